"""CPU oracle for the tgcn time-vertex Chebyshev graph-convolution hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in ``tgcn_b200/`` (the product) may import
this package.  Allowed importers: ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs -- there only as the
checker or the timed CPU baseline, never as the thing shipped.

Contents (each function cites the reference file:line it restates; reference
root = cassianobecker/tgcn):

* ``layers_np``   numpy float64 restatement of ``tgcn/nn/gcn.py`` layers
                  (forward + analytic backward) and of ``gcn_pool*``.
* ``layers_torch`` functional torch-CPU port that performs the same ATen calls
                  as the reference (dense-L einsum, stacked basis, autograd) --
                  this is what ``bench.py`` times as the CPU baseline ("port").
* ``graph_np``    restatement of ``gcn/graph.py`` (grid, kNN, adjacency,
                  laplacian, rescale_L).
* ``coarsening_np`` restatement of ``gcn/coarsening.py`` (metis, compute_perm,
                  perm_adjacency, coarsen, perm_data).
* ``model_torch`` the reference's model compositions (pytorch_hcp_tgcn.py:93-155,
                  pytorch_mnist_tgcn.py:67-92) over ``layers_torch``.

Everything is Python (numpy / scipy / torch-CPU), like the reference itself: there is
no C restatement and nothing to build.

Parity pinning: the reference has exactly one known-answer test on this path
(``gcn/coarsening.py:216-217``, the ``compute_perm`` KAT) -- checked in
``tests/test_oracle_coarsening.py``.  Everything else is pinned against outputs
of the UNMODIFIED reference run in the build container and committed under
``tests/golden/*.npz`` together with the generating script
``tests/golden/make_golden.py``.
"""

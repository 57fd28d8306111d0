"""tgcn_b200 -- B200-native (sm_100a) time-vertex Chebyshev graph-convolution hot path.

Drop-in for the `tgcn/nn` PyTorch layers of cassianobecker/tgcn:

    from tgcn_b200.nn.gcn import TGCNCheb, TGCNCheb_H, GCNCheb, gcn_pool, gcn_pool_4

plus the host-side input producers (`tgcn_b200.graph`, `tgcn_b200.coarsening`) with the
reference's function names.  All device work runs in hand-written CUDA kernels behind the C-ABI
in include/tgcn_b200.h; there is no CPU fallback.
"""
__version__ = "0.1.0"

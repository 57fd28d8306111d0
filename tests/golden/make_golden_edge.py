"""Golden outputs of the reference's edge-index operators (tgcn/nn/gcn.py:348-538, ChebConv / ChebTimeConv).

The reference imports three functions from packages that are NOT under /root/reference and are not
installed in this image (the reference pins no versions; current upstream releases are named):
  * torch_geometric.utils.degree(index, num_nodes, dtype)      (PyG 2.x)   -- occurrences of each index
  * torch_geometric.utils.remove_self_loops(edge_index, attr)  (PyG 2.x)   -- drop edges with row == col
  * torch_scatter.scatter_add(src, index, dim, dim_size)       (2.1.x)     -- out.index_add_(dim, index, src)
They are restated below from their published documentation and injected as the stub modules' attributes;
everything else (the operators themselves) is the UNMODIFIED reference code, run on CPU.

    python tests/golden/make_golden_edge.py     # writes tests/golden/edge_*.npz
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_loader  # noqa: E402


def degree(index, num_nodes=None, dtype=None):
    n = int(index.max()) + 1 if num_nodes is None else num_nodes
    out = torch.zeros((n,), dtype=dtype, device=index.device)
    return out.scatter_add_(0, index, out.new_ones((index.size(0),)))


def remove_self_loops(edge_index, edge_attr=None):
    mask = edge_index[0] != edge_index[1]
    edge_index = edge_index[:, mask]
    return (edge_index, None) if edge_attr is None else (edge_index, edge_attr[mask])


def scatter_add(src, index, dim=-1, out=None, dim_size=None):
    shape = list(src.shape)
    shape[dim] = dim_size if dim_size is not None else int(index.max()) + 1
    out = torch.zeros(shape, dtype=src.dtype, device=src.device)
    return out.index_add_(dim, index, src)


def main():
    g, c, n, m = ref_loader.load()
    n.degree, n.remove_self_loops, n.scatter_add = degree, remove_self_loops, scatter_add
    rng = np.random.default_rng(0)
    cases = {
        # name: (cls, N, E, Q, F, G, K, H, weighted, bias)
        "chebconv_w": ("ChebConv", 40, 160, 3, 4, 6, 5, None, True, True),
        "chebconv_unweighted_2d": ("ChebConv", 25, 90, 2, 1, 3, 4, None, False, True),
        "chebconv_k1_nobias": ("ChebConv", 12, 30, 2, 3, 2, 1, None, True, False),
        "chebtime_w": ("ChebTimeConv", 30, 120, 3, 2, 5, 6, 7, True, True),
        "chebtime_3d": ("ChebTimeConv", 20, 70, 2, 1, 4, 3, 5, False, True),
    }
    for name, (cls, N, E, Q, F, G, K, H, weighted, bias) in cases.items():
        row = rng.integers(0, N, E)
        col = rng.integers(0, N, E)            # random directed multigraph: self-loops and duplicates included
        ei = torch.tensor(np.stack([row, col]), dtype=torch.long)
        ew = torch.tensor(rng.random(E).astype(np.float32) + 0.1) if weighted else None
        torch.manual_seed(1)
        if cls == "ChebConv":
            lay = n.ChebConv(F, G, K, bias=bias)
            shape = (Q, N) if name.endswith("2d") else (Q, N, F)
        else:
            lay = n.ChebTimeConv(F, G, K, H, bias=bias)
            shape = (Q, N, H) if name.endswith("3d") else (Q, N, H, F)
        x = torch.tensor(rng.standard_normal(shape).astype(np.float32), requires_grad=True)
        out = lay(x, ei, ew)
        dout = torch.tensor(rng.standard_normal(tuple(out.shape)).astype(np.float32))
        out.backward(dout)
        rec = dict(cls=cls, edge_index=ei.numpy(), x=x.detach().numpy(), W=lay.weight.detach().numpy(),
                   out=out.detach().numpy(), dout=dout.numpy(), dW=lay.weight.grad.numpy(), dx=x.grad.numpy(),
                   K=K, F=F, G=G, H=-1 if H is None else H)
        if ew is not None:
            rec["edge_weight"] = ew.numpy()
        if bias:
            rec["b"] = lay.bias.detach().numpy()
            rec["db"] = lay.bias.grad.numpy()
        np.savez_compressed(os.path.join(HERE, "edge_%s.npz" % name), **rec)
        print("wrote edge_%s.npz" % name, tuple(out.shape))


if __name__ == "__main__":
    main()

"""Sample-resident fused layer kernels (engine="resident", csrc/resident.cu) against the goldens of
the unmodified reference, the float64 oracle, and the streaming path; fused ReLU + max-pool epilogue
against the unfused chain (values, indices and gradients).

Tolerance (north_star): 1e-4 relative (conftest.rel_err) for activations and gradients; pool
indices bit-exact.
"""
import numpy as np
import pytest
import torch

from conftest import LAYER_CASES, csr_from, load_golden, rel_err
from oracle import layers_np

pytestmark = pytest.mark.gpu
TOL = 1e-4
KIND = {"TGCNCheb_H": "tgcn_h", "GCNCheb": "gcn", "TGCNCheb": "tgcn"}


def make_layer(r, L, **kw):
    from tgcn_b200.nn import gcn as G
    cls = getattr(G, str(r["cls"]))
    K, fin, fout, H = int(r["K"]), int(r["in_ch"]), int(r["out_ch"]), int(r["H"])
    bias = "b" in r
    lay = cls(L, fin, fout, K, H, bias=bias, **kw) if str(r["cls"]) == "TGCNCheb_H" else cls(L, fin, fout, K, bias=bias, **kw)
    with torch.no_grad():
        lay.weight.copy_(torch.tensor(r["W"]))
        if bias:
            lay.bias.copy_(torch.tensor(r["b"]))
    return lay.cuda()


@pytest.mark.parametrize("case", LAYER_CASES)
def test_resident_layer_vs_reference_golden(case):
    r = load_golden(case)
    Lsp = csr_from(r, "L")
    dense = torch.tensor(np.asarray(Lsp.todense()), dtype=torch.float)
    lay = make_layer(r, dense, engine="resident")
    x = torch.tensor(r["x"], device="cuda", requires_grad=True)
    out = lay(x)
    assert out.shape == r["out"].shape and out.is_contiguous()
    assert rel_err(out.detach().cpu().numpy(), r["out"]) < TOL
    out.backward(torch.tensor(r["dout"], device="cuda"))
    assert rel_err(lay.weight.grad.cpu().numpy(), r["dW"]) < TOL
    assert rel_err(x.grad.cpu().numpy(), r["dx"]) < TOL
    if "b" in r:
        assert lay.bias.grad.shape == r["db"].shape
        assert rel_err(lay.bias.grad.cpu().numpy(), r["db"]) < TOL


@pytest.mark.parametrize("case", LAYER_CASES)
@pytest.mark.parametrize("recursion", ["reference", "chebyshev"])
def test_resident_layer_vs_numpy_oracle(case, recursion):
    r = load_golden(case)
    Lsp = csr_from(r, "L")
    kind = KIND[str(r["cls"])]
    rng = np.random.default_rng(11)
    x = rng.standard_normal(r["x"].shape).astype(np.float32)
    W = rng.standard_normal(r["W"].shape).astype(np.float32) * 0.3
    b = rng.standard_normal(r["b"].shape).astype(np.float32) if "b" in r else None
    dout = rng.standard_normal(r["out"].shape).astype(np.float32)
    r2 = dict(r); r2["W"] = W
    if b is not None:
        r2["b"] = b
    lay = make_layer(r2, torch.tensor(np.asarray(Lsp.todense()), dtype=torch.float), recursion=recursion, engine="resident")
    xt = torch.tensor(x, device="cuda", requires_grad=True)
    out = lay(xt)
    out.backward(torch.tensor(dout, device="cuda"))
    ref_out = layers_np.layer_forward(Lsp, x, W, b, kind=kind, recursion=recursion)
    dW, db, dx = layers_np.layer_backward(Lsp, x, W, dout, None if b is None else b.shape, kind=kind, recursion=recursion)
    assert rel_err(out.detach().cpu().numpy(), ref_out) < TOL
    assert rel_err(lay.weight.grad.cpu().numpy(), dW) < TOL
    assert rel_err(xt.grad.cpu().numpy(), dx) < TOL
    if b is not None:
        assert rel_err(lay.bias.grad.cpu().numpy(), db) < TOL


def _rand_graph(n, density, rng, symmetric=True):
    A = (rng.random((n, n)) < density) * rng.random((n, n))
    if symmetric:
        A = np.maximum(A, A.T)
    np.fill_diagonal(A, 0)
    d = A.sum(0) + 1e-30
    return (-(A / np.sqrt(d)[:, None]) / np.sqrt(d)[None, :]).astype(np.float32)


SHAPES = [  # cls, Q, N, H, F, G, K, density
    ("TGCNCheb_H", 64, 384, 15, 1, 32, 10, 0.045),     # hcp360 layer 1
    ("GCNCheb", 64, 96, 1, 32, 64, 10, 0.48),          # hcp360 layer 2
    ("TGCNCheb_H", 100, 992, 12, 1, 15, 10, 0.0065),   # mnist layer 1 (G = 15: padded filter tile)
    ("TGCNCheb_H", 5, 368, 15, 1, 32, 4, 1.0),         # fully dense L~: CSR stays in global memory
    ("TGCNCheb", 3, 30, 1, 7, 9, 3, 0.3),              # N % 4 != 0, odd D and G
    ("GCNCheb", 2, 18, 1, 5, 6, 1, 0.3),               # K = 1
    ("TGCNCheb_H", 4, 24, 3, 5, 8, 20, 0.2),           # K = 20 (> one reduce pass)
    ("GCNCheb", 1, 4, 1, 1, 1, 2, 1.0),                # degenerate
]


def _build(shape, recursion, engine, seed=0, bias=True):
    from tgcn_b200.nn import gcn as Gm
    cls, Q, N, H, F, G, K, dens = shape
    rng = np.random.default_rng(seed)
    L = _rand_graph(N, dens, rng, symmetric=(seed % 2 == 0))
    Lt = torch.tensor(L)
    torch.manual_seed(seed)
    if cls == "TGCNCheb_H":
        lay = Gm.TGCNCheb_H(Lt, F, G, K, H, bias=bias, recursion=recursion, engine=engine).cuda()
        x = rng.standard_normal((Q, N, H, F)).astype(np.float32)
        kind = "tgcn_h"
    elif cls == "TGCNCheb":
        lay = Gm.TGCNCheb(Lt, F, G, K, bias=bias, recursion=recursion, engine=engine).cuda()
        x = rng.standard_normal((Q, N, F)).astype(np.float32)
        kind = "tgcn"
    else:
        lay = Gm.GCNCheb(Lt, F, G, K, bias=bias, recursion=recursion, engine=engine).cuda()
        x = rng.standard_normal((Q, N, F)).astype(np.float32)
        kind = "gcn"
    return lay, L, x, kind


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("recursion", ["reference", "chebyshev"])
@pytest.mark.parametrize("seed", [0, 1])          # seed 1: non-symmetric L~ (exercises the true transpose)
def test_resident_shapes_vs_oracle(shape, recursion, seed):
    lay, L, x, kind = _build(shape, recursion, "resident", seed=seed, bias=(seed == 0))
    xt = torch.tensor(x, device="cuda", requires_grad=True)
    out = lay(xt)
    rng = np.random.default_rng(5)
    dout = rng.standard_normal(tuple(out.shape)).astype(np.float32)
    out.backward(torch.tensor(dout, device="cuda"))
    W = lay.weight.detach().cpu().numpy()
    b = None if lay.bias is None else lay.bias.detach().cpu().numpy()
    ref = layers_np.layer_forward(L, x, W, b, kind=kind, recursion=recursion)
    dW, db, dx = layers_np.layer_backward(L, x, W, dout, None if b is None else b.shape, kind=kind, recursion=recursion)
    assert rel_err(out.detach().cpu().numpy(), ref) < TOL
    assert rel_err(lay.weight.grad.cpu().numpy(), dW) < TOL
    assert rel_err(xt.grad.cpu().numpy(), dx) < TOL
    if b is not None:
        assert rel_err(lay.bias.grad.cpu().numpy(), db) < TOL


@pytest.mark.parametrize("shape", SHAPES[:3] + SHAPES[4:5])
@pytest.mark.parametrize("p", [2, 4])
def test_fused_relu_pool_matches_unfused_chain(shape, p):
    """forward_relu_pool == gcn_pool(F.relu(layer(x))): values bit-equal to pooling the layer's own
    output, indices equal to the oracle's torch.max rule, gradients equal to the unfused chain."""
    from tgcn_b200.nn import gcn as Gm
    if shape[2] % p:
        pytest.skip("N not divisible by the pool size")
    lay, L, x, kind = _build(shape, "reference", "resident", seed=0)
    xa = torch.tensor(x, device="cuda", requires_grad=True)
    y = lay.forward_relu_pool(xa, p)
    g = torch.Generator(device="cuda").manual_seed(3)
    dy = torch.randn(y.shape, device="cuda", generator=g)
    y.backward(dy)
    gW, gb, gx = lay.weight.grad.clone(), lay.bias.grad.clone(), xa.grad.clone()
    lay.zero_grad()
    xb = torch.tensor(x, device="cuda", requires_grad=True)
    out = lay(xb)
    y2, idx2 = Gm.gcn_pool_with_indices(torch.relu(out), p)
    assert torch.equal(y, y2)
    pv, pi = layers_np.pool_forward(np.maximum(out.detach().cpu().numpy(), 0), p)
    assert np.array_equal(y.detach().cpu().numpy(), pv)
    assert np.array_equal(idx2.cpu().numpy(), pi)
    y2.backward(dy)
    assert rel_err(gW.cpu().numpy(), lay.weight.grad.cpu().numpy()) < 1e-6
    assert rel_err(gb.cpu().numpy(), lay.bias.grad.cpu().numpy()) < 1e-6
    assert rel_err(gx.cpu().numpy(), xb.grad.cpu().numpy()) < 1e-6


def test_fused_pool_indices_and_nan_rule():
    """Ties pick the first sibling, NaN wins (torch.max rule); the ReLU mask follows autograd:
    zero gradient where the selected activation is <= 0, NaN passes."""
    from tgcn_b200 import _lib
    from tgcn_b200.csr import build_csr
    lib = _lib.load()
    N, Q, D, G, K = 8, 2, 1, 4, 1
    plan = build_csr(torch.zeros(N, N), torch.device("cuda"))
    x = torch.zeros(Q, N, D, device="cuda")
    W = torch.zeros(K, D, G, device="cuda")
    bias = torch.tensor([[1, 1, -1, float("nan")], [1, 2, -2, 0], [1, 2, -1, 5], [0, float("nan"), -3, 5],
                         [0, 0, 0, 0], [float("inf"), -1, 0, 0], [float("inf"), -1, 0, 1], [1, -1, 0, 1]],
                        dtype=torch.float32, device="cuda")
    y = torch.empty(Q, N // 4, G, device="cuda")
    idx = torch.empty(Q, N // 4, G, dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    stack = torch.empty(int(lib.tgcn_resident_stack_bytes(Q, N, D, K)) // 4, device="cuda")
    wimg = torch.empty(int(lib.tgcn_resident_weights_bytes(D, G, K)) // 4, device="cuda")
    rowinfo, entries, E = plan.packed(lib.tgcn_resident_pack_classes(Q, N, D, 0))
    rc = lib.tgcn_resident_layer_fwd(rowinfo.data_ptr(), entries.data_ptr(), N, E,
                                     x.data_ptr(), W.data_ptr(), bias.data_ptr(), 1, None, y.data_ptr(), idx.data_ptr(), 4, 1, None,
                                     stack.data_ptr(), wimg.data_ptr(), Q, D, G, K, 0, st)
    assert rc == 0, _lib.last_error()
    ref_v, ref_i = torch.max(torch.relu(bias).reshape(1, N // 4, 4, G).expand(Q, -1, -1, -1), dim=2)
    assert torch.equal(idx.long(), ref_i)
    assert torch.equal(torch.nan_to_num(y, nan=-7.0), torch.nan_to_num(ref_v, nan=-7.0))
    # gradient routing against autograd on the same activations
    a = bias.clone().requires_grad_(True)
    v, _ = torch.max(torch.relu(a).reshape(1, N // 4, 4, G), dim=2)
    dy = torch.arange(1, 1 + (N // 4) * G, dtype=torch.float32, device="cuda").reshape(1, N // 4, G)
    v.backward(dy)
    dW = torch.empty(K, D, G, device="cuda")
    db = torch.empty(N, G, device="cuda")
    ws = torch.empty(max(int(lib.tgcn_resident_bwd_workspace(Q, N, D, G, K)) // 4, 1), device="cuda")
    dyq = dy.expand(Q, -1, -1).contiguous()
    rc = lib.tgcn_resident_layer_bwd(rowinfo.data_ptr(), entries.data_ptr(), N, E,
                                     None, dyq.data_ptr(), idx.data_ptr(), y.data_ptr(), 4, 1, None, stack.data_ptr(), wimg.data_ptr(),
                                     dW.data_ptr(), db.data_ptr(), 1, None, ws.data_ptr(), Q, D, G, K, 0, st)
    assert rc == 0, _lib.last_error()
    assert torch.equal(torch.nan_to_num(db, nan=-7.0), torch.nan_to_num(a.grad * Q, nan=-7.0))


def test_resident_is_deterministic_and_batch_independent():
    """Run-to-run bit stability, and a sample's output does not depend on the batch it rides in."""
    lay, L, x, kind = _build(SHAPES[0], "reference", "resident")
    xt = torch.tensor(x, device="cuda")
    with torch.no_grad():
        a = lay(xt)
        b = lay(xt)
        c = lay(xt[5:9].contiguous())
    assert torch.equal(a, b)
    assert torch.equal(a[5:9], c)
    xg = torch.tensor(x, device="cuda", requires_grad=True)
    grads = []
    for _ in range(2):
        lay.zero_grad()
        lay(xg).square().sum().backward()
        grads.append((lay.weight.grad.clone(), lay.bias.grad.clone()))
    assert torch.equal(grads[0][0], grads[1][0]) and torch.equal(grads[0][1], grads[1][1])


def test_resident_matches_streaming_path():
    for shape in SHAPES[:2]:
        lr, L, x, kind = _build(shape, "reference", "resident")
        ls, _, _, _ = _build(shape, "reference", "ffma")
        ls.load_state_dict(lr.state_dict())
        xa = torch.tensor(x, device="cuda", requires_grad=True)
        xb = torch.tensor(x, device="cuda", requires_grad=True)
        oa, ob = lr(xa), ls(xb)
        assert rel_err(oa.detach().cpu().numpy(), ob.detach().cpu().numpy()) < 1e-5
        g = torch.randn(oa.shape, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))
        oa.backward(g); ob.backward(g)
        assert rel_err(lr.weight.grad.cpu().numpy(), ls.weight.grad.cpu().numpy()) < 1e-5
        assert rel_err(xa.grad.cpu().numpy(), xb.grad.cpu().numpy()) < 1e-5


def test_resident_refuses_what_does_not_fit():
    from tgcn_b200 import _lib
    from tgcn_b200.nn import gcn as Gm
    lib = _lib.load()
    assert lib.tgcn_resident_supported(384, 15, 32, 10, 6540) == 1
    assert lib.tgcn_resident_supported(41856, 30, 32, 10, 194940) == 0
    n = 6000
    L = torch.eye(n).to_sparse()
    lay = Gm.TGCNCheb_H(L, 1, 8, 3, 30, engine="resident").cuda()
    with pytest.raises(RuntimeError, match="resident"):
        lay(torch.zeros(1, n, 30, device="cuda"))


@pytest.mark.parametrize("N", [20, 4, 12])
def test_cluster_split_with_an_empty_trailing_rank(N):
    """Q small enough for a 4-CTA cluster while ceil(row_groups / 4) leaves the last rank without rows
    (N = 20: 5 row groups -> 2, 2, 1, 0)."""
    shape = ("TGCNCheb_H", 2, N, 5, 1, 4, 3, 0.4)
    lay, L, x, kind = _build(shape, "chebyshev", "resident", seed=0)
    xt = torch.tensor(x, device="cuda", requires_grad=True)
    out = lay(xt)
    dout = np.random.default_rng(1).standard_normal(tuple(out.shape)).astype(np.float32)
    out.backward(torch.tensor(dout, device="cuda"))
    W, b = lay.weight.detach().cpu().numpy(), lay.bias.detach().cpu().numpy()
    ref = layers_np.layer_forward(L, x, W, b, kind=kind, recursion="chebyshev")
    dW, db, dx = layers_np.layer_backward(L, x, W, dout, b.shape, kind=kind, recursion="chebyshev")
    assert rel_err(out.detach().cpu().numpy(), ref) < TOL
    assert rel_err(lay.weight.grad.cpu().numpy(), dW) < TOL
    assert rel_err(xt.grad.cpu().numpy(), dx) < TOL


@pytest.mark.parametrize("shape", SHAPES[:2] + SHAPES[4:5] + SHAPES[6:7])
def test_tcgen05_and_ffma_contraction_of_the_resident_forward_agree(shape):
    """RES_TC tuning key: the contraction of the resident forward kernel on tcgen05 (3xTF32, TMEM accumulators)
    against the same kernel with the fp32 FFMA contraction, and both against the float64 oracle."""
    from tgcn_b200 import _lib
    lib = _lib.load()
    lay, L, x, kind = _build(shape, "reference", "resident", seed=0)
    xt = torch.tensor(x, device="cuda")
    outs = {}
    try:
        for tc in (1, 0):
            assert lib.tgcn_set_tuning(b"RES_TC", tc) == 0
            with torch.no_grad():
                outs[tc] = lay(xt).cpu().numpy()
    finally:
        lib.tgcn_set_tuning(b"RES_TC", -1)
    W, b = lay.weight.detach().cpu().numpy(), lay.bias.detach().cpu().numpy()
    ref = layers_np.layer_forward(L, x, W, b, kind=kind, recursion="reference")
    assert rel_err(outs[0], ref) < 2e-6
    assert rel_err(outs[1], ref) < 2e-5
    assert rel_err(outs[1], outs[0]) < 2e-5


def test_register_cached_csr_variant_is_bit_identical():
    """RES_ENT tuning key: the forward variant that keeps each thread's CSR entries in registers across the K
    steps follows the same even/odd summation order, so its output is bit-identical to the default kernel."""
    from tgcn_b200 import _lib
    lib = _lib.load()
    for shape in (SHAPES[0], SHAPES[1], SHAPES[4]):
        lay, L, x, kind = _build(shape, "reference", "resident", seed=0)
        xt = torch.tensor(x, device="cuda")
        outs = []
        try:
            for v in (0, 1):
                assert lib.tgcn_set_tuning(b"RES_ENT", v) == 0
                with torch.no_grad():
                    outs.append(lay(xt))
        finally:
            lib.tgcn_set_tuning(b"RES_ENT", -1)
        assert torch.equal(outs[0], outs[1])

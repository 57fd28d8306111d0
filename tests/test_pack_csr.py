"""Host-side packing of the CSR operand for the sample-resident kernels (tgcn_pack_csr_host):
a pure CPU function of the C-ABI, checked against a numpy restatement."""
import numpy as np
import pytest
import scipy.sparse as sp

from tgcn_b200 import _lib


def _pack(m, classes):
    lib = _lib.load()
    m = m.tocsr(); m.sort_indices()
    rp = m.indptr.astype(np.int32); c = m.indices.astype(np.int32); v = m.data.astype(np.float32)
    n = m.shape[0]
    rowinfo = np.zeros(((n + 1) & ~1, 2), np.int32)      # an odd N needs one zero padding row (16-byte multiple)
    E = lib.tgcn_pack_csr_host(rp.ctypes.data, c.ctypes.data, v.ctypes.data, n, classes, rowinfo.ctypes.data, None)
    entries = np.full((max(E, 1), 2), -7, np.int32)
    E2 = lib.tgcn_pack_csr_host(rp.ctypes.data, c.ctypes.data, v.ctypes.data, n, classes, rowinfo.ctypes.data, entries.ctypes.data)
    assert E == E2
    return rp, c, v, rowinfo, entries, E


@pytest.mark.parametrize("classes", [1, 2, 4, 8])
def test_pack_is_a_row_wise_permutation_with_even_starts(classes):
    rng = np.random.default_rng(classes)
    m = sp.random(57, 57, density=0.2, random_state=3, format="csr", dtype=np.float32)
    rp, c, v, rowinfo, entries, E = _pack(m, classes)
    assert E % 2 == 0 and E == int(sum((l + 1) & ~1 for l in np.diff(rp)))
    for n in range(57):
        start, ln = rowinfo[n]
        assert start % 2 == 0 and ln == rp[n + 1] - rp[n]
        got_c = entries[start:start + ln, 0]
        got_v = entries[start:start + ln, 1].view(np.float32)
        # same multiset of (col, val) pairs
        ref = sorted(zip(c[rp[n]:rp[n + 1]].tolist(), v[rp[n]:rp[n + 1]].tolist()))
        assert sorted(zip(got_c.tolist(), got_v.tolist())) == ref
        # ordered by bank class starting at the row's own class, stable inside a class
        key = (got_c - n) % classes
        assert np.all(np.diff(key) >= 0)
        for k in range(classes):
            assert np.all(np.diff(got_c[key == k]) > 0)


def test_pack_rejects_bad_arguments_and_handles_empty_rows():
    lib = _lib.load()
    m = sp.csr_matrix((6, 6), dtype=np.float32)
    rp, c, v, rowinfo, entries, E = _pack(m, 2)
    assert E == 0 and np.all(rowinfo[:, 1] == 0)
    rp = np.zeros(7, np.int32)
    assert lib.tgcn_pack_csr_host(rp.ctypes.data, None, None, 6, 3, None, None) == -1
    assert lib.tgcn_resident_pack_classes(64, 384, 15, 0) == 2      # 4 float4 per row -> rows alternate bank halves
    assert lib.tgcn_resident_pack_classes(64, 96, 32, 0) == 1       # 128-byte rows: no conflicts to avoid

"""The C-ABI library loads and exports every symbol include/tgcn_b200.h declares (no compute calls)."""
import ctypes
import os
import re

from conftest import ROOT
from tgcn_b200 import _lib


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "tgcn_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tgcn_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_documented_surface():
    names = declared_symbols()
    for must in ("tgcn_spmm_step", "tgcn_contract_fwd", "tgcn_contract_bwd_w", "tgcn_contract_bwd_x",
                 "tgcn_pool_max_fwd", "tgcn_pool_max_bwd", "tgcn_layer_fwd", "tgcn_layer_bwd", "tgcn_last_error"):
        assert must in names


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(raw, name), "missing export: " + name
        assert name in _lib.SIGNATURES, "ctypes signature missing for " + name
    assert set(_lib.SIGNATURES) == set(declared_symbols())
    assert lib.tgcn_version() >= 100
    assert isinstance(_lib.last_error(), str)


def test_workspace_queries_are_pure_host_calls():
    lib = _lib.load()
    assert lib.tgcn_layer_bwd_workspace(8, 1000, 30, 32, 10) > 0
    assert lib.tgcn_contract_bwd_w_workspace(0, 0, 1, 1, 1) >= 0
    assert lib.tgcn_layer_bwd_workspace(-1, 5, 1, 1, 1) == 0


def test_slab_width_policy_is_a_pure_host_call():
    """tgcn_layer_slab_width (include/tgcn_b200.h): D padded to whole 128-byte blocks only when that costs <= 12.5 % and the
    tensor-core engine can run; never below D; the FFMA engine never pads."""
    lib = _lib.load()
    assert lib.tgcn_layer_slab_width(8, 41856, 30, 32, 10, _lib.ENGINE_AUTO) == 32        # cortical mesh layer 1
    assert lib.tgcn_layer_slab_width(8, 41856, 30, 32, 10, _lib.ENGINE_FFMA) == 30
    assert lib.tgcn_layer_slab_width(8, 10464, 32, 64, 10, _lib.ENGINE_AUTO) == 32        # already whole blocks
    assert lib.tgcn_layer_slab_width(1, 1000000, 192, 64, 8, _lib.ENGINE_AUTO) == 192
    assert lib.tgcn_layer_slab_width(8, 41856, 5, 32, 10, _lib.ENGINE_AUTO) == 5          # 5 -> 32 would cost 540 %
    assert lib.tgcn_layer_slab_width(8, 41856, 28, 32, 10, _lib.ENGINE_AUTO) == 28        # 14 % > 12.5 %
    assert lib.tgcn_layer_slab_width(8, 41856, 29, 32, 10, _lib.ENGINE_AUTO) == 32
    for D in (1, 7, 30, 33, 60, 100, 191):
        for eng in (_lib.ENGINE_AUTO, _lib.ENGINE_FFMA, _lib.ENGINE_TCGEN05):
            Dp = lib.tgcn_layer_slab_width(4, 5000, D, 32, 5, eng)
            assert D <= Dp <= (D + 31) // 32 * 32 and (Dp == D or Dp % 32 == 0)
    assert lib.tgcn_head_workspace(8, 167424, 200) > 0                                   # large head: partial sums
    assert lib.tgcn_head_workspace(64, 1536, 200) == 0                                   # small head: none
    assert lib.tgcn_head_fused_update_supported(8, 167424, 200) == 1
    assert lib.tgcn_head_fused_update_supported(64, 1536, 200) == 0

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

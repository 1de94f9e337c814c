"""CPU checks of the drop-in boundary: the C-ABI library builds, loads and exports every
symbol include/azgnn_b200.h declares; the Python binding table covers them all; compute
entry points refuse to run without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import pytest
import torch

import azgnn_b200
from azgnn_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "azgnn_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(azg_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from importlib import import_module
    build = import_module("azgnn_b200.build")
    path = build.build()
    cdll = ctypes.CDLL(path)
    syms = declared_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(cdll, s), f"{s} declared in include/azgnn_b200.h but not exported"
    assert cdll.azg_abi_version() == 3


def test_binding_table_matches_header():
    assert sorted(_lib.SIGNATURES) == declared_symbols()


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    from azgnn_b200.nets import B200Connect4GNNWrapper
    from azgnn_b200.arena import DeviceArena

    class G:
        def getBoardSize(self):
            return (7, 7)

        def getActionSize(self):
            return 8
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        B200Connect4GNNWrapper(G(), dict(lr=1e-3))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        DeviceArena("connect4", 7, 4, 10, 1.0)

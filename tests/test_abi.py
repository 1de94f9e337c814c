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


def test_header_constants_match_the_python_binding():
    """#define AZG_PREC_* / AZG_EVAL_* in include/azgnn_b200.h == the integers `_lib.py` passes through ctypes; every precision
    name a config may use (`b200_precision`) maps to one of them; precisions that share another precision's packed weight
    images (`PACKED_AS`) point at a real precision."""
    import os
    import re
    from azgnn_b200 import _lib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    text = open(os.path.join(root, "include", "azgnn_b200.h")).read()
    defs = {m.group(1): int(m.group(2)) for m in re.finditer(r"#define\s+(AZG_(?:PREC|EVAL)_\w+)\s+(-?\d+)", text)}
    want = {"AZG_PREC_FP32": _lib.PREC_FP32, "AZG_PREC_BF16X3": _lib.PREC_BF16X3, "AZG_PREC_BF16": _lib.PREC_BF16,
            "AZG_PREC_F16F8": _lib.PREC_F16F8, "AZG_PREC_F16F8_KS": _lib.PREC_F16F8_KS, "AZG_PREC_BF16X3_KS": _lib.PREC_BF16X3_KS,
            "AZG_EVAL_STD": _lib.EVAL_STD, "AZG_EVAL_GNN": _lib.EVAL_GNN, "AZG_EVAL_FOLD": _lib.EVAL_FOLD}
    assert {k: defs.get(k) for k in want} == want
    assert {k for k in defs if k.startswith("AZG_PREC_")} == {k for k in want if k.startswith("AZG_PREC_")}
    assert set(_lib.PRECISIONS.values()) == {v for k, v in want.items() if k.startswith("AZG_PREC_")} | {_lib.PREC_AUTO}
    assert all(src in _lib.PRECISION_NAMES and dst in _lib.PRECISION_NAMES for src, dst in _lib.PACKED_AS.items())

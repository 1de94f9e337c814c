"""TEST INFRASTRUCTURE: the arena's host/device core compiled for x86 (libazg_hostcheck.so),
driven through the SAME Python code as the CUDA arena (azgnn_b200.arena.DeviceArena), with
torch CPU tensors instead of CUDA tensors.  Lets the arena source be checked bit for bit
against the oracle on machines without a GPU.  Never used by the package itself."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import build as _build  # noqa: E402

from azgnn_b200 import _lib  # noqa: E402
from azgnn_b200.arena import DeviceArena  # noqa: E402

_NAMES = [n for n in _lib.SIGNATURES if n.startswith("azg_arena_") or n in ("azg_rules_eval", "azg_last_error")]


class _HostLib:
    def __init__(self):
        self.cdll = C.CDLL(_build.build())
        for name in _NAMES:
            fn = getattr(self.cdll, name.replace("azg_", "azgh_", 1))
            fn.restype, fn.argtypes = _lib.SIGNATURES[name]
            setattr(self, name, fn)
        self.cdll.azgh_np_sum.restype = C.c_double
        self.cdll.azgh_np_sum.argtypes = [C.c_void_p, C.c_int]


_hostlib = None


def hostlib():
    global _hostlib
    if _hostlib is None:
        _hostlib = _HostLib()
    return _hostlib


class HostArena(DeviceArena):
    def _open(self, device):
        self.lib = hostlib()
        self.device = torch.device("cpu")

    def _stream(self):
        return C.c_void_p(None)

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError(f"hostcheck arena error {rc}: {self.lib.azg_last_error().decode()}")

    def to_host(self, t):
        return t.detach().numpy().copy()  # CPU tensors would otherwise alias the arena's buffers

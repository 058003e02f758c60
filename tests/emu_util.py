"""Run libpgpd's CUDA-core kernels + host orchestration on the CPU through the SIMT emulator
(tests/simt_emu).  TEST INFRASTRUCTURE ONLY: numpy arrays stand in for device memory."""
import ctypes as C
import importlib.util
import os

import numpy as np

from pointnetgpd_b200 import _abi as A

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def emu_lib():
    global _lib
    if _lib is None:
        spec = importlib.util.spec_from_file_location("emu_build", os.path.join(_HERE, "simt_emu", "build.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        _lib = A.bind(C.CDLL(mod.build()))
    return _lib


def _addr(a):
    return a.ctypes.data


class Guarded:
    """numpy buffer with NaN-poisoned guard zones, to catch out-of-bounds writes of a kernel."""
    PAD = 64

    def __init__(self, nbytes, align=256):
        self.raw = np.zeros(nbytes + 2 * self.PAD * 4 + align, dtype=np.uint8)
        base = self.raw.ctypes.data + self.PAD * 4
        self.off = (-base) % align + self.PAD * 4
        self.nbytes = nbytes
        self.raw[:] = 0xA5
        self.snapshot_lo = self.raw[:self.off].copy()
        self.snapshot_hi = self.raw[self.off + nbytes:].copy()

    @property
    def addr(self):
        return self.raw.ctypes.data + self.off

    def check(self):
        assert (self.raw[:self.off] == self.snapshot_lo).all(), "write before the workspace"
        assert (self.raw[self.off + self.nbytes:] == self.snapshot_hi).all(), "write past the workspace"


def run_model(state, x, what=A.PGPD_CLS, train=True, dout=None, dtrans=None, backward=False, flags_extra=0, split_backward=False):
    """state: dict of numpy arrays (float32 / int64), modified in place for running stats.
    Returns dict(out, trans, grads)."""
    lib = emu_lib()
    st = {k: np.ascontiguousarray(v) for k, v in state.items()}
    for k in list(st):
        if st[k].ndim == 0:
            st[k] = st[k].reshape(1)
    B, _, N = x.shape
    k = int(st["fc3.weight"].shape[0]) if what == A.PGPD_CLS else 1
    x = np.ascontiguousarray(x, dtype=np.float32)
    flags = (A.F_TRAIN if train else 0) | (A.F_SAVE if backward else 0) | flags_extra
    model = A.build_model(lambda key: _addr(st[key]), what)
    nbytes = lib.pgpd_workspace_bytes(what, B, N, k, flags)
    ws = Guarded(nbytes)
    out = np.full((B, k if what == A.PGPD_CLS else 1024), np.nan, dtype=np.float32)
    trans = np.full((B, 3, 3), np.nan, dtype=np.float32)
    rc = lib.pgpd_forward(what, C.byref(model), _addr(x), B, N, k, flags, _addr(out), _addr(trans),
                          ws.addr, nbytes, None)
    A.check(lib, rc)
    ws.check()
    res = dict(out=out, trans=trans, state=st)
    if backward:
        grads = {key: np.full(st[key].shape, np.nan, dtype=np.float32) for key in A.param_keys(what)}
        g = A.build_grads(lambda key: _addr(grads[key]), what)
        dout_a = None if dout is None else np.ascontiguousarray(dout, dtype=np.float32)
        dtr_a = None if dtrans is None else np.ascontiguousarray(dtrans, dtype=np.float32)
        for extra in ((A.F_BWD_HEAD, A.F_BWD_STN) if split_backward else (0,)):     # two calls: the halves a data-parallel trainer overlaps
            rc = lib.pgpd_backward(what, C.byref(model), C.byref(g), _addr(x), B, N, k, flags | extra,
                                   None if dout_a is None else _addr(dout_a),
                                   None if dtr_a is None else _addr(dtr_a), ws.addr, nbytes, None)
            A.check(lib, rc)
        ws.check()
        res["grads"] = grads
    return res

"""ORACLE: ctypes front of c/refpath.c (OpenMP loop nests of the reference's kernels; CPU baseline)."""
import ctypes as C
import time

import numpy as np

from .femdict import lib as _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def set_threads(n):
    _lib().ora_set_threads(int(n))


def get_threads():
    return int(_lib().ora_get_threads())


def _i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def var_basic(itp, slot, shift, cp, el, host, x):
    nsel, (_, _, na, nq) = len(el), itp.shape
    out = np.empty((nsel, nq))
    cp = np.ascontiguousarray(cp, dtype=np.int32)
    _lib().ora_var_basic(nq, na, C.c_int64(nsel), _p(itp), int(slot), C.c_int64(int(shift)), _p(cp), C.c_int64(cp.shape[1]),
                         _p(_i64(el)), _p(_i64(host)), _p(x), _p(out))
    return out


def kval_basic(itp, dslot, bslot, vals, sid, shift, el, host, K):
    nsel, (_, _, na, nq) = len(el), itp.shape
    vals = np.ascontiguousarray(vals, dtype=np.float64)
    _lib().ora_kval_basic(nq, na, C.c_int64(nsel), _p(itp), int(dslot), int(bslot), _p(vals), _p(sid),
                          C.c_int64(sid.shape[2]), C.c_int64(int(shift)), _p(_i64(el)), _p(_i64(host)), _p(K))


def res_basic(itp, slot, vals, shift, cp, el, host, residue):
    nsel, (_, _, na, nq) = len(el), itp.shape
    vals = np.ascontiguousarray(vals, dtype=np.float64)
    cp = np.ascontiguousarray(cp, dtype=np.int32)
    _lib().ora_res_basic(nq, na, C.c_int64(nsel), _p(itp), int(slot), _p(vals), C.c_int64(int(shift)), _p(cp),
                         C.c_int64(cp.shape[1]), _p(_i64(el)), _p(_i64(host)), _p(residue))


class CsrOperator:
    """A @ x through the OpenMP CSR kernel (1-based arrays of GlobalField)."""

    calls, seconds = 0, 0.0      # class-wide counters (bench.py's cpu_baseline sub-metrics)

    def __init__(self, ptr, col, val, n):
        self.ptr = np.ascontiguousarray(ptr, dtype=np.int32)
        self.col = np.ascontiguousarray(col, dtype=np.int32)
        self.data = np.ascontiguousarray(val, dtype=np.float64)
        self.shape = (n, n)

    def __matmul__(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.empty(self.shape[0])
        t0 = time.perf_counter()
        _lib().ora_spmv_csr(C.c_int64(self.shape[0]), _p(self.ptr), _p(self.col), _p(self.data), _p(x), _p(y))
        CsrOperator.seconds += time.perf_counter() - t0
        CsrOperator.calls += 1
        return y

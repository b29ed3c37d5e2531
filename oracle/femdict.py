"""ORACLE: FEM_Dict (reference src/misc/06_GPU_Dict.jl), sequential-insertion restatement.

Slot IDs are 1-based like the reference. Heavy loops live in c/femdict.c.
"""
import ctypes
import numpy as np

from . import build as _build

_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(_build.build())
        _lib.ora_wang64.restype = ctypes.c_uint64
        _lib.ora_wang64.argtypes = [ctypes.c_uint64]
        _lib.ora_dict_size.restype = ctypes.c_int64
        _lib.ora_dict_size.argtypes = [ctypes.c_int64]
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def dict_size(n: int) -> int:
    """_DictSize, 06_GPU_Dict.jl:13."""
    return int(lib().ora_dict_size(int(n)))


# key packing helpers, 06_GPU_Dict.jl:227-238
def I32I32_To_UI64(x, y):
    x = np.asarray(x).astype(np.int64).astype(np.uint64) & np.uint64(0xFFFFFFFF)
    y = np.asarray(y).astype(np.int64).astype(np.uint64) & np.uint64(0xFFFFFFFF)
    return (x << np.uint64(32)) + y


def UI64_To_UpperHalf(k):
    return ((np.asarray(k, dtype=np.uint64) >> np.uint64(32)) & np.uint64(0xFFFFFFFF)).astype(np.int64).astype(np.int32)


def UI64_To_LowerHalf(k):
    return (np.asarray(k, dtype=np.uint64) & np.uint64(0xFFFFFFFF)).astype(np.int64).astype(np.int32)


def I4I30I30_To_UI64(x, y, z):
    m30 = np.uint64(0x3FFFFFFF)
    x = np.asarray(x).astype(np.int64)
    y = np.asarray(y).astype(np.int64).astype(np.uint64) & m30
    z = np.asarray(z).astype(np.int64).astype(np.uint64) & m30
    return ((x + 1).astype(np.uint64) << np.uint64(60)) + (y << np.uint64(30)) + z


class FemDict:
    """dumb_FEM_Dict_Init(ArrayType, Int32): keys/hashs/hash_init/hash_prev/hash_next + vals."""

    def __init__(self, size_hint: int = 16):
        self._alloc(dict_size(size_hint))

    def _alloc(self, n):
        self.keys = np.zeros(n, np.uint64)
        self.hashs = np.zeros(n, np.uint64)
        self.hash_init = np.zeros(n, np.int32)
        self.hash_prev = np.zeros(n, np.int32)
        self.hash_next = np.zeros(n, np.int32)
        self.vals = np.zeros(n, np.int32)

    def total_ids(self):
        """get_Total_IDs: occupied slots, ascending, 1-based (06_GPU_Dict.jl:21)."""
        return (np.nonzero(self.keys != 0)[0] + 1).astype(np.int32)

    def _insert(self, new_keys):
        new_keys = np.ascontiguousarray(new_keys, dtype=np.uint64)
        ids = np.zeros(len(new_keys), np.int32)
        lib().ora_dict_set(ctypes.c_int64(len(self.keys)), _p(self.keys), _p(self.hashs), _p(self.hash_init),
                           _p(self.hash_prev), _p(self.hash_next), ctypes.c_int64(len(new_keys)),
                           _p(new_keys), _p(ids))
        return ids

    def set_ids(self, new_keys):
        """FEM_Dict_SetID! (06_GPU_Dict.jl:45-93)."""
        new_keys = np.ascontiguousarray(new_keys, dtype=np.uint64)
        if len(new_keys) == 0:
            return np.zeros(0, np.int32)
        src = self.total_ids()
        est = dict_size(len(new_keys) + len(src))
        if est != len(self.keys):
            src_keys = self.keys[src - 1].copy()
            src_vals = self.vals[src - 1].copy()
            self._alloc(est)
            if len(src):
                mapped = self._insert(src_keys)
                self.vals[mapped - 1] = src_vals
        return self._insert(new_keys)

    def get_ids(self, target_keys):
        """FEM_Dict_GetID (06_GPU_Dict.jl:95-100, kernel :164-188)."""
        target_keys = np.ascontiguousarray(target_keys, dtype=np.uint64)
        ids = np.zeros(len(target_keys), np.int32)
        lib().ora_dict_get(ctypes.c_int64(len(self.keys)), _p(self.keys), _p(self.hashs), _p(self.hash_init),
                           _p(self.hash_next), ctypes.c_int64(len(target_keys)), _p(target_keys), _p(ids))
        return ids

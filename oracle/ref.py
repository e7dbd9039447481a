"""Window into the UNMODIFIED reference build (oracle/_ref/libjincresize_ref.so, see Makefile target `ref` and
oracle/ref_shim.cpp): where the library is, and a reader for the tables the reference built for a filter instance.
TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

REF_PLUGIN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "libjincresize_ref.so")


def available() -> bool:
    return os.path.exists(REF_PLUGIN)


class RefTables:
    """Reads the reference's own tables out of a filter built by oracle/_ref (via oracle/ref_shim.cpp)."""

    def __init__(self, ref_lib_path: str = REF_PLUGIN):
        self.lib = C.CDLL(ref_lib_path)
        self.lib.ref_table_count.restype = C.c_int
        self.lib.ref_table_count.argtypes = [C.c_void_p]
        self.lib.ref_table_view.restype = C.c_int
        self.lib.ref_table_view.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                            C.POINTER(C.POINTER(C.c_int)), C.POINTER(C.POINTER(C.c_float))]
        self.lib.ref_lut.restype = C.POINTER(C.c_double)
        self.lib.ref_lut.argtypes = [C.c_void_p]

    def count(self, clip) -> int:
        return self.lib.ref_table_count(clip.filter_info)

    def lut(self, clip) -> np.ndarray:
        p = self.lib.ref_lut(clip.filter_info)
        return np.ctypeslib.as_array(p, shape=(1024,)).copy()

    def table(self, clip, k: int, dst_w: int, dst_h: int):
        """Returns (filter_size, coeff_stride, meta[h,w,3] int32, factor float32 flat)."""
        fs, cs = C.c_int(), C.c_int()
        meta, factor = C.POINTER(C.c_int)(), C.POINTER(C.c_float)()
        rc = self.lib.ref_table_view(clip.filter_info, k, C.byref(fs), C.byref(cs), C.byref(meta), C.byref(factor))
        if rc != 0:
            raise IndexError(k)
        m = np.ctypeslib.as_array(meta, shape=(dst_h, dst_w, 3)).copy()
        nfl = int(m[..., 2].max()) + fs.value * cs.value
        f = np.ctypeslib.as_array(factor, shape=(nfl,)).copy()
        return fs.value, cs.value, m, f

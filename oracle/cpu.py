"""ctypes binding of oracle/libjinc_oracle.so (the plain-C restatement in jinc_oracle.c).

TEST INFRASTRUCTURE: imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libjinc_oracle.so")


class TableParams(C.Structure):
    _fields_ = [("quant_x", C.c_int), ("quant_y", C.c_int), ("src_w", C.c_int), ("src_h", C.c_int),
                ("dst_w", C.c_int), ("dst_h", C.c_int), ("radius", C.c_double), ("crop_left", C.c_double),
                ("crop_top", C.c_double), ("crop_w", C.c_double), ("crop_h", C.c_double)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class _CTable(C.Structure):
    _fields_ = [("filter_size", C.c_int), ("coeff_stride", C.c_int), ("dst_w", C.c_int), ("dst_h", C.c_int),
                ("meta", C.POINTER(C.c_int32)), ("factor", C.POINTER(C.c_float)), ("factor_len", C.c_size_t),
                ("border", C.POINTER(C.c_uint8)), ("phase", C.POINTER(C.c_int32)), ("n_blocks", C.c_size_t)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} missing: run `make host`")
        L = C.CDLL(LIB_PATH)
        L.jo_jinc_sqr.restype = C.c_double
        L.jo_jinc_sqr.argtypes = [C.c_double]
        L.jo_radius_for_tap.restype = C.c_double
        L.jo_radius_for_tap.argtypes = [C.c_int]
        L.jo_lut_init.restype = None
        L.jo_lut_init.argtypes = [C.POINTER(C.c_double), C.c_double, C.c_double]
        L.jo_table_generate.restype = C.c_int
        L.jo_table_generate.argtypes = [C.POINTER(TableParams), C.POINTER(C.c_double), C.POINTER(_CTable)]
        L.jo_table_free.restype = None
        L.jo_table_free.argtypes = [C.POINTER(_CTable)]
        L.jo_plane_params.restype = C.c_int
        L.jo_plane_params.argtypes = [C.c_int] * 4 + [C.c_double] * 4 + [C.c_int] * 6 + [C.POINTER(TableParams)]
        L.jo_resize_rows.restype = None
        L.jo_resize_rows.argtypes = [C.POINTER(_CTable), C.c_int, C.c_void_p, C.c_ssize_t, C.c_void_p, C.c_ssize_t,
                                     C.c_float, C.c_int, C.c_int]
        _lib = L
    return _lib


def radius_for_tap(tap: int) -> float:
    return lib().jo_radius_for_tap(tap)


def make_lut(tap: int, blur: float = 1.0) -> np.ndarray:
    """1024 doubles; `blur` is what the script passes (float32-rounded by AviSynth; 0 means 1.0)."""
    lut = np.zeros(1024, dtype=np.float64)
    lib().jo_lut_init(lut.ctypes.data_as(C.POINTER(C.c_double)), radius_for_tap(tap), float(np.float32(blur)))
    return lut


CPLACE = {"mpeg2": 0, "mpeg1": 1, "topleft": 2}


def plane_params(src_w, src_h, target_w, target_h, *, src_left=0.0, src_top=0.0, src_width=None, src_height=None,
                 quant_x=256, quant_y=256, tap=3, sub_w=0, sub_h=0, cplace="mpeg2"):
    """Per-table geometry exactly as Create_JincResize derives it; script floats are rounded to float32 first."""
    f32 = lambda v: float(np.float32(v))
    out = (TableParams * 2)()
    n = lib().jo_plane_params(src_w, src_h, target_w, target_h, f32(src_left), f32(src_top),
                              f32(src_width) if src_width is not None else float(src_w),
                              f32(src_height) if src_height is not None else float(src_h),
                              quant_x, quant_y, tap, sub_w, sub_h, CPLACE[cplace.lower()], out)
    res = []
    for i in range(n):
        p = TableParams()
        C.memmove(C.byref(p), C.byref(out[i]), C.sizeof(TableParams))
        res.append(p)
    return res


class Table:
    """Owns one jo_table; numpy views are copies taken at construction."""

    def __init__(self, params: TableParams, lut: np.ndarray):
        self._c = _CTable()
        self.params = params
        lut = np.ascontiguousarray(lut, dtype=np.float64)
        rc = lib().jo_table_generate(C.byref(params), lut.ctypes.data_as(C.POINTER(C.c_double)), C.byref(self._c))
        if rc != 0:
            raise MemoryError("jo_table_generate")
        c = self._c
        self.filter_size, self.coeff_stride = c.filter_size, c.coeff_stride
        self.dst_w, self.dst_h = c.dst_w, c.dst_h
        self.n_blocks = c.n_blocks
        n = c.dst_w * c.dst_h
        self.meta = np.ctypeslib.as_array(c.meta, shape=(c.dst_h, c.dst_w, 3))
        self.border = np.ctypeslib.as_array(c.border, shape=(c.dst_h, c.dst_w))
        self.phase = np.ctypeslib.as_array(c.phase, shape=(c.dst_h, c.dst_w, 2))
        self.factor = np.ctypeslib.as_array(c.factor, shape=(max(int(c.factor_len), 1),))[: c.factor_len]

    def block(self, y: int, x: int) -> np.ndarray:
        """filter_size x filter_size weights of output pixel (y, x)."""
        fs, cs = self.filter_size, self.coeff_stride
        o = int(self.meta[y, x, 2])
        return self.factor[o:o + fs * cs].reshape(fs, cs)[:, :fs]

    def resize(self, src: np.ndarray, peak: float = 0.0, rows=None) -> np.ndarray:
        src = np.ascontiguousarray(src)
        dst = np.zeros((self.dst_h, self.dst_w), dtype=src.dtype)
        y0, y1 = rows if rows is not None else (0, self.dst_h)
        lib().jo_resize_rows(C.byref(self._c), src.dtype.itemsize, src.ctypes.data, src.strides[0] // src.dtype.itemsize,
                             dst.ctypes.data, dst.strides[0] // dst.dtype.itemsize, float(peak), y0, y1)
        return dst

    def close(self):
        if self._c.meta:
            self.meta = self.border = self.phase = self.factor = None
            lib().jo_table_free(C.byref(self._c))

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

"""ctypes driver for the in-process AviSynth+ C-API stand-in (minihost/minihost.cpp).

Used by tests and bench.py to drive C plugins -- the unmodified reference build
(oracle/_ref/libjincresize_ref.so) and the B200 plugin -- through the same
AviSynth call sequence: load plugin, build a source clip, invoke a script
function with positional/named arguments, pull frames, read planes and frame
properties.  Test/bench infrastructure; not on the product path.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

MINIHOST_LIB = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libavs_minihost.so")


# ---- pixel types (mirror minihost/include/avisynth_c.h) -------------------
CS_YUVA = 1 << 27
CS_BGR = 1 << 28
CS_YUV = 1 << 29
CS_INTERLEAVED = 1 << 30
CS_PLANAR = -(1 << 31)  # sign bit of a 32-bit int
CS_VPLANEFIRST = 1 << 3
SUB_W = {1: 3, 2: 0, 4: 1}
SUB_H = {1: 3 << 8, 2: 0 << 8, 4: 1 << 8}
BITS = {8: 0 << 16, 16: 1 << 16, 32: 2 << 16, 10: 5 << 16, 12: 6 << 16, 14: 7 << 16}

PLANAR_Y, PLANAR_U, PLANAR_V, PLANAR_A, PLANAR_R, PLANAR_G, PLANAR_B = 1, 2, 4, 16, 32, 64, 128


def _i32(v: int) -> int:
    v &= 0xFFFFFFFF
    return v - (1 << 32) if v & 0x80000000 else v


def pixel_type(family: str, bits: int) -> int:
    """family: 'y', '420', '422', '444', '411', 'rgbp', 'rgbap', 'yuva420', 'yuva422', 'yuva444', 'bgr32'."""
    b = BITS[bits]
    if family == "y":
        return _i32(CS_PLANAR | CS_INTERLEAVED | CS_YUV | b)
    if family in ("420", "422", "444", "411"):
        sw, sh = {"420": (2, 2), "422": (2, 1), "444": (1, 1), "411": (4, 1)}[family]
        return _i32(CS_PLANAR | CS_YUV | CS_VPLANEFIRST | SUB_W[sw] | SUB_H[sh] | b)
    if family in ("yuva420", "yuva422", "yuva444"):
        sw, sh = {"yuva420": (2, 2), "yuva422": (2, 1), "yuva444": (1, 1)}[family]
        return _i32(CS_PLANAR | CS_YUVA | CS_VPLANEFIRST | SUB_W[sw] | SUB_H[sh] | b)
    if family == "rgbp":
        return _i32(CS_PLANAR | CS_BGR | 1 | b)
    if family == "rgbap":
        return _i32(CS_PLANAR | CS_BGR | 2 | b)
    if family == "bgr32":
        return _i32(2 | CS_BGR | CS_INTERLEAVED)
    raise ValueError(family)


@dataclass(frozen=True)
class Format:
    """Planar clip format: plane ids in the reference's processing order
    (src/JincResize.cpp:539-540: Y,U,V,A or G,B,R,A)."""

    family: str
    bits: int

    @property
    def pixel_type(self) -> int:
        return pixel_type(self.family, self.bits)

    @property
    def dtype(self):
        return np.uint8 if self.bits == 8 else (np.float32 if self.bits == 32 else np.uint16)

    @property
    def planes(self):
        if self.family == "y":
            return [PLANAR_Y]
        if self.family in ("rgbp",):
            return [PLANAR_G, PLANAR_B, PLANAR_R]
        if self.family in ("rgbap",):
            return [PLANAR_G, PLANAR_B, PLANAR_R, PLANAR_A]
        if self.family.startswith("yuva"):
            return [PLANAR_Y, PLANAR_U, PLANAR_V, PLANAR_A]
        return [PLANAR_Y, PLANAR_U, PLANAR_V]

    @property
    def subsampling(self):
        """(log2 w, log2 h) of the chroma planes."""
        return {"420": (1, 1), "422": (1, 0), "411": (2, 0), "yuva420": (1, 1), "yuva422": (1, 0)}.get(self.family, (0, 0))

    def plane_shape(self, i: int, width: int, height: int):
        if i in (1, 2) and self.family not in ("rgbp", "rgbap"):
            sw, sh = self.subsampling
            return height >> sh, width >> sw
        return height, width

    @property
    def peak(self) -> int:
        return (1 << self.bits) - 1


YV12 = Format("420", 8)
YUV420P8 = YV12
YUV420P10 = Format("420", 10)
YUV444P16 = Format("444", 16)
RGBPS = Format("rgbp", 32)


class _Lib:
    _inst = None

    @classmethod
    def get(cls):
        if cls._inst is None:
            path = MINIHOST_LIB
            if not os.path.exists(path):
                raise RuntimeError(f"{path} is missing: run `make host` (or __graft_entry__.build())")
            lib = C.CDLL(path, mode=C.RTLD_GLOBAL)  # plugins resolve avs_* against it
            vp, ci, cp = C.c_void_p, C.c_int, C.c_char_p
            sig = {
                "mh_env_create": (vp, []),
                "mh_env_destroy": (None, [vp]),
                "mh_env_set_interface": (None, [vp, ci, ci]),
                "mh_env_set_cpu_flags": (None, [vp, ci]),
                "mh_last_error": (cp, [vp]),
                "mh_load_plugin": (cp, [vp, cp]),
                "mh_function_params": (cp, [vp, cp]),
                "mh_source_create": (vp, [vp, ci, ci, ci, ci, ci]),
                "mh_source_fill_plane": (ci, [vp, ci, ci, vp, C.c_ssize_t]),
                "mh_source_set_prop_int": (None, [vp, cp, C.c_int64]),
                "mh_source_clear_prop": (None, [vp, cp]),
                "mh_args_create": (vp, []),
                "mh_args_destroy": (None, [vp]),
                "mh_args_add_clip": (None, [vp, vp, cp]),
                "mh_args_add_int": (None, [vp, ci, cp]),
                "mh_args_add_float": (None, [vp, C.c_float, cp]),
                "mh_args_add_string": (None, [vp, cp, cp]),
                "mh_invoke_clip": (vp, [vp, cp, vp]),
                "mh_clip_filter_info": (vp, [vp]),
                "mh_clip_mt_mode": (ci, [vp]),
                "mh_frame_prop_int": (ci, [vp, vp, cp, C.POINTER(C.c_int64)]),
                "mh_live_frames": (C.c_long, []),
                "mh_live_clips": (C.c_long, []),
                "mh_pull_frames": (C.c_double, [vp, ci, ci, ci]),
                "avs_release_clip": (None, [vp]),
                "avs_get_frame": (vp, [vp, ci]),
                "avs_release_video_frame": (None, [vp]),
                "avs_get_video_info": (vp, [vp]),
                "avs_clip_get_error": (cp, [vp]),
                "avs_get_pitch_p": (ci, [vp, ci]),
                "avs_get_row_size_p": (ci, [vp, ci]),
                "avs_get_height_p": (ci, [vp, ci]),
                "avs_get_read_ptr_p": (vp, [vp, ci]),
            }
            for name, (res, args) in sig.items():
                fn = getattr(lib, name)
                fn.restype, fn.argtypes = res, args
            cls._inst = lib
        return cls._inst


class AvsError(RuntimeError):
    """An AVS_Value error returned by a script function (the text is the plugin's)."""


class Clip:
    def __init__(self, env: "Env", handle: int, fmt: Format | None):
        self.env, self.handle, self.fmt = env, handle, fmt

    def release(self):
        if self.handle:
            _Lib.get().avs_release_clip(self.handle)
            self.handle = None

    def __del__(self):  # best effort
        try:
            self.release()
        except Exception:
            pass

    @property
    def filter_info(self) -> int:
        return _Lib.get().mh_clip_filter_info(self.handle)

    @property
    def mt_mode(self) -> int:
        return _Lib.get().mh_clip_mt_mode(self.handle)

    @property
    def size(self):
        vi = C.cast(_Lib.get().avs_get_video_info(self.handle), C.POINTER(C.c_int))
        return vi[0], vi[1]  # width, height

    def get_frame(self, n: int):
        """Returns (planes, props) where planes is a list of numpy arrays in processing order."""
        lib = _Lib.get()
        f = lib.avs_get_frame(self.handle, n)
        if not f:
            raise AvsError((lib.avs_clip_get_error(self.handle) or b"avs_get_frame returned NULL").decode())
        try:
            err = lib.avs_clip_get_error(self.handle)
            if err:
                raise AvsError(err.decode())
            planes = []
            for p in self.fmt.planes:
                pitch, rs, h = lib.avs_get_pitch_p(f, p), lib.avs_get_row_size_p(f, p), lib.avs_get_height_p(f, p)
                ptr = lib.avs_get_read_ptr_p(f, p)
                raw = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(h, pitch))
                planes.append(raw[:, :rs].copy().view(self.fmt.dtype))
            props = {}
            v = C.c_int64()
            for key in ("_ChromaLocation", "_Matrix", "_Primaries", "_Transfer", "_ColorRange", "_FieldBased"):
                if lib.mh_frame_prop_int(self.env.handle, f, key.encode(), C.byref(v)):
                    props[key] = v.value
            return planes, props
        finally:
            lib.avs_release_video_frame(f)

    def pull(self, first: int, count: int, threads: int = 1) -> float:
        """Frame-parallel pull of `count` frames with `threads` host threads; returns seconds."""
        t = _Lib.get().mh_pull_frames(self.handle, first, count, threads)
        if t < 0:
            raise AvsError("get_frame failed while pulling frames")
        return t


class Env:
    def __init__(self):
        self.lib = _Lib.get()
        self.handle = self.lib.mh_env_create()

    def set_interface(self, version: int, bugfix: int = 0):
        self.lib.mh_env_set_interface(self.handle, version, bugfix)

    def set_cpu_flags(self, flags: int):
        self.lib.mh_env_set_cpu_flags(self.handle, flags)

    def load_plugin(self, path: str) -> str:
        r = self.lib.mh_load_plugin(self.handle, path.encode())
        if r is None:
            raise RuntimeError(self.lib.mh_last_error(self.handle).decode())
        return r.decode()

    def function_params(self, name: str):
        r = self.lib.mh_function_params(self.handle, name.encode())
        return r.decode() if r is not None else None

    def source(self, fmt: Format, width: int, height: int, frames, num_frames: int | None = None, props=None) -> Clip:
        """frames: list (one entry per stored frame) of lists of 2-D numpy plane arrays."""
        n_distinct = len(frames)
        h = self.lib.mh_source_create(self.handle, width, height, fmt.pixel_type, num_frames or n_distinct, n_distinct)
        if not h:
            raise RuntimeError(self.lib.mh_last_error(self.handle).decode())
        clip = Clip(self, h, fmt)
        for k, planes in enumerate(frames):
            assert len(planes) == len(fmt.planes)
            for i, (pid, arr) in enumerate(zip(fmt.planes, planes)):
                arr = np.ascontiguousarray(arr, dtype=fmt.dtype)
                assert arr.shape == fmt.plane_shape(i, width, height), (arr.shape, fmt.plane_shape(i, width, height))
                rc = self.lib.mh_source_fill_plane(h, k, pid, arr.ctypes.data, arr.strides[0])
                assert rc == 0
        for key, val in (props or {}).items():
            self.lib.mh_source_set_prop_int(h, key.encode(), int(val))
        return clip

    def invoke(self, name: str, clip: Clip, *positional, **named) -> Clip:
        """Call a script function: clip first, then positional ints, then named arguments.
        Python float -> script float (32-bit), int -> int, str -> string."""
        a = self.lib.mh_args_create()
        try:
            self.lib.mh_args_add_clip(a, clip.handle, None)
            for v in positional:
                self._add(a, v, None)
            for k, v in named.items():
                if v is not None:
                    self._add(a, v, k.encode())
            h = self.lib.mh_invoke_clip(self.handle, name.encode(), a)
            if not h:
                raise AvsError(self.lib.mh_last_error(self.handle).decode())
            return Clip(self, h, clip.fmt)
        finally:
            self.lib.mh_args_destroy(a)

    def _add(self, a, v, name):
        if isinstance(v, bool):
            raise TypeError("bool arguments are not used by these functions")
        if isinstance(v, (int, np.integer)):
            self.lib.mh_args_add_int(a, int(v), name)
        elif isinstance(v, (float, np.floating)):
            self.lib.mh_args_add_float(a, float(v), name)
        elif isinstance(v, str):
            self.lib.mh_args_add_string(a, v.encode(), name)
        else:
            raise TypeError(type(v))

    @staticmethod
    def live_objects():
        lib = _Lib.get()
        return lib.mh_live_frames(), lib.mh_live_clips()

"""oracle/c_oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes binding of ``oracle/libt360oracle.so`` (the plain-C restatement in
``t360_oracle.c``).  Only tests/, ``__graft_entry__.smoke()`` and bench.py's
cpu_baseline / reference legs may import this.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
ORACLE_SO = _HERE / "libt360oracle.so"


class Segment(C.Structure):
    _fields_ = [("left", C.c_int), ("top", C.c_int), ("width", C.c_int), ("height", C.c_int),
                ("nkx", C.c_int), ("nky", C.c_int), ("kx_off", C.c_int), ("ky_off", C.c_int)]


_lib = None


def build(force: bool = False) -> None:
    if force or not ORACLE_SO.exists() or ORACLE_SO.stat().st_mtime < (_HERE / "t360_oracle.c").stat().st_mtime:
        subprocess.run(["make", "-C", str(_HERE), "oracle"], check=True, capture_output=True)


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(str(ORACLE_SO))
        vp, ci, sz = C.c_void_p, C.c_int, C.c_size_t
        L.t360o_scaled_dims.argtypes = [vp, ci, ci, C.POINTER(ci), C.POINTER(ci)]
        L.t360o_generate_map.argtypes = [vp, ci, ci, ci, ci, vp]
        L.t360o_transform_pos.argtypes = [vp, C.c_float, C.c_float, C.c_float, C.POINTER(C.c_float), C.POINTER(C.c_float)]
        L.t360o_filter_plan.argtypes = [vp, ci, ci, ci, ci, vp, ci, vp, ci, C.POINTER(ci)]
        L.t360o_remap_ksize.argtypes = [ci]
        L.t360o_build_itab.argtypes = [ci, vp]
        L.t360o_build_itab.restype = None
        L.t360o_remap_u8.argtypes = [vp, ci, ci, sz, vp, ci, ci, sz, vp, ci, ci]
        L.t360o_remap_u8.restype = None
        L.t360o_sepfilter_roi_u8.argtypes = [vp, ci, ci, sz, ci, ci, ci, ci, vp, sz, vp, ci, vp, ci]
        L.t360o_sepfilter_roi_u8.restype = None
        L.t360o_filter_plane.argtypes = [vp, vp, ci, ci, sz, vp, sz, vp, ci, vp]
        L.t360o_filter_plane.restype = None
        L.t360o_transform_plane.argtypes = [vp, vp, ci, ci, sz, vp, ci, ci, sz, vp, ci, ci, ci, vp, ci, vp]
        L.t360o_resize_area_u8.argtypes = [vp, ci, ci, sz, vp, ci, ci, sz]
        L.t360o_noise_plane.argtypes = [vp, ci, ci, sz, C.c_uint32, C.c_uint32]
        L.t360o_noise_plane.restype = None
        L.t360o_fnv1a64.argtypes = [vp, sz]
        L.t360o_fnv1a64.restype = C.c_uint64
        _lib = L
    return _lib


def _ctxp(ctx):
    assert C.sizeof(ctx) == 112
    return C.addressof(ctx)


def scaled_dims(ctx, out_w, out_h):
    sw, sh = C.c_int(), C.c_int()
    lib().t360o_scaled_dims(_ctxp(ctx), out_w, out_h, C.byref(sw), C.byref(sh))
    return sw.value, sh.value


def generate_map(ctx, in_w, in_h, out_w, out_h) -> np.ndarray:
    sw, sh = scaled_dims(ctx, out_w, out_h)
    m = np.empty((sh, sw, 2), np.float32)
    if not lib().t360o_generate_map(_ctxp(ctx), in_w, in_h, out_w, out_h, m.ctypes.data):
        raise RuntimeError("t360o_generate_map failed")
    return m


def filter_plan(ctx, in_w, in_h, scaled_out_w, scaled_out_h, max_segs=1 << 18, max_taps=1 << 24):
    """Returns (segments ctypes array, nsegs, taps float32 array)."""
    segs = (Segment * max_segs)()
    taps = np.zeros(max_taps, np.float32)
    nt = C.c_int()
    n = lib().t360o_filter_plan(_ctxp(ctx), in_w, in_h, scaled_out_w, scaled_out_h, segs, max_segs,
                                taps.ctypes.data, max_taps, C.byref(nt))
    if n < 0:
        raise RuntimeError("t360o_filter_plan: buffers too small")
    return segs, n, taps[:nt.value].copy()


def plan_as_list(segs, n, taps):
    return [(s.left, s.top, s.width, s.height, taps[s.kx_off:s.kx_off + s.nkx], taps[s.ky_off:s.ky_off + s.nky])
            for s in segs[:n]]


def build_itab(interp) -> np.ndarray:
    k = lib().t360o_remap_ksize(interp)
    t = np.zeros((1024, k, k), np.int16)
    lib().t360o_build_itab(interp, t.ctypes.data)
    return t


def remap_u8(src: np.ndarray, mapxy: np.ndarray, interp: int, border: int = 3, dst: np.ndarray | None = None):
    assert src.dtype == np.uint8 and src.strides[1] == 1
    mapxy = np.ascontiguousarray(mapxy, np.float32)
    dh, dw = mapxy.shape[:2]
    if dst is None:
        dst = np.zeros((dh, dw), np.uint8)
    lib().t360o_remap_u8(src.ctypes.data, src.shape[1], src.shape[0], src.strides[0], dst.ctypes.data, dw, dh,
                         dst.strides[0], mapxy.ctypes.data, interp, border)
    return dst


def sepfilter_roi(parent: np.ndarray, rx, ry, rw, rh, kx, ky, dst: np.ndarray):
    kx = np.ascontiguousarray(kx, np.float32)
    ky = np.ascontiguousarray(ky, np.float32)
    lib().t360o_sepfilter_roi_u8(parent.ctypes.data, parent.shape[1], parent.shape[0], parent.strides[0], rx, ry, rw,
                                 rh, dst.ctypes.data, dst.strides[0], kx.ctypes.data, kx.size, ky.ctypes.data, ky.size)
    return dst


def filter_plane(ctx, src: np.ndarray, segs, n, taps) -> np.ndarray:
    dst = np.zeros_like(src)
    lib().t360o_filter_plane(_ctxp(ctx), src.ctypes.data, src.shape[1], src.shape[0], src.strides[0], dst.ctypes.data,
                             dst.strides[0], segs, n, taps.ctypes.data)
    return dst


class OraclePlan:
    """Everything generateMapForPlane (cpp:504-576) caches for one plan index, from the C restatement."""

    def __init__(self, ctx, in_w, in_h, out_w, out_h):
        self.ctx, self.in_w, self.in_h = ctx, in_w, in_h
        self.map = generate_map(ctx, in_w, in_h, out_w, out_h)
        self.segs, self.nsegs, self.taps = None, 0, np.zeros(1, np.float32)
        if ctx.enable_low_pass_filter:
            sh, sw = self.map.shape[:2]
            self.segs, self.nsegs, self.taps = filter_plan(ctx, in_w, in_h, sw, sh)


def transform_plane(ctx, plan: OraclePlan, src: np.ndarray, out_w, out_h, map_index=0, prefill=0) -> np.ndarray:
    """The reference's transformPlane (cpp:707-794), non-resize branch, from the C restatement."""
    assert src.dtype == np.uint8 and src.strides[1] == 1
    dst = np.full((out_h, out_w), prefill, np.uint8)
    mh, mw = plan.map.shape[:2]
    ok = lib().t360o_transform_plane(_ctxp(ctx), src.ctypes.data, src.shape[1], src.shape[0], src.strides[0],
                                     dst.ctypes.data, out_w, out_h, dst.strides[0], plan.map.ctypes.data, mw, mh,
                                     map_index, plan.segs, plan.nsegs, plan.taps.ctypes.data)
    if not ok:
        raise RuntimeError("oracle: t360o_transform_plane failed")
    return dst


def resize_area(src: np.ndarray, dw: int, dh: int) -> np.ndarray:
    dst = np.zeros((dh, dw), np.uint8)
    if not lib().t360o_resize_area_u8(src.ctypes.data, src.shape[1], src.shape[0], src.strides[0], dst.ctypes.data, dw, dh,
                                      dst.strides[0]):
        raise ValueError("oracle: INTER_AREA enlarging is not restated")
    return dst


def noise_plane(w, h, plane=0, frame=0, pitch=None) -> np.ndarray:
    pitch = pitch or w
    a = np.zeros((h, pitch), np.uint8)
    lib().t360o_noise_plane(a.ctypes.data, w, h, pitch, plane, frame)
    return a[:, :w] if pitch == w else a


def fnv1a64(a) -> str:
    b = np.ascontiguousarray(a).view(np.uint8).ravel()
    return "%016x" % lib().t360o_fnv1a64(b.ctypes.data, b.size)

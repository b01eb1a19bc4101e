"""oracle/ff_harness.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Drives the reference's UNMODIFIED ffmpeg filter source (``Transform360/vf_transform360.c``), compiled against the
libavfilter stand-in ``oracle/ffshim`` and linked either with the reference library (``variant="ref"``: oracle/_ref/
libt360ref.so + cv2) or with the product (``variant="b200"``: transform360_b200/lib/libTransform360.so).  Arguments use
the filter's own option names (``"cube_edge_length=256:interpolation_alg=cubic:..."``), defaults come from its
AVOption table, the output size from its ``config_output``.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIBS = {"ref": _HERE / "_ref" / "libvf_t360_ref.so", "b200": _HERE / "_ref" / "libvf_t360_b200.so",
         "cuda": _HERE / "_ref" / "libvf_t360_cuda.so"}  # "cuda": the product's transform360_cuda filter (device frames)
_loaded = {}


def available(variant: str) -> bool:
    return _LIBS[variant].exists()


def _lib(variant: str):
    if variant not in _loaded:
        if variant == "ref":
            from . import ref_harness
            ref_harness.ref_lib()  # installs the cv2 hooks in libt360ref.so (the loader then reuses that object)
        L = C.CDLL(str(_LIBS[variant]), mode=os.RTLD_LOCAL)
        L.t360f_open.restype = C.c_void_p
        L.t360f_open.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int)]
        L.t360f_close.argtypes = [C.c_void_p]
        L.t360f_close.restype = None
        L.t360f_out_size.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.t360f_out_size.restype = None
        L.t360f_filter.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.t360f_filter.restype = C.c_int
        L.t360f_open_cuda.restype = C.c_void_p
        L.t360f_open_cuda.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]
        L.t360f_filter_cuda.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.t360f_filter_cuda.restype = C.c_int
        _loaded[variant] = L
    return _loaded[variant]


class Filter:
    """One instance of the ``transform360`` filter on a yuv420p link of in_w x in_h."""

    def __init__(self, variant: str, args: str, in_w: int, in_h: int):
        self.L = _lib(variant)
        err = C.c_int()
        self.h = self.L.t360f_open(args.encode(), in_w, in_h, 0, C.byref(err))
        if not self.h:
            raise ValueError(f"filter rejected {args!r}: AVERROR {err.value}")
        self.in_w, self.in_h = in_w, in_h
        w, h = C.c_int(), C.c_int()
        self.L.t360f_out_size(self.h, C.byref(w), C.byref(h))
        self.out_w, self.out_h = w.value, h.value

    def close(self):
        if getattr(self, "h", None):
            self.L.t360f_close(self.h)
            self.h = None

    __del__ = close

    def filter(self, planes):
        """planes: [Y, U, V] uint8 arrays; returns the output [Y, U, V]."""
        outs = [np.zeros((self.out_h, self.out_w), np.uint8),
                np.zeros(((self.out_h + 1) // 2, (self.out_w + 1) // 2), np.uint8),
                np.zeros(((self.out_h + 1) // 2, (self.out_w + 1) // 2), np.uint8)]
        VP, IA = C.c_void_p * 3, C.c_int * 3
        rc = self.L.t360f_filter(self.h, VP(*[p.ctypes.data for p in planes]), IA(*[p.strides[0] for p in planes]),
                                 VP(*[o.ctypes.data for o in outs]), IA(*[o.strides[0] for o in outs]))
        if rc:
            raise RuntimeError(f"filter_frame failed: AVERROR {rc}")
        return outs


AV_PIX_FMT_YUV420P, AV_PIX_FMT_GRAY8, AV_PIX_FMT_NV12 = 0, 8, 23


class CudaFilter:
    """One instance of the product's ``transform360_cuda`` filter on a link of AV_PIX_FMT_CUDA frames.

    ``device=False`` configures the graph only (option table, format negotiation, config_output: no GPU needed).  With a
    device the filter is given torch's current CUDA context and ``stream`` (a torch.cuda.Stream), the CUDA driver's own
    cuCtxPushCurrent / cuCtxPopCurrent / cuStreamSynchronize as ffmpeg's dynlink table would, and planes that are
    torch tensors on the GPU."""

    def __init__(self, args: str, in_w: int, in_h: int, sw_format: int = AV_PIX_FMT_YUV420P, device: bool = True, stream=None):
        self.L = _lib("cuda")
        err = C.c_int()
        ctx, fns, self.stream = None, None, stream
        if device:
            import torch
            torch.cuda.init()
            torch.zeros(1, device="cuda")  # makes the primary context current
            drv = C.CDLL("libcuda.so.1")
            cur = C.c_void_p()
            assert drv.cuCtxGetCurrent(C.byref(cur)) == 0 and cur.value
            ctx = cur
            self._fns = (C.c_void_p * 3)(C.cast(drv.cuCtxPushCurrent_v2, C.c_void_p), C.cast(drv.cuCtxPopCurrent_v2, C.c_void_p),
                                         C.cast(drv.cuStreamSynchronize, C.c_void_p))
            fns = self._fns
        self.h = self.L.t360f_open_cuda(args.encode(), in_w, in_h, sw_format, ctx, C.c_void_p(stream.cuda_stream if stream is not None else 0),
                                        fns, C.byref(err))
        if not self.h:
            raise ValueError(f"filter rejected {args!r}: AVERROR {err.value}")
        self.in_w, self.in_h, self.planes = in_w, in_h, 1 if sw_format == AV_PIX_FMT_GRAY8 else 3
        w, h = C.c_int(), C.c_int()
        self.L.t360f_out_size(self.h, C.byref(w), C.byref(h))
        self.out_w, self.out_h = w.value, h.value

    def close(self):
        if getattr(self, "h", None):
            self.L.t360f_close(self.h)
            self.h = None

    __del__ = close

    def filter(self, planes, out_pitch_pad: int = 0):
        """planes: uint8 CUDA tensors (2-D, any row stride); returns the output planes as CUDA tensors."""
        import torch
        shapes = [(self.out_h, self.out_w)] + [((self.out_h + 1) // 2, (self.out_w + 1) // 2)] * (self.planes - 1)
        store = [torch.full((h, w + out_pitch_pad), 0xA5, dtype=torch.uint8, device="cuda") for h, w in shapes]
        outs = [s[:, :w] for s, (h, w) in zip(store, shapes)]
        for o in outs:
            o.zero_()  # a fresh software frame of the stand-in is zeroed too; BORDER_TRANSPARENT layouts keep what they find
        n = self.planes
        VP, IA = C.c_void_p * n, C.c_int * n
        rc = self.L.t360f_filter_cuda(self.h, VP(*[p.data_ptr() for p in planes]), IA(*[p.stride(0) for p in planes]),
                                      VP(*[o.data_ptr() for o in outs]), IA(*[o.stride(0) for o in outs]))
        if rc:
            raise RuntimeError(f"filter_frame failed: AVERROR {rc}")
        self.last_store = store
        return outs

"""ctypes wrapper of the bench-only upstream-structured GPU stand-in (libstandin.so, see standin.cu).

NOT part of the product: only bench.py's `gpu_baseline` leg and tests/test_gpu_standin.py load it.  It gives
libghr's stages K2-K7 a GPU-class denominator (global CUB sort + blocking read of the instance count + one CTA
per tile + per-thread atomicAdd) when the reference's own rasterizer package
(/root/reference/environment.yml:129, imported at /root/reference/tgs/models/renderer_one_shot.py:3) is absent.
The per-Gaussian projected geometry comes from libghr's own preprocess (the geometry block of a ghr_forward
state) -- preprocess is NOT part of what the stand-in times."""
import ctypes as C
import os

import torch

from . import build as _build

_lib = None


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libstandin.so")
        if not os.path.exists(path):
            path = _build.build()
        L = C.CDLL(path)
        vp = C.c_void_p
        L.sgs_create.argtypes = [C.POINTER(vp)]
        L.sgs_destroy.argtypes = [vp]
        L.sgs_unpack.argtypes = [vp, C.c_int, vp, vp]
        L.sgs_forward.argtypes = [vp, C.c_int, C.c_int, vp, vp, vp]
        L.sgs_backward.argtypes = [vp, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp, vp]
        L.sgs_num_rendered.argtypes = [vp]
        L.sgs_num_rendered.restype = C.c_uint32
        for f in ("sgs_n_contrib", "sgs_final_T", "sgs_ranges", "sgs_sorted_keys"):
            getattr(L, f).argtypes = [vp]
            getattr(L, f).restype = vp
        L.sgs_copy.argtypes = [vp, vp, C.c_size_t, vp]
        _lib = L
    return _lib


def _stream(dev):
    return torch._C._cuda_getCurrentRawStream(dev.index if dev.index is not None else torch.cuda.current_device())


class Standin:
    """One view's binning + blend, upstream-structured.  geom: the [P, 16] float32 geometry block of ONE view of a
    libghr forward state (api.forward_raw(...).state at layout off_geom)."""

    def __init__(self, geom: torch.Tensor, H: int, W: int, bg: torch.Tensor):
        self.L = lib()
        self.h = C.c_void_p()
        if self.L.sgs_create(C.byref(self.h)):
            raise RuntimeError("sgs_create failed")
        assert geom.is_cuda and geom.dtype == torch.float32 and geom.is_contiguous() and geom.data_ptr() % 16 == 0
        self.dev = geom.device
        self.P, self.H, self.W = geom.numel() // 16, H, W
        self.bg = bg.to(self.dev, torch.float32).contiguous()
        self.geom = geom
        if self.L.sgs_unpack(self.h, self.P, geom.data_ptr(), _stream(self.dev)):
            raise RuntimeError("sgs_unpack failed")
        self.color = torch.empty(3, H, W, dtype=torch.float32, device=self.dev)
        P = max(self.P, 1)
        self.dL_dmean2D = torch.empty(P, 2, dtype=torch.float32, device=self.dev)
        self.dL_dconic = torch.empty(P, 3, dtype=torch.float32, device=self.dev)
        self.dL_dopacity = torch.empty(P, dtype=torch.float32, device=self.dev)
        self.dL_dcolors = torch.empty(P, 3, dtype=torch.float32, device=self.dev)

    def forward(self) -> torch.Tensor:
        """K2-K6; blocks the host once on the instance count like upstream."""
        if self.L.sgs_forward(self.h, self.H, self.W, self.bg.data_ptr(), self.color.data_ptr(), _stream(self.dev)):
            raise RuntimeError("sgs_forward failed")
        return self.color

    def backward(self, dL_dout: torch.Tensor):
        """K7.  dL_dout [3,H,W]."""
        assert dL_dout.is_contiguous() and dL_dout.dtype == torch.float32
        if self.L.sgs_backward(self.h, self.H, self.W, self.bg.data_ptr(), dL_dout.data_ptr(),
                               self.dL_dmean2D.data_ptr(), self.dL_dconic.data_ptr(), self.dL_dopacity.data_ptr(),
                               self.dL_dcolors.data_ptr(), _stream(self.dev)):
            raise RuntimeError("sgs_backward failed")
        return self.dL_dmean2D, self.dL_dconic, self.dL_dopacity, self.dL_dcolors

    @property
    def R(self) -> int:
        return int(self.L.sgs_num_rendered(self.h))

    def _view(self, fn, nbytes, dtype):
        p = getattr(self.L, fn)(self.h)
        out = torch.empty(max(nbytes, 8), dtype=torch.uint8, device=self.dev)
        if nbytes and self.L.sgs_copy(out.data_ptr(), p, nbytes, _stream(self.dev)):
            raise RuntimeError("sgs_copy failed")
        return out[:nbytes].view(dtype)

    def n_contrib(self):
        return self._view("sgs_n_contrib", self.H * self.W * 4, torch.int32).view(self.H, self.W)

    def final_T(self):
        return self._view("sgs_final_T", self.H * self.W * 4, torch.float32).view(self.H, self.W)

    def sorted_keys(self):
        return self._view("sgs_sorted_keys", self.R * 8, torch.int64)

    def ranges(self):
        gx, gy = (self.W + 15) // 16, (self.H + 15) // 16
        return self._view("sgs_ranges", gx * gy * 8, torch.int32).view(gx * gy, 2)

    def close(self):
        if self.h:
            torch.cuda.synchronize(self.dev)
            self.L.sgs_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

"""Thin object layer over the C ABI: contexts and device handles with deterministic clean-up.

Mirrors what the Julia glue does with finalizers (julia/src/libblr.jl).  PyTorch is optional here and used
only as plumbing (wrapping CUDA tensors as borrowed device pointers, torch.distributed to ship the NCCL id).
"""
from __future__ import annotations

import ctypes as C
import os
import weakref
from typing import Optional, Tuple

import numpy as np

from . import _lib as L


def _f64(a, order="F") -> np.ndarray:
    return np.require(a, dtype=np.float64, requirements=["F_CONTIGUOUS" if order == "F" else "C_CONTIGUOUS", "ALIGNED"])


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Context:
    """One context drives one GPU (blr_ctx_create).  Not re-entrant."""

    def __init__(self, device: Optional[int] = None):
        self.lib = L.load()
        if device is None:
            device = int(os.environ.get("LOCAL_RANK", "0")) if "LOCAL_RANK" in os.environ else 0
        h = C.c_void_p()
        rc = self.lib.blr_ctx_create(C.byref(h), int(device))
        if rc == L.E_NODEVICE:
            raise L.BLRError(rc, "no sm_100 (B200) device available; libblr_cuda has no CPU fallback")
        if rc != 0:
            raise L.BLRError(rc, "blr_ctx_create failed")
        self.handle = h
        self.device = int(device)
        self.rank, self.world = 0, 1
        self._pinned_free, self._pinned_all = {}, []
        self._fin = weakref.finalize(self, self.lib.blr_ctx_destroy, h)

    # ---- pinned result buffers ---------------------------------------------------------------------
    def empty_pinned(self, shape, order="F", min_bytes: int = 1 << 15) -> np.ndarray:
        """np.empty in page-locked memory (results download at PCIe speed and truly asynchronously: a D2H copy into pageable
        memory is staged by the driver and blocks the host).  Buffers are recycled through a per-context free list once
        every numpy view of them is gone; results below `min_bytes` stay pageable."""
        n = int(np.prod(shape))
        nbytes = n * 8
        if nbytes < min_bytes:
            return np.empty(shape, dtype=np.float64, order=order)
        free = self._pinned_free.setdefault(nbytes, [])
        if free:
            ptr = free.pop()
        else:
            out = C.c_void_p()
            self.check(self.lib.blr_host_alloc(self.handle, nbytes, C.byref(out)))
            ptr = out.value
            self._pinned_all.append(ptr)
        buf = (C.c_char * nbytes).from_address(ptr)
        weakref.finalize(buf, free.append, ptr)  # back to the free list when the last view dies
        return np.frombuffer(buf, dtype=np.float64).reshape(shape, order=order)

    # ---- error mapping -------------------------------------------------------------------------
    def check(self, rc: int):
        if rc == 0:
            return
        if rc > 0:
            raise L.PosDefException(rc)
        msg = (self.lib.blr_last_error(self.handle) or b"").decode("utf-8", "replace")
        if rc == L.E_DIM:
            raise L.DimensionMismatch(rc, msg)
        raise L.BLRError(rc, msg)

    def sync(self):
        self.check(self.lib.blr_ctx_sync(self.handle))

    def set_form(self, form: str):
        """Numerical form of the D x D phase and the factor applications: "direct" (default; chol(Λw + G), inverse-factor GEMMs)
        or "whitened" -- the reference's literal evaluation order (src/bayesian_linear_regression.jl:81,86,64-68 and the
        triangular solves of :36,:41,:51).  See include/blr_cuda.h, blr_ctx_set_form."""
        forms = {"direct": 0, "whitened": 1}
        if form not in forms:
            raise ValueError(f"form must be one of {sorted(forms)}")
        self.check(self.lib.blr_ctx_set_form(self.handle, forms[form]))

    def form(self) -> str:
        v = C.c_int()
        self.check(self.lib.blr_ctx_get_form(self.handle, C.byref(v)))
        return ("direct", "whitened")[v.value]

    def stream(self) -> int:
        s = C.c_void_p()
        self.check(self.lib.blr_ctx_stream(self.handle, C.byref(s)))
        return int(s.value or 0)

    def wait_torch_stream(self, t=None):
        """Order this context's (private, non-blocking) stream after the work already enqueued on torch's current stream
        of this device: borrowed CUDA tensors may still be being written by the caller's kernels."""
        import torch

        s = torch.cuda.current_stream(t.device if t is not None else self.device)
        self.check(self.lib.blr_ctx_wait_stream(self.handle, C.c_void_p(s.cuda_stream)))

    def release_to_torch_stream(self, device=None):
        """The other direction: torch's current stream waits for everything this context has enqueued (results written
        into borrowed tensors by the *_dev entry points)."""
        import torch

        s = torch.cuda.current_stream(self.device if device is None else device)
        self.check(self.lib.blr_stream_wait_ctx(self.handle, C.c_void_p(s.cuda_stream)))

    def launch_count(self) -> int:
        return int(self.lib.blr_launch_count(self.handle))

    def last_timings(self) -> dict:
        out = (C.c_double * 8)()
        self.check(self.lib.blr_last_timings(self.handle, out))
        return {"prep_ms": out[0], "gram_ms": out[1], "reduce_ms": out[2], "solve_ms": out[3]}

    # ---- multi-GPU -----------------------------------------------------------------------------
    def init_comm(self, nranks: int, rank: int, unique_id: bytes):
        buf = C.create_string_buffer(unique_id, 128)
        self.check(self.lib.blr_comm_init_rank(self.handle, buf, int(nranks), int(rank)))
        self.rank, self.world = int(rank), int(nranks)

    def new_unique_id(self) -> bytes:
        buf = C.create_string_buffer(128)
        rc = self.lib.blr_nccl_unique_id(buf)
        if rc != 0:
            raise L.BLRError(rc, "ncclGetUniqueId failed (libnccl not loadable?)")
        return buf.raw

    def init_comm_from_torch(self):
        """Create the NCCL communicator of this context with torch.distributed carrying the 128-byte id."""
        import torch.distributed as dist

        if not dist.is_initialized() or dist.get_world_size() == 1:
            return
        rank, world = dist.get_rank(), dist.get_world_size()
        box = [self.new_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        self.init_comm(world, rank, box[0])

    @staticmethod
    def init_comm_all(ctxs):
        """One process, several GPUs: make `ctxs` (one Context per device) the ranks 0..n-1 of one communicator."""
        arr = (C.c_void_p * len(ctxs))(*[c.handle for c in ctxs])
        ctxs[0].check(ctxs[0].lib.blr_comm_init_all(arr, len(ctxs)))
        for i, c in enumerate(ctxs):
            c.rank, c.world = i, len(ctxs)

    # ---- calibration ---------------------------------------------------------------------------
    def calibrate(self) -> dict:
        out = {}
        for key, fn in (("dmma_tflops", self.lib.blr_calibrate_dmma), ("dfma_tflops", self.lib.blr_calibrate_dfma),
                        ("hbm_gbs", self.lib.blr_calibrate_hbm)):
            v = C.c_double()
            self.check(fn(self.handle, C.byref(v)))
            out[key] = v.value
        return out


_default_ctx: Optional[Context] = None


def default_context() -> Context:
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context()
    return _default_ctx


def set_default_context(ctx: Optional[Context]):
    global _default_ctx
    _default_ctx = ctx


class DeviceMatrix:
    """Device-resident design matrix (blr_x): D features x N inputs in ColVecs or RowVecs layout."""

    def __init__(self, ctx: Context, handle, D: int, N: int, layout: int, keepalive=None):
        self.ctx, self.handle, self.D, self.N, self.layout = ctx, handle, int(D), int(N), int(layout)
        self._keep = keepalive
        self._fin = weakref.finalize(self, ctx.lib.blr_x_free, ctx.handle, handle)

    @staticmethod
    def upload(ctx: Context, A: np.ndarray, layout: int) -> "DeviceMatrix":
        """A is the Julia-side matrix: D x N for ColVecs, N x D for RowVecs; sent column-major."""
        A = np.asarray(A, dtype=np.float64)
        if A.ndim != 2:
            raise L.BLRError(L.E_INVALID, "expected a matrix")
        Af = _f64(A, "F")
        D, N = (Af.shape if layout == L.COLVECS else Af.shape[::-1])
        h = C.c_void_p()
        ctx.check(ctx.lib.blr_x_upload(ctx.handle, _ptr(Af), D, N, max(Af.shape[0], 1), layout, C.byref(h)))
        return DeviceMatrix(ctx, h, D, N, layout)

    @staticmethod
    def alloc(ctx: Context, D: int, N: int, layout: int = L.COLVECS) -> "DeviceMatrix":
        h = C.c_void_p()
        ctx.check(ctx.lib.blr_x_alloc(ctx.handle, D, N, layout, C.byref(h)))
        return DeviceMatrix(ctx, h, D, N, layout)

    @staticmethod
    def wrap_torch(ctx: Context, t, layout: int) -> "DeviceMatrix":
        """Borrow a 2-D float64 CUDA tensor holding the column-major matrix as its TRANSPOSE in torch's
        row-major convention: ColVecs -> tensor of shape (N, D); RowVecs -> tensor of shape (D, N)."""
        if not (t.is_cuda and t.dim() == 2 and t.dtype.is_floating_point and t.element_size() == 8 and t.stride(1) == 1):
            raise L.BLRError(L.E_INVALID, "expected a 2-D float64 CUDA tensor with unit inner stride")
        cols, rows = t.shape
        D, N = (rows, cols) if layout == L.COLVECS else (cols, rows)
        ctx.wait_torch_stream(t)  # the tensor's producer may still be running on torch's stream
        h = C.c_void_p()
        ctx.check(ctx.lib.blr_x_wrap_device(ctx.handle, C.c_void_p(t.data_ptr()), D, N, max(t.stride(0), 1), layout, C.byref(h)))
        return DeviceMatrix(ctx, h, D, N, layout, keepalive=t)

    def as_torch(self):
        """Zero-copy torch view of the device matrix (the inverse of wrap_torch): ColVecs -> tensor of shape (N, D), RowVecs ->
        (D, N), row stride = the leading dimension.  torch's current stream is ordered after this context's pending work;
        the view keeps the handle alive."""
        import torch

        self.ctx.release_to_torch_stream()
        ptr, ld = self.device_ptr()
        rows, cols = (self.N, self.D) if self.layout == L.COLVECS else (self.D, self.N)
        if rows == 0 or cols == 0:
            return torch.empty((rows, cols), dtype=torch.float64, device=f"cuda:{self.ctx.device}")

        class _View:
            pass

        v = _View()
        v.__cuda_array_interface__ = {"shape": (rows, cols), "typestr": "<f8", "data": (ptr, False), "version": 2,
                                      "strides": (ld * 8, 8)}
        v.owner = self
        t = torch.as_tensor(v, device=f"cuda:{self.ctx.device}")
        t._blr_owner = self  # the tensor must not outlive the library-owned memory
        return t

    def device_ptr(self) -> Tuple[int, int]:
        p, ld = C.c_void_p(), C.c_int64()
        self.ctx.check(self.ctx.lib.blr_x_device_ptr(self.ctx.handle, self.handle, C.byref(p), C.byref(ld)))
        return int(p.value or 0), int(ld.value)

    def synth_(self, seed: int, n_offset: int = 0) -> "DeviceMatrix":
        self.ctx.check(self.ctx.lib.blr_x_synth(self.ctx.handle, self.handle, seed, n_offset))
        return self

    def __len__(self):
        return self.N


class DeviceVector:
    """Device-resident N-vector (blr_vec)."""

    def __init__(self, ctx: Context, handle, n: int, keepalive=None):
        self.ctx, self.handle, self.n = ctx, handle, int(n)
        self._keep = keepalive
        self._fin = weakref.finalize(self, ctx.lib.blr_vec_free, ctx.handle, handle)

    @staticmethod
    def upload(ctx: Context, v) -> "DeviceVector":
        v = _f64(np.asarray(v, dtype=np.float64).reshape(-1))
        h = C.c_void_p()
        ctx.check(ctx.lib.blr_vec_upload(ctx.handle, _ptr(v), v.shape[0], C.byref(h)))
        return DeviceVector(ctx, h, v.shape[0])

    @staticmethod
    def alloc(ctx: Context, n: int) -> "DeviceVector":
        h = C.c_void_p()
        ctx.check(ctx.lib.blr_vec_alloc(ctx.handle, n, C.byref(h)))
        return DeviceVector(ctx, h, n)

    @staticmethod
    def wrap_torch(ctx: Context, t) -> "DeviceVector":
        if not (t.is_cuda and t.dim() == 1 and t.element_size() == 8 and t.is_contiguous()):
            raise L.BLRError(L.E_INVALID, "expected a contiguous 1-D float64 CUDA tensor")
        ctx.wait_torch_stream(t)
        h = C.c_void_p()
        ctx.check(ctx.lib.blr_vec_wrap_device(ctx.handle, C.c_void_p(t.data_ptr()), t.shape[0], C.byref(h)))
        return DeviceVector(ctx, h, t.shape[0], keepalive=t)

    def device_ptr(self) -> int:
        p = C.c_void_p()
        self.ctx.check(self.ctx.lib.blr_vec_device_ptr(self.ctx.handle, self.handle, C.byref(p)))
        return int(p.value or 0)

    def download(self) -> np.ndarray:
        out = np.empty(self.n, dtype=np.float64)
        self.ctx.check(self.ctx.lib.blr_vec_download(self.ctx.handle, self.handle, _ptr(out)))
        return out

    def __len__(self):
        return self.n


class Stats:
    """Packed sufficient statistics [G | r | q | ℓ | n] on the device (blr_stats)."""

    def __init__(self, ctx: Context, D: int):
        h = C.c_void_p()
        ctx.check(ctx.lib.blr_stats_create(ctx.handle, D, C.byref(h)))
        self.ctx, self.handle, self.D = ctx, h, int(D)
        self._fin = weakref.finalize(self, ctx.lib.blr_stats_free, ctx.handle, h)

    def __len__(self):
        return self.D * self.D + self.D + 3

    def zero(self):
        self.ctx.check(self.ctx.lib.blr_stats_zero(self.ctx.handle, self.handle))

    def allreduce(self):
        self.ctx.check(self.ctx.lib.blr_stats_allreduce(self.ctx.handle, self.handle))

    @staticmethod
    def allreduce_all(stats):
        """Grouped in-place sum over the per-device statistics of a single-process communicator (Context.init_comm_all)."""
        ctxs = (C.c_void_p * len(stats))(*[s.ctx.handle for s in stats])
        hs = (C.c_void_p * len(stats))(*[s.handle for s in stats])
        stats[0].ctx.check(stats[0].ctx.lib.blr_stats_allreduce_all(ctxs, hs, len(stats)))

    def download(self) -> np.ndarray:
        out = np.empty(len(self), dtype=np.float64)
        self.ctx.check(self.ctx.lib.blr_stats_download(self.ctx.handle, self.handle, _ptr(out)))
        return out

    def upload(self, packed: np.ndarray):
        packed = _f64(np.asarray(packed).reshape(-1))
        if packed.shape[0] != len(self):
            raise L.BLRError(L.E_INVALID, "packed statistics have the wrong length")
        self.ctx.check(self.ctx.lib.blr_stats_upload(self.ctx.handle, self.handle, _ptr(packed)))

    def unpack(self):
        p, D = self.download(), self.D
        return p[: D * D].reshape(D, D, order="F"), p[D * D : D * D + D], p[D * D + D], p[D * D + D + 1], p[D * D + D + 2]


class DevicePosterior:
    """Device-resident regressor (blr_post): mw, chol(Λw) and (lazily) its inverse factor, cached for predict."""

    def __init__(self, ctx: Context, handle, D: int):
        self.ctx, self.handle, self.D = ctx, handle, int(D)
        self._fin = weakref.finalize(self, ctx.lib.blr_post_free, ctx.handle, handle)


def make_noise(ctx: Context, Σy, N: int):
    """FiniteGP noise normalisation -> (Noise struct, keepalive).  Real -> Diagonal(Fill), vector -> Diagonal."""
    from .model import Diagonal  # local import: model imports runtime

    from .model import PDMat, Symmetric

    if isinstance(Σy, DeviceVector):
        return L.Noise(L.NOISE_VECTOR, 0.0, Σy.handle, None, 0), Σy
    if isinstance(Σy, Diagonal):
        Σy = Σy.diag
    if isinstance(Σy, (Symmetric, PDMat)):
        Σy = Σy.dense()
    if isinstance(Σy, DeviceVector):
        return L.Noise(L.NOISE_VECTOR, 0.0, Σy.handle, None, 0), Σy
    try:
        import torch

        if isinstance(Σy, torch.Tensor) and Σy.is_cuda:
            v = DeviceVector.wrap_torch(ctx, Σy)
            return L.Noise(L.NOISE_VECTOR, 0.0, v.handle, None, 0), v
    except ImportError:  # pragma: no cover
        pass
    a = np.asarray(Σy, dtype=np.float64)
    if a.ndim == 0:
        return L.Noise(L.NOISE_SCALAR, float(a), None, None, 0), None
    if a.ndim == 1:
        if a.shape[0] != N:
            raise L.DimensionMismatch(L.E_DIM, "length(diag(Σy)) != number of inputs")
        v = DeviceVector.upload(ctx, a)
        return L.Noise(L.NOISE_VECTOR, 0.0, v.handle, None, 0), v
    if a.ndim == 2:  # dense Σy: the small-N side path (whitening on the device)
        if a.shape != (N, N):
            raise L.DimensionMismatch(L.E_DIM, "size(Σy) does not match the number of inputs")
        af = _f64(a, "F")
        return L.Noise(L.NOISE_DENSE, 0.0, None, af.ctypes.data_as(C.c_void_p), N), af
    raise L.BLRError(L.E_INVALID, "unsupported observation-noise argument")

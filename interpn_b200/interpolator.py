"""Resident interpolators: the reference's struct API (`X::new(..)?.interp(obs, out)`,
e.g. /root/reference/src/multilinear/regular.rs:225-283) with the grid kept in HBM.

This is a thin ctypes mirror of the `interpn_b200_*_new_*` / `interp_eval_*` entry points of
include/interpn_b200.h. Host arrays go through `eval`; device-resident data (raw device pointers
or torch CUDA tensors) go through `eval_device` / `eval_torch`, which only enqueue a kernel on the
given stream.
"""

from __future__ import annotations

import ctypes as C
from collections.abc import Sequence

import numpy as np

from . import _lib
from ._lib import lib
from .raw import _CT, _arr, _dims, _p, _ptrs

_METHODS = {"linear": _lib.LINEAR, "cubic": _lib.CUBIC, "nearest": _lib.NEAREST}


def _sfx(dtype) -> str:
    dtype = np.dtype(dtype)
    if dtype == np.float64:
        return "f64"
    if dtype == np.float32:
        return "f32"
    raise TypeError(f"Unexpected data type: {dtype}")


def _is_torch_cuda(x) -> bool:
    return hasattr(x, "data_ptr") and hasattr(x, "is_cuda") and bool(x.is_cuda)


class Interpolator:
    """A grid resident on the current CUDA device plus an evaluation method."""

    def __init__(self, handle: int, sfx: str, ndims: int):
        self._h = C.c_void_p(handle)
        self._sfx = sfx
        self.ndims = ndims
        self.dtype, self._ct = _CT[sfx]

    # ---- constructors -------------------------------------------------------------------
    @staticmethod
    def _vals_arg(vals, dt, ct, nvals_if_uninit):
        """(pointer, nvals, location) for host ndarray / torch CUDA tensor / None (uninitialised)."""
        if vals is None:
            return None, int(nvals_if_uninit), _lib.VALS_UNINIT, None
        if _is_torch_cuda(vals):
            import torch

            want = torch.float64 if dt == np.float64 else torch.float32
            if vals.dtype != want or not vals.is_contiguous():
                raise TypeError("device `vals` must be a contiguous CUDA tensor of the interpolator dtype")
            return C.cast(C.c_void_p(vals.data_ptr()), C.POINTER(ct)), vals.numel(), _lib.VALS_DEVICE, vals
        v = _arr(vals, dt, "vals")
        return _p(v, ct), v.size, _lib.VALS_HOST, v

    @classmethod
    def regular(cls, method: str, dims, starts, steps, vals, linearize_extrapolation: bool = True, dtype=None):
        """`MultilinearRegular::new` / `MulticubicRegular::new` / `NearestRegular::new`.

        `vals` may be a host ndarray, a torch CUDA tensor (copied device-to-device) or None
        (storage allocated, left for the caller to fill — see `vals_ptr`, used by the multi-GPU
        broadcast in interpn_b200.distributed).
        """
        sfx = _sfx(dtype if dtype is not None else (vals.dtype if isinstance(vals, np.ndarray) else np.asarray(starts).dtype))
        dt, ct = _CT[sfx]
        dims_a = _dims(dims)
        starts, steps = _arr(starts, dt, "starts"), _arr(steps, dt, "steps")
        vptr, nvals, loc, keep = cls._vals_arg(vals, dt, ct, int(np.prod(dims_a, dtype=np.uint64)) if len(dims_a) else 0)
        h = C.c_void_p()
        fn = getattr(lib, f"interpn_b200_regular_new_{sfx}")
        _lib.check(
            fn(
                C.c_int(_METHODS[method]),
                dims_a.ctypes.data_as(C.POINTER(C.c_size_t)), C.c_size_t(dims_a.size),
                _p(starts, ct), C.c_size_t(starts.size), _p(steps, ct), C.c_size_t(steps.size),
                vptr, C.c_size_t(nvals), C.c_int(int(bool(linearize_extrapolation))), C.c_int(loc), C.byref(h),
            )  # fmt: skip
        )
        del keep
        return cls(h.value, sfx, int(dims_a.size))

    @classmethod
    def rectilinear(cls, method: str, grids: Sequence[np.ndarray], vals, linearize_extrapolation: bool = True, dtype=None):
        """`MultilinearRectilinear::new` / `MulticubicRectilinear::new` / `NearestRectilinear::new`."""
        sfx = _sfx(dtype if dtype is not None else grids[0].dtype)
        dt, ct = _CT[sfx]
        grids = [_arr(g, dt, "grids") for g in grids]
        gptrs, glens = _ptrs(grids, ct)
        nv = 1
        for g in grids:
            nv *= g.size
        vptr, nvals, loc, keep = cls._vals_arg(vals, dt, ct, nv)
        h = C.c_void_p()
        fn = getattr(lib, f"interpn_b200_rectilinear_new_{sfx}")
        _lib.check(
            fn(
                C.c_int(_METHODS[method]), gptrs, glens, C.c_size_t(len(grids)),
                vptr, C.c_size_t(nvals), C.c_int(int(bool(linearize_extrapolation))), C.c_int(loc), C.byref(h),
            )  # fmt: skip
        )
        del keep
        return cls(h.value, sfx, len(grids))

    # ---- evaluation ---------------------------------------------------------------------
    def eval(self, obs: Sequence[np.ndarray], out: np.ndarray | None = None) -> np.ndarray:
        """`.interp(obs, out)` on host arrays (copies in and out; reference error semantics)."""
        obs = [_arr(o, self.dtype, "obs") for o in obs]
        if out is None:
            out = np.zeros_like(obs[0]) if obs else np.zeros(0, dtype=self.dtype)
        out = _arr(out, self.dtype, "out", writable=True)
        optrs, olens = _ptrs(obs, self._ct)
        fn = getattr(lib, f"interpn_b200_interp_eval_host_{self._sfx}")
        fb = C.c_size_t(_lib.NO_BAD)
        st = fn(self._h, optrs, olens, C.c_size_t(len(obs)), _p(out, self._ct), C.c_size_t(out.size), C.byref(fb))
        self.first_bad = None if fb.value == _lib.NO_BAD else int(fb.value)
        _lib.check(st)
        return out

    def eval_device(self, obs_ptrs: Sequence[int], n: int, out_ptr: int, stream: int = 0) -> None:
        """Enqueue `.interp` on raw device pointers (stream-ordered, no synchronisation)."""
        k = len(obs_ptrs)
        ptrs = (C.POINTER(self._ct) * max(k, 1))()
        for i, p in enumerate(obs_ptrs):
            ptrs[i] = C.cast(C.c_void_p(int(p)), C.POINTER(self._ct))
        fn = getattr(lib, f"interpn_b200_interp_eval_device_{self._sfx}")
        _lib.check(
            fn(self._h, ptrs, C.c_size_t(k), C.c_size_t(int(n)), C.cast(C.c_void_p(int(out_ptr)), C.POINTER(self._ct)),
               C.c_void_p(int(stream)))  # fmt: skip
        )

    @staticmethod
    def eval_fields_device(interps: Sequence["Interpolator"], obs_ptrs: Sequence[int], n: int, out_ptrs: Sequence[int],
                           stream: int = 0) -> None:
        """Several fields over one grid and one query batch: `interps` were built over the same grid
        with the same method, `out_ptrs[k]` receives field k. Multilinear and nearest fields on grids within L2 share
        one cell location per point; results equal `len(interps)` separate `eval_device` calls bit for bit. Failures
        are latched on `interps[0]` (`interps[0].status(stream)`)."""
        first = interps[0]
        if any(i._sfx != first._sfx for i in interps) or len(out_ptrs) != len(interps):
            raise AssertionError("Dimension mismatch")
        ct = first._ct
        hs = (C.c_void_p * len(interps))(*[i._h for i in interps])
        k = len(obs_ptrs)
        optrs = (C.POINTER(ct) * max(k, 1))()
        for i, p in enumerate(obs_ptrs):
            optrs[i] = C.cast(C.c_void_p(int(p)), C.POINTER(ct))
        outs = (C.POINTER(ct) * len(interps))()
        for i, p in enumerate(out_ptrs):
            outs[i] = C.cast(C.c_void_p(int(p)), C.POINTER(ct))
        fn = getattr(lib, f"interpn_b200_interp_eval_fields_device_{first._sfx}")
        _lib.check(fn(hs, C.c_size_t(len(interps)), optrs, C.c_size_t(k), C.c_size_t(int(n)), outs, C.c_void_p(int(stream))))

    @staticmethod
    def eval_fields_torch(interps: Sequence["Interpolator"], obs, outs=None, stream=None):
        """`eval_fields_device` on torch CUDA tensors (torch's current stream unless `stream` is given)."""
        import torch

        n = obs[0].numel() if len(obs) else 0
        want = torch.float64 if interps[0].dtype == np.float64 else torch.float32
        for o in obs:
            if not (o.is_cuda and o.dtype == want and o.is_contiguous() and o.dim() == 1 and o.numel() == n):
                raise TypeError("obs must be contiguous 1-D CUDA tensors of the interpolators' dtype and one length")
        if outs is None:
            outs = [torch.empty(n, dtype=want, device=obs[0].device) for _ in interps]
        s = stream if stream is not None else torch.cuda.current_stream(obs[0].device)
        Interpolator.eval_fields_device(interps, [o.data_ptr() for o in obs], n, [o.data_ptr() for o in outs], s.cuda_stream)
        return outs

    def eval_cuda_arrays(self, obs, out, stream: int = 0) -> None:
        """Enqueue `.interp` on any objects that expose ``__cuda_array_interface__`` (CuPy and Numba arrays, torch
        CUDA tensors, RMM buffers): zero-copy, stream-ordered, no synchronisation. Every array must
        be a contiguous 1-D array of the interpolator's dtype; `out` must be writable."""
        want = "<f8" if self.dtype == np.float64 else "<f4"

        def ptr_len(a, what, writable=False):
            cai = a.__cuda_array_interface__
            if cai["typestr"] != want or len(cai["shape"]) != 1:
                raise TypeError(f"{what} must be a 1-D CUDA array of dtype {self.dtype}")
            st = cai.get("strides")
            if st is not None and tuple(st) != (self.dtype.itemsize,):
                raise TypeError(f"{what} must be contiguous")
            p, readonly = cai["data"]
            if writable and readonly:
                raise TypeError(f"{what} must be writable")
            return int(p), int(cai["shape"][0])

        pl = [ptr_len(o, "obs") for o in obs]
        op, n = ptr_len(out, "out", writable=True)
        if any(m != n for _, m in pl):
            raise AssertionError("Dimension mismatch")
        self.eval_device([p for p, _ in pl], n, op, stream)

    def eval_torch(self, obs, out=None, stream=None):
        """Enqueue `.interp` on torch CUDA tensors on torch's current stream (or `stream`)."""
        import torch

        want = torch.float64 if self.dtype == np.float64 else torch.float32
        for o in obs:
            if not (o.is_cuda and o.dtype == want and o.is_contiguous() and o.dim() == 1):
                raise TypeError("obs must be contiguous 1-D CUDA tensors of the interpolator dtype")
        n = obs[0].numel() if len(obs) else 0
        if any(o.numel() != n for o in obs):
            raise AssertionError("Dimension mismatch")
        if out is None:
            out = torch.empty(n, dtype=want, device=obs[0].device)
        elif out.numel() != n:
            raise AssertionError("Dimension mismatch")
        s = stream if stream is not None else torch.cuda.current_stream(obs[0].device)
        self.eval_device([o.data_ptr() for o in obs], n, out.data_ptr(), s.cuda_stream)
        return out

    def status(self, stream: int = 0) -> None:
        """Synchronise `stream` and raise AssertionError("Unrepresentable coordinate value") if any
        device evaluation since the last call met such a point; `.first_bad` holds its index."""
        fb = C.c_size_t(0)
        st = lib.interpn_b200_interp_status(self._h, C.c_void_p(int(stream)), C.byref(fb))
        self.first_bad = None if fb.value == _lib.NO_BAD else int(fb.value)
        _lib.check(st)

    # ---- residency ------------------------------------------------------------------------
    @property
    def vals_ptr(self) -> int:
        return int(lib.interpn_b200_interp_vals_ptr(self._h) or 0)

    @property
    def vals_len(self) -> int:
        return int(lib.interpn_b200_interp_vals_len(self._h))

    def vals_tensor(self):
        """A torch CUDA tensor aliasing the resident `vals` storage (no copy) — the buffer a
        `torch.distributed.broadcast` fills when the grid is replicated across ranks."""
        import torch

        class _Alias:
            pass

        a = _Alias()
        a.__cuda_array_interface__ = {
            "shape": (self.vals_len,),
            "typestr": "<f8" if self.dtype == np.float64 else "<f4",
            "data": (self.vals_ptr, False),
            "version": 2,
        }
        t = torch.as_tensor(a, device="cuda")
        t._interpn_owner = self  # keep the storage alive while the alias exists
        return t

    def vals_updated(self, stream: int = 0) -> None:
        """Tell the library the resident `vals` were rewritten through `vals_ptr` / `vals_tensor()`
        (e.g. by the grid broadcast): refreshes its gather-optimised copies, ordered on `stream`."""
        _lib.check(lib.interpn_b200_interp_vals_updated(self._h, C.c_void_p(int(stream))))

    def close(self) -> None:
        if self._h:
            lib.interpn_b200_interp_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

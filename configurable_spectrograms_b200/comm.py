"""Rank plumbing for the one exchange step of the path (global extrema).

``torch.distributed`` carries the collectives: NCCL over NVLink/NVSwitch on the GPU box
(tensors stay on the device), gloo in the CPU tests.  The payloads are tiny (bucket totals
of the radix histograms, a few scalars), so latency, not bandwidth, is what matters.
"""

from __future__ import annotations

import numpy as np


class TorchComm:
    def __init__(self, dist, device=None):
        import torch

        self.torch = torch
        self.dist = dist
        self.rank = dist.get_rank()
        self.size = dist.get_world_size()
        self.device = device if device is not None else torch.device("cpu")

    def allgather(self, arr: np.ndarray) -> list[np.ndarray]:
        """All-gather equally shaped numpy arrays (histogram-merge payload)."""
        torch = self.torch
        arr = np.ascontiguousarray(arr)
        view = arr.view(np.uint8).reshape(-1)
        t = torch.from_numpy(view.copy()).to(self.device)
        outs = [torch.empty_like(t) for _ in range(self.size)]
        self.dist.all_gather(outs, t)
        return [o.cpu().numpy().view(arr.dtype).reshape(arr.shape) for o in outs]

    def allgather_object(self, obj):
        outs = [None] * self.size
        self.dist.all_gather_object(outs, obj)
        return outs

    def barrier(self):
        self.dist.barrier()

    # ---- device-pointer collectives (NCCL): the payload never leaves HBM, nothing blocks the host
    def _tensor(self, ptr: int, nbytes: int):
        key = (ptr, nbytes)
        cache = self.__dict__.setdefault("_tensors", {})
        t = cache.get(key)
        if t is None:
            holder = type("CudaPtr", (), {})()
            holder.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3}
            t = cache[key] = self.torch.as_tensor(holder, device=self.device)
        return t

    def _on(self, stream_handle: int):
        """Order the collective behind (and the following work after) the given cudaStream_t --
        the libcsgpu context's stream -- whatever torch's current stream happens to be."""
        torch = self.torch
        if not stream_handle:
            return torch.cuda.stream(torch.cuda.current_stream(self.device))
        cache = self.__dict__.setdefault("_streams", {})
        ext = cache.get(stream_handle)
        if ext is None:
            ext = cache[stream_handle] = torch.cuda.ExternalStream(stream_handle, device=self.device)
        return torch.cuda.stream(ext)

    def stream_scope(self, stream_handle: int):
        """Context manager: inside it the device collectives may be called with
        ``stream_handle=None`` (no per-call stream switch -- a dozen tiny collectives per step)."""
        return self._on(stream_handle)

    def allgather_dev(self, src_ptr: int, dst_ptr: int, nbytes: int, stream_handle: int | None = 0):
        """dst[rank][nbytes] <- every rank's src[nbytes], enqueued behind the context stream's work."""
        dst, src = self._tensor(dst_ptr, nbytes * self.size), self._tensor(src_ptr, nbytes)
        if stream_handle is None:
            self.dist.all_gather_into_tensor(dst, src)
            return
        with self._on(stream_handle):
            self.dist.all_gather_into_tensor(dst, src)

    def allreduce_max_dev(self, ptr: int, count: int, kind: str, stream_handle: int | None = 0):
        key = (ptr, count, kind)
        cache = self.__dict__.setdefault("_typed", {})
        t = cache.get(key)
        if t is None:
            torch = self.torch
            dt = {"i8": torch.int64, "f8": torch.float64, "i4": torch.int32}[kind]
            t = cache[key] = self._tensor(ptr, count * dt.itemsize).view(dt)
        if stream_handle is None:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            return
        with self._on(stream_handle):
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)

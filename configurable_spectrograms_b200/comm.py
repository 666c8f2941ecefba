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

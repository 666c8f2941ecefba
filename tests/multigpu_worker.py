"""torchrun worker for the multi-GPU parity test: every rank takes a contiguous block of the golden
tree's orbits; the pooled extrema exchanged over NCCL must equal the reference's JSON, and the
directory driver must write the same PNG tree as on one GPU."""

import json
import os
import sys


ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist

    from configurable_spectrograms_b200 import _lib, cdf_utils
    from configurable_spectrograms_b200.comm import TorchComm
    from configurable_spectrograms_b200.fast.batch_directory import FAST_plot_spectrograms_directory
    from configurable_spectrograms_b200.fast.extrema import extrema_from_shard
    from configurable_spectrograms_b200.fast.pipeline import ShardPlan
    from tests.helpers import load_json
    from tests.test_gpu_pipeline import ORDER, _tree

    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    rank, world = dist.get_rank(), dist.get_world_size()
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ctx = _lib.Context(local, stream=stream.cuda_stream)
    comm = TorchComm(dist, dev)
    tree = _tree()
    gold = load_json("extrema_tree.json")
    sequence = [(o, {i: True for i in dsets}) for o, dsets, _ in tree]
    per = (len(tree) + world - 1) // world
    lo, hi = rank * per, min(len(tree), (rank + 1) * per)
    for ys, zs, mins, p, key in (("linear", "log", False, 99.0, "batch_extrema"), ("linear", "linear", True, 95.0, "pool95_mins")):
        shard = ShardPlan(ctx, ys, zs, instrument_order=ORDER)
        shard.first_orbit_index = lo
        for o, dsets, lines in tree[lo:hi]:
            shard.add_orbit(o, dsets, lines)
        shard.upload()
        shard.collapse()
        state = extrema_from_shard(shard, sequence, ORDER, ys, zs, {}, max_percentile=p, compute_mins=mins, comm=comm)
        assert state == gold[key], (rank, key, state, gold[key])
    # the peer exchange on its own, through real CUDA IPC: payloads of several sizes, then mailboxes that
    # must grow (every rank unmaps its imports before any exporter frees) and further exchanges
    import numpy as np

    from configurable_spectrograms_b200.comm import IpcPeerExchange

    ex = IpcPeerExchange(comm, ctx)
    for slot, sizes in ((4096, (16, 4096, 48)), (1 << 20, (1 << 20, 64, 4096, 1 << 19, 32))):
        ex.ensure(slot)
        for k, nbytes in enumerate(sizes):
            mine = np.full(nbytes, (17 * rank + k) % 251, dtype=np.uint8)
            src = ctx.to_device(mine)
            gathered = ex.allgather(src.ptr, nbytes)
            host = np.empty(world * nbytes, np.uint8)
            ctx._check(ctx.lib.csg_d2h(ctx.handle, host.ctypes.data, gathered, world * nbytes))
            ctx.sync()
            want = np.concatenate([np.full(nbytes, (17 * r + k) % 251, dtype=np.uint8) for r in range(world)])
            assert np.array_equal(host, want), (rank, slot, nbytes)
    stats = ex.wait_stats(8)
    assert 0.0 <= stats["mean_us"] <= stats["max_us"] < 5e6, stats
    ex.close()  # collective: every rank unmaps its imports, barrier, then frees its own mailbox
    assert ex.handle is None
    dist.barrier()
    # the directory driver across ranks (shared working directory prepared by the parent test)
    work = sys.argv[1]
    os.chdir(work)
    cdf_utils.filtered_orbits_cache.clear()
    res = FAST_plot_spectrograms_directory(
        "./FAST_data", output_base="./FAST_plots/", y_scale="linear", z_scale="log", colormap="cividis",
        max_processing_percentile=99, max_workers=2, progress_json_path="./progress.json",
    )
    dist.barrier()
    if rank == 0:
        assert sorted((r["orbit"], r["status"]) for r in res) == [tuple(x) for x in gold["batch_status"]]
        pngs = []
        for dirpath, _dirs, fs in os.walk("./FAST_plots"):
            pngs += [os.path.relpath(os.path.join(dirpath, fn), "./FAST_plots") for fn in fs]
        assert sorted(pngs) == gold["batch_pngs"], sorted(pngs)
        assert json.load(open("./FAST_calculated_extrema.json")) == gold["batch_extrema"]
        print(f"MULTIGPU_OK world={world}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

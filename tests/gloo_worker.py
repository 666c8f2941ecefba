"""World-size-2 worker (gloo, CPU): the multi-rank host logic of the pooled-extrema selection.

Every rank generates the same synthetic sequence of collapsed matrices, keeps its contiguous
block, and resolves the prefix-pool percentiles through ``pool_select.prefix_percentiles``
over ``TorchComm`` (gloo) with the numpy kernel stand-in; the answers must equal numpy's
``nanpercentile`` of the concatenated pools (what ``fast/extrema.py:280-300`` computes).
usage: gloo_worker.py RANK WORLD PORT
"""

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def make_sequence(dtype, n_inst=3, n_files=7, seed=11):
    rng = np.random.default_rng(seed)
    seq = []
    for k in range(n_files):
        for i in range(n_inst):
            if rng.random() < 0.15:
                continue  # instrument missing for this orbit
            T, E = int(rng.integers(3, 40)), 12
            scale = 50.0 if k == 1 else 1.0  # an early "storm" file: running max != final percentile
            m = (rng.gamma(2.0, 3.0, (T, E)) * scale).astype(dtype)
            m[rng.random((T, E)) < 0.2] = 0
            m[rng.random((T, E)) < 0.05] = np.nan
            if k == 3:
                m[0, 0] = np.inf
            if k == 4 and i == 0:
                m[:] = 0  # a file without positives repeats the candidate
            seq.append((k, i, m))
    return seq


def main():
    rank, world, port = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=port, RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist

    from configurable_spectrograms_b200._lib import POOL_ITEM
    from configurable_spectrograms_b200.comm import TorchComm
    from configurable_spectrograms_b200.fast.extrema import energy_candidates
    from configurable_spectrograms_b200.pool_select import prefix_percentiles
    from oracle.pool_backend import NumpyPoolBackend

    dist.init_process_group("gloo", rank=rank, world_size=world)
    comm = TorchComm(dist)
    for dtype in (np.float32, np.float64):
        n_inst, n_files = 3, 7
        seq = make_sequence(dtype, n_inst, n_files)
        per = (n_files + world - 1) // world
        lo, hi = rank * per, min(n_files, (rank + 1) * per)
        mine = [(k, i, m) for k, i, m in seq if lo <= k < hi]
        flat, rows, pos_of = [], [], [0] * n_inst
        off = 0
        for k, i, m in mine:
            rows.append((off, m.shape[0], m.shape[1], i, pos_of[i]))
            pos_of[i] += 1
            flat.append(m.reshape(-1))
            off += m.size
        items = np.array(rows, dtype=POOL_ITEM) if rows else np.zeros(0, POOL_ITEM)
        mats = np.concatenate(flat) if flat else np.zeros(0, dtype)
        inst_len = np.array(pos_of, dtype=np.int32)
        requests = [{"inst": i, "p": 99.0, "mode": "running_max"} for i in range(n_inst)]
        requests += [{"inst": i, "p": 1, "mode": "last"} for i in range(n_inst)]
        requests += [{"inst": 0, "p": 50.0, "mode": "running_max"}]
        vals, counts, npos = prefix_percentiles(NumpyPoolBackend(mats), dtype, items, n_inst, inst_len, 12, requests, comm=comm)
        # ---- brute force over the global sequence (every rank can do it: same seed)
        for r, rq in enumerate(requests):
            pool, best, last = [], None, None
            for k, i, m in seq:
                if i != rq["inst"]:
                    continue
                with np.errstate(invalid="ignore"):
                    pos = m[np.isfinite(m) & (m > 0)]
                if pos.size:
                    pool.append(pos)
                if pool:
                    last = float(np.nanpercentile(np.concatenate(pool), rq["p"]))
                    best = last if best is None else max(best, last)
            want = best if rq["mode"] == "running_max" else last
            assert vals[r] == want, (rank, dtype, r, rq, vals[r], want)
        # ---- per-energy counts: gathered rows reproduce the single-process candidates
        all_counts = np.concatenate(comm.allgather(np.pad(counts, ((0, 32 - len(counts)), (0, 0)))))
        all_n = [len(c) for c in comm.allgather_object(counts)]
        energy = np.geomspace(4000.0, 5.0, 12)
        for i in range(n_inst):
            glob, single = [], []
            for rk in range(world):
                blk = [(k, ii, m) for k, ii, m in seq if rk * per <= k < min(n_files, (rk + 1) * per)]
                row = 0
                for k, ii, m in blk:
                    if ii == i:
                        glob.append(all_counts[rk * 32 + row])
                    row += 1
            for k, ii, m in seq:
                if ii == i:
                    with np.errstate(invalid="ignore"):
                        single.append((np.isfinite(m) & (m > 0)).sum(axis=0))
            assert sum(all_n) == len(seq)
            if single:
                assert np.array_equal(np.asarray(glob), np.asarray(single)), (rank, i)
                a = energy_candidates([energy] * len(single), np.asarray(glob))
                b = energy_candidates([energy] * len(single), np.asarray(single))
                assert a == b
    dist.barrier()
    if rank == 0:
        print(f"GLOO_OK world={world}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""Exact order statistics of every prefix pool of an instrument's file sequence.

The reference recomputes ``np.nanpercentile(np.concatenate(blocks so far), p)`` after
every (orbit, instrument) step and keeps a running maximum
(``fast/extrema.py:280-300``).  Here the pooled samples stay in HBM as the collapsed
total matrices; per-file radix-digit histograms are scanned along the file sequence so
one row describes one prefix pool, and the two neighbours of every prefix's percentile
are located digit by digit (``csrc/pool.cu``).  Prefixes that provably cannot hold the
running maximum are dropped between digits, so the refinement passes follow only a
handful of buckets.

Multi-GPU: ranks hold contiguous blocks of the ascending-orbit sequence.  Per digit one
all-gather of each rank's bucket totals gives every rank the counts held by lower
ranks (added by ``csg_pool_scan``); one all-gather of the surviving (instrument,
prefix) pairs and their lower bounds keeps the slot tables identical everywhere.
"""

from __future__ import annotations

import numpy as np

from . import _lib
from ._lib import POOL_ITEM, POOL_QUERY

FIRST_BITS = 11
NEXT_BITS = 10
MAX_SLOTS = 64


def key_bits(dtype) -> int:
    return 31 if np.dtype(dtype) == np.float32 else 63


def digit_plan(dtype) -> list[tuple[int, int]]:
    """[(shift, bits)] from the most significant digit down to shift 0."""
    total = key_bits(dtype)
    plan = [(total - FIRST_BITS, FIRST_BITS)]
    shift = total - FIRST_BITS
    while shift > 0:
        b = min(NEXT_BITS, shift)
        shift -= b
        plan.append((shift, b))
    return plan


def bits_to_value(bits: np.ndarray, dtype) -> np.ndarray:
    dt = np.dtype(dtype)
    if dt == np.float32:
        return bits.astype(np.uint32).view(np.float32)
    return bits.astype(np.uint64).view(np.float64)


def percentile_ranks(n: np.ndarray, p, dtype):
    """Vectorised numpy ``_get_indexes`` / ``_get_gamma`` in dtype D (n >= 1)."""
    D = np.dtype(dtype).type
    q = D(p) / D(100)
    nm1 = (n - 1).astype(dtype)
    v = nm1 * q
    above = v >= nm1
    below = v < 0
    fl = np.floor(v)
    lo = fl.astype(np.int64)
    hi = lo + 1
    gamma = (v - fl).astype(dtype)
    lo = np.where(above, n - 1, np.where(below, 0, lo))
    hi = np.where(above, n - 1, np.where(below, 0, hi))
    hi = np.minimum(hi, n - 1)
    lo = np.minimum(lo, n - 1)
    gamma = np.where(above | below, D(0), gamma).astype(dtype)
    return lo, hi, gamma


def lerp(a, b, g, dtype):
    """numpy ``_lerp`` rounded after every operation in D."""
    D = np.dtype(dtype).type
    with np.errstate(invalid="ignore", over="ignore"):
        a, b, g = D(a), D(b), D(g)
        d = D(b - a)
        r = D(a + D(d * g))
        if g >= 0.5:
            r = D(b - D(d * D(D(1) - g)))
    return r


class GpuPoolBackend:
    """The four pool kernels over a :class:`engine.Batch`'s sums buffer."""

    def __init__(self, batch):
        self.batch = batch
        self.ctx = batch.ctx
        self.d_items = None
        self.n_items = 0

    def set_items(self, items: np.ndarray, n_inst: int, max_pos: int, inst_len: np.ndarray, max_E: int):
        self.items = items
        self.n_items = len(items)
        self.n_inst, self.max_pos, self.max_E = n_inst, max(max_pos, 1), max(max_E, 1)
        self.d_items = self.ctx.to_device(items) if len(items) else None
        self.d_inst_len = self.ctx.to_device(np.ascontiguousarray(inst_len, dtype=np.int32))

    def hist_first(self, bits: int):
        nb = 1 << bits
        self.n_slots, self.bits = 1, bits
        self.d_hist = self.ctx.alloc(self.n_inst * self.max_pos * nb * 4)
        self.d_hist.zero()
        d_counts = self.ctx.alloc(max(self.n_items, 1) * self.max_E * 4)
        d_npos = self.ctx.alloc(max(self.n_items, 1) * 4)
        if self.n_items:
            self.ctx._check(
                self.ctx.lib.csg_pool_hist_first(
                    self.ctx.handle, self.batch.d_sums.ptr, self.batch.code, self.d_items.ptr, self.n_items,
                    self.max_pos, bits, self.max_E, self.d_hist.ptr, d_counts.ptr, d_npos.ptr,
                )
            )
        counts = d_counts.download(np.int32, self.n_items * self.max_E, sync=False).reshape(self.n_items, self.max_E)
        npos = d_npos.download(np.int32, self.n_items)
        return counts, npos

    def hist_refine(self, slot_prefix: np.ndarray, prefix_shift: int, shift: int, bits: int):
        n_slots = slot_prefix.shape[1]
        nb = 1 << bits
        self.n_slots, self.bits = n_slots, bits
        self.d_hist = self.ctx.alloc(self.n_inst * self.max_pos * n_slots * nb * 4)
        self.d_hist.zero()
        d_pref = self.ctx.to_device(np.ascontiguousarray(slot_prefix, dtype=np.uint64))
        if self.n_items:
            self.ctx._check(
                self.ctx.lib.csg_pool_hist_refine(
                    self.ctx.handle, self.batch.d_sums.ptr, self.batch.code, self.d_items.ptr, self.n_items,
                    self.max_pos, n_slots, d_pref.ptr, prefix_shift, shift, bits, self.d_hist.ptr,
                )
            )

    def scan(self, want_totals: bool):
        """Inclusive scan along the file sequence; returns this rank's bucket totals when asked."""
        nb = 1 << self.bits
        n = self.n_inst * self.n_slots * nb
        d_tot = self.ctx.alloc(n * 4) if want_totals else None
        self.ctx._check(
            self.ctx.lib.csg_pool_scan(
                self.ctx.handle, self.d_hist.ptr, self.n_inst, self.max_pos, self.d_inst_len.ptr, self.n_slots,
                self.bits, d_tot.ptr if d_tot is not None else None,
            )
        )
        if want_totals:
            return d_tot.download(np.uint32, n).reshape(self.n_inst, self.n_slots, nb)
        return None

    def locate(self, queries: np.ndarray, base: np.ndarray | None) -> np.ndarray:
        if len(queries) == 0:
            return queries
        d_q = self.ctx.to_device(queries)
        d_base = self.ctx.to_device(np.ascontiguousarray(base, dtype=np.uint32)) if base is not None else None
        self.ctx._check(
            self.ctx.lib.csg_pool_locate(
                self.ctx.handle, self.d_hist.ptr, self.max_pos, self.n_slots, self.bits,
                d_base.ptr if d_base is not None else None, d_q.ptr, len(queries),
            )
        )
        return d_q.download(POOL_QUERY, len(queries))


class SingleRank:
    rank, size = 0, 1

    def allgather(self, arr: np.ndarray) -> list[np.ndarray]:
        return [arr]

    def allgather_object(self, obj):
        return [obj]


def prefix_percentiles(backend, dtype, items: np.ndarray, n_inst: int, inst_len: np.ndarray, max_E: int,
                       requests: list[dict], comm=None):
    """Resolve percentile requests over prefix pools.

    ``items`` (POOL_ITEM) are this rank's files, ``pos`` = local position in the
    instrument's sequence.  ``requests`` entries::

        {"inst": i, "p": percentile, "mode": "running_max" | "last"}

    ``running_max`` -> max over every prefix of nanpercentile(prefix pool, p) (NaN-free: the
    pool holds finite positives only); ``last`` -> the percentile of the whole pool.
    Returns ``(values, counts, npos)``: one float per request (``None`` when the
    instrument's pool is empty everywhere), the per-item per-energy positive counts and
    per-item positive totals (host side of ``fast/extrema.py:260-264``).
    """
    comm = comm or SingleRank()
    D = np.dtype(dtype)
    plan = digit_plan(D)
    max_pos = int(inst_len.max()) if len(inst_len) else 0
    backend.set_items(items, n_inst, max_pos, inst_len, max_E)

    # ---- digit 0: histograms, positive counts
    shift0, bits0 = plan[0]
    counts, npos = backend.hist_first(bits0)
    # pool size after each local prefix, plus what lower ranks hold
    local_tot = np.zeros(n_inst, dtype=np.int64)
    n_after = np.zeros((n_inst, max(max_pos, 1)), dtype=np.int64)
    for i in range(n_inst):
        sel = np.flatnonzero(items["inst"] == i)
        order = sel[np.argsort(items["pos"][sel])]
        c = np.cumsum(npos[order].astype(np.int64))
        n_after[i, : len(c)] = c
        local_tot[i] = c[-1] if len(c) else 0
    all_tot = comm.allgather(local_tot)
    below = np.sum(all_tot[: comm.rank], axis=0).astype(np.int64) if comm.rank > 0 else np.zeros(n_inst, np.int64)
    grand = np.sum(all_tot, axis=0).astype(np.int64)
    n_after += below[:, None]

    # ---- queries: (request, pos) pairs -> two rank targets each
    q_req, q_pos, q_lo, q_hi, q_gamma = [], [], [], [], []
    for r, req in enumerate(requests):
        i = req["inst"]
        L = int(inst_len[i])
        if req["mode"] == "last":
            # the pool after the globally last file: owned by the highest rank holding files of inst
            holders = [rk for rk in range(comm.size) if all_tot[rk][i] > 0]
            if not holders or holders[-1] != comm.rank or L == 0:
                continue
            # last local position with a non-empty cumulative pool
            pos_list = [L - 1]
        else:
            pos_list = list(range(L))
        for k in pos_list:
            n = int(n_after[i, k])
            if n <= 0:
                continue
            if req["mode"] == "running_max" and k > 0 and n == int(n_after[i, k - 1]):
                continue  # file added no positive sample: same pool, same candidate
            lo, hi, g = percentile_ranks(np.array([n]), req["p"], D)
            q_req.append(r), q_pos.append(k), q_lo.append(int(lo[0])), q_hi.append(int(hi[0])), q_gamma.append(g[0])
    nq = len(q_req)
    q_req = np.array(q_req, dtype=np.int64)
    q_pos = np.array(q_pos, dtype=np.int64)
    q_inst = np.array([requests[r]["inst"] for r in q_req], dtype=np.int64) if nq else np.zeros(0, np.int64)
    # targets: index 2*j (lo) and 2*j+1 (hi)
    t_rank = np.empty(2 * nq, dtype=np.int64)
    t_rank[0::2], t_rank[1::2] = q_lo, q_hi
    t_prefix = np.zeros(2 * nq, dtype=np.uint64)
    t_inst = np.repeat(q_inst, 2)
    t_pos = np.repeat(q_pos, 2)
    active = np.ones(nq, dtype=bool)

    def exchange_base(totals):
        """counts held by lower ranks for every (inst, slot, bin)"""
        if comm.size == 1 or totals is None:
            return None
        allt = comm.allgather(totals)
        if comm.rank == 0:
            return None
        return np.sum(np.stack(allt[: comm.rank]).astype(np.uint64), axis=0).astype(np.uint32)

    prev_shift = None
    for level, (shift, bits) in enumerate(plan):
        if level == 0:
            base = exchange_base(backend.scan(comm.size > 1))
            t_slot = np.zeros(2 * nq, dtype=np.int64)
            pending = np.repeat(active, 2)
            _locate(backend, base, t_inst, t_pos, t_slot, t_rank, t_prefix, pending, bits)
        else:
            # distinct (inst, prefix) among active targets, agreed across ranks
            act_t = np.repeat(active, 2)
            mine = sorted({(int(i), int(p)) for i, p in zip(t_inst[act_t], t_prefix[act_t])})
            union = sorted(set().union(*[set(x) for x in comm.allgather_object(mine)]))
            todo = {i: [p for (ii, p) in union if ii == i] for i in range(n_inst)}
            done_t = ~act_t  # inactive targets need no refinement
            while True:
                n_slots = min(MAX_SLOTS, max((len(v) for v in todo.values()), default=0))
                if n_slots == 0:
                    break
                table = np.full((n_inst, n_slots), np.iinfo(np.uint64).max, dtype=np.uint64)
                taken = {}
                for i in range(n_inst):
                    chunk = todo[i][:n_slots]
                    todo[i] = todo[i][n_slots:]
                    table[i, : len(chunk)] = np.array(chunk, dtype=np.uint64)
                    taken[i] = {p: s for s, p in enumerate(chunk)}
                backend.hist_refine(table, prev_shift, shift, bits)
                base = exchange_base(backend.scan(comm.size > 1))
                t_slot = np.full(2 * nq, -1, dtype=np.int64)
                for j in np.flatnonzero(~done_t):
                    s = taken[int(t_inst[j])].get(int(t_prefix[j]))
                    if s is not None:
                        t_slot[j] = s
                pending = t_slot >= 0
                _locate(backend, base, t_inst, t_pos, t_slot, t_rank, t_prefix, pending, bits)
                done_t |= pending
        prev_shift = shift
        # ---- prune prefixes that cannot hold the running maximum
        lo_bound = bits_to_value(t_prefix[0::2] << np.uint64(shift), D).astype(np.float64)
        hi_bound = bits_to_value(((t_prefix[1::2] + np.uint64(1)) << np.uint64(shift)) - np.uint64(1), D).astype(np.float64)
        best_local = {}
        for r, req in enumerate(requests):
            if req["mode"] != "running_max":
                continue
            sel = active & (q_req == r)
            best_local[r] = float(lo_bound[sel].max()) if sel.any() else -np.inf
        best = {}
        for d in comm.allgather_object(best_local):
            for r, v in d.items():
                best[r] = max(best.get(r, -np.inf), v)
        for r, req in enumerate(requests):
            if req["mode"] != "running_max":
                continue
            sel = active & (q_req == r)
            active[sel & (hi_bound < best[r])] = False

    # ---- exact neighbours -> numpy's lerp -> per-request reduction
    vals = bits_to_value(t_prefix, D)
    local = {}
    for j in np.flatnonzero(active):
        r = int(q_req[j])
        v = float(lerp(vals[2 * j], vals[2 * j + 1], q_gamma[j], D))
        if requests[r]["mode"] == "running_max":
            local[r] = max(local.get(r, -np.inf), v)
        else:
            local[r] = v
    merged: dict[int, float] = {}
    for d in comm.allgather_object(local):
        for r, v in d.items():
            merged[r] = max(merged[r], v) if (r in merged and requests[r]["mode"] == "running_max") else v
    out = []
    for r, req in enumerate(requests):
        out.append(merged.get(r) if grand[req["inst"]] > 0 else None)
    return out, counts, npos


def _locate(backend, base, t_inst, t_pos, t_slot, t_rank, t_prefix, pending, bits):
    idx = np.flatnonzero(pending)
    if len(idx) == 0:
        return
    q = np.zeros(len(idx), dtype=POOL_QUERY)
    q["inst"], q["pos"], q["slot"], q["rank"] = t_inst[idx], t_pos[idx], t_slot[idx], t_rank[idx]
    q["bin"] = -1
    res = backend.locate(q, base)
    if np.any(res["bin"] < 0):
        raise _lib.CsgError("pool_locate: a rank fell outside its bucket (histogram / scan mismatch)")
    t_prefix[idx] = (t_prefix[idx] << np.uint64(bits)) | res["bin"].astype(np.uint64)
    t_rank[idx] = res["rank"]

"""TEST INFRASTRUCTURE -- numpy stand-in for the four pool kernels (``csrc/pool.cu``).

Implements the backend interface ``pool_select.prefix_percentiles`` drives
(``set_items / hist_first / hist_refine / scan / locate``) on the host, so the multi-rank
host logic of the pooled-extrema selection (reference ``fast/extrema.py:259-300``: positive
pool, per-energy counts, percentile of every prefix pool) can be exercised with ``gloo`` and
no GPU.  Only ``tests/`` imports this module; the product path never does.
"""

from __future__ import annotations

import numpy as np


class NumpyPoolBackend:
    def __init__(self, mats: np.ndarray):
        """``mats``: flat array (float32 | float64) holding every collapsed total matrix."""
        self.mats = mats
        self.dtype = mats.dtype
        self.utype = np.uint32 if self.dtype == np.float32 else np.uint64

    def set_items(self, items, n_inst, max_pos, inst_len, max_E):
        self.items, self.n_inst = items, n_inst
        self.max_pos, self.max_E = max(max_pos, 1), max(max_E, 1)
        self.inst_len = np.asarray(inst_len)
        self._keys = []
        for it in items:
            m = self.mats[it["mat_off"] : it["mat_off"] + it["T"] * it["E"]].reshape(it["T"], it["E"])
            with np.errstate(invalid="ignore"):
                pos = np.isfinite(m) & (m > 0)  # fast/extrema.py:260
            self._keys.append((m, pos, m[pos].view(self.utype).astype(np.uint64)))

    def hist_first(self, bits):
        nb = 1 << bits
        total_bits = 31 if self.dtype == np.float32 else 63
        shift = total_bits - bits
        self.n_slots, self.bits = 1, bits
        self.hist = np.zeros((self.n_inst, self.max_pos, 1, nb), dtype=np.uint32)
        counts = np.zeros((len(self.items), self.max_E), dtype=np.int32)
        npos = np.zeros(len(self.items), dtype=np.int32)
        for k, (it, (m, pos, keys)) in enumerate(zip(self.items, self._keys)):
            counts[k, : it["E"]] = pos.sum(axis=0)  # fast/extrema.py:261-264
            npos[k] = pos.sum()
            self.hist[it["inst"], it["pos"], 0] += np.bincount((keys >> np.uint64(shift)).astype(np.int64), minlength=nb).astype(np.uint32)
        return counts, npos

    def hist_refine(self, slot_prefix, prefix_shift, shift, bits):
        nb = 1 << bits
        n_slots = slot_prefix.shape[1]
        self.n_slots, self.bits = n_slots, bits
        self.hist = np.zeros((self.n_inst, self.max_pos, n_slots, nb), dtype=np.uint32)
        for it, (m, pos, keys) in zip(self.items, self._keys):
            pref = keys >> np.uint64(prefix_shift)
            digit = ((keys >> np.uint64(shift)) & np.uint64(nb - 1)).astype(np.int64)
            for s in range(n_slots):
                want = slot_prefix[it["inst"], s]
                if want == np.iinfo(np.uint64).max:
                    continue
                sel = pref == want
                if sel.any():
                    self.hist[it["inst"], it["pos"], s] += np.bincount(digit[sel], minlength=nb).astype(np.uint32)

    def scan(self, want_totals):
        for i in range(self.n_inst):
            L = int(self.inst_len[i])
            if L:
                self.hist[i, :L] = np.cumsum(self.hist[i, :L].astype(np.uint64), axis=0).astype(np.uint32)
        if not want_totals:
            return None
        tot = np.zeros((self.n_inst, self.n_slots, 1 << self.bits), dtype=np.uint32)
        for i in range(self.n_inst):
            L = int(self.inst_len[i])
            if L:
                tot[i] = self.hist[i, L - 1]
        return tot

    def locate(self, queries, base):
        out = queries.copy()
        for q in out:
            row = self.hist[q["inst"], q["pos"], q["slot"]].astype(np.int64)
            if base is not None:
                row = row + base[q["inst"], q["slot"]].astype(np.int64)
            cum = np.cumsum(row)
            b = int(np.searchsorted(cum, q["rank"], side="right"))
            if b >= len(row):
                q["bin"] = -1
                continue
            q["bin"] = b
            q["rank"] = q["rank"] - (cum[b] - row[b])
            q["row_total"] = cum[-1]
        return out

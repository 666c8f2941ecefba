"""Test infrastructure: a minimal CDF v3 *writer* (zVariables only), the counterpart of the native
reader in ``csrc/cdf.cpp``.  Written from the published CDF 3.x internal format (64-bit offsets,
big-endian internal records): CDR 1, GDR 2, VXR 6, VVR 7, zVDR 8, CCR 10, CPR 11, CVVR 13.

cdflib and real FAST files are not available offline, so reader and writer pin each other, not
NASA's library -- ``csrc/cdf.cpp`` says "parity unpinned" for that reason.  The writer deliberately
exercises what real files do: network (big-endian) and IBMPC (little-endian) encodings, gzip-
compressed variables split into several CVVRs, multi-entry and chained VXRs, a second VXR level,
never-written (sparse) records with and without a pad value, whole-file gzip compression.
"""

from __future__ import annotations

import struct
import zlib

import numpy as np

CDF_TYPES = {np.dtype("float32"): 44, np.dtype("float64"): 45, np.dtype("int32"): 4, np.dtype("int64"): 8,
             np.dtype("uint8"): 11, np.dtype("int16"): 2}
NETWORK, IBMPC = 1, 6


def _rec(rtype: int, body: bytes) -> bytes:
    return struct.pack(">qi", 12 + len(body), rtype) + body


class _Image:
    def __init__(self):
        self.buf = bytearray()

    def tell(self):
        return len(self.buf)

    def put(self, data: bytes) -> int:
        off = len(self.buf)
        self.buf += data
        return off

    def patch64(self, at: int, value: int):
        self.buf[at : at + 8] = struct.pack(">q", value)


def write_cdf(path, variables, encoding=IBMPC, file_gzip=False):
    """``variables``: list of dicts ``{"name", "data" (n_rec, *dims) ndarray, "gzip": level|None,
    "records_per_block": int, "sparse": set of record numbers NOT written, "pad": scalar|None,
    "rec_vary": bool, "two_level": bool}``."""
    big = encoding == NETWORK
    img = _Image()
    img.put(struct.pack(">II", 0xCDF30001, 0x0000FFFF))
    # ---- CDR (GDR offset patched later)
    copyright_ = b"CDF test file written by tests/cdf_writer.py".ljust(256, b"\0")
    cdr = struct.pack(">qiiiiiiiii", 0, 3, 9, encoding, 0b11, 0, 0, 0, 0, 0) + copyright_
    cdr_off = img.put(_rec(1, cdr))
    # ---- GDR (heads patched later)
    gdr_body = struct.pack(">qqqqiiiiiqiii", 0, 0, 0, 0, 0, 0, -1, 0, len(variables), 0, 0, 20170101, 0)
    gdr_off = img.put(_rec(2, gdr_body))
    img.patch64(cdr_off + 12, gdr_off)
    prev_next_at = gdr_off + 20  # zVDRhead
    for num, var in enumerate(variables):
        data = np.asarray(var["data"])
        n_rec = data.shape[0]
        dims = data.shape[1:]
        dt = data.dtype
        raw = data.astype(dt.newbyteorder(">" if big else "<"), copy=False)
        sparse = set(var.get("sparse") or ())
        gz = var.get("gzip")
        per_block = int(var.get("records_per_block") or max(1, n_rec))
        # ---- data records first: VVR / CVVR per run of written records, at most per_block records each
        entries = []
        r = 0
        while r < n_rec:
            if r in sparse:
                r += 1
                continue
            end = r
            while end + 1 < n_rec and (end + 1) not in sparse and end + 1 - r < per_block:
                end += 1
            payload = raw[r : end + 1].tobytes()
            if gz is not None:
                comp = zlib.compress(payload, gz)
                off = img.put(_rec(13, struct.pack(">iq", 0, len(comp)) + comp))
            else:
                off = img.put(_rec(7, payload))
            entries.append((r, end, off))
            r = end + 1
        # ---- index: VXRs of at most 3 entries, chained; optionally under a top-level VXR
        def vxr(ents, n_slots=None):
            n_slots = n_slots or len(ents)
            first = [e[0] for e in ents] + [-1] * (n_slots - len(ents))
            last = [e[1] for e in ents] + [-1] * (n_slots - len(ents))
            offs = [e[2] for e in ents] + [-1] * (n_slots - len(ents))
            body = struct.pack(">qii", 0, n_slots, len(ents)) + struct.pack(f">{n_slots}i", *first) + \
                struct.pack(f">{n_slots}i", *last) + struct.pack(f">{n_slots}q", *offs)
            return img.put(_rec(6, body))

        leaves = []
        for k in range(0, len(entries), 3):
            chunk = entries[k : k + 3]
            off = vxr(chunk, n_slots=4)
            leaves.append((chunk[0][0], chunk[-1][1], off))
        vxr_head = vxr_tail = 0
        if leaves:
            if var.get("two_level"):
                vxr_head = vxr_tail = vxr(leaves)
            else:
                for a, b in zip(leaves[:-1], leaves[1:]):
                    img.patch64(a[2] + 12, b[2])  # VXRnext
                vxr_head, vxr_tail = leaves[0][2], leaves[-1][2]
        cpr_off = -1
        if gz is not None:
            cpr_off = img.put(_rec(11, struct.pack(">iiii", 5, 0, 1, gz)))
        flags = (1 if var.get("rec_vary", True) else 0) | (2 if var.get("pad") is not None else 0) | (4 if gz is not None else 0)
        name = var["name"].encode().ljust(256, b"\0")
        max_rec = n_rec - 1
        body = struct.pack(">qiiqqiiiiiiiqi", 0, CDF_TYPES[dt], max_rec, vxr_head, vxr_tail, flags, 0, 0, 0, 0, 1, num, cpr_off,
                           per_block) + name
        body += struct.pack(">i", len(dims)) + struct.pack(f">{len(dims)}i", *dims) + struct.pack(f">{len(dims)}i", *([-1] * len(dims)))
        if var.get("pad") is not None:
            body += np.asarray(var["pad"], dtype=dt.newbyteorder(">" if big else "<")).tobytes()
        vdr_off = img.put(_rec(8, body))
        img.patch64(prev_next_at, vdr_off)
        prev_next_at = vdr_off + 12  # VDRnext
    img.patch64(gdr_off + 36, img.tell())  # eof
    blob = bytes(img.buf)
    if file_gzip:
        comp = zlib.compress(blob[8:], 6)
        ccr_size = 32 + len(comp)
        cpr_at = 8 + ccr_size
        ccr = struct.pack(">qiqqi", ccr_size, 10, cpr_at, len(blob) - 8, 0) + comp
        cpr = _rec(11, struct.pack(">iiii", 5, 0, 1, 6))
        blob = struct.pack(">II", 0xCDF30001, 0xCCCC0001) + ccr + cpr
    with open(path, "wb") as f:
        f.write(blob)


def write_fast_cdf(path, arrays, encoding=IBMPC, gzip=None, records_per_block=64, file_gzip=False):
    """The four FAST ESA variables of ``synth.make_file_arrays`` as a CDF file.  Like the real files,
    ``energy`` holds a few records and ``pitch_angle`` one record per time step (``FAST CDF variables.txt``)
    although only record 0 of either is ever used (``CS/cdf_utils.py:252-253``)."""
    T = len(arrays["time_unix"])
    energy = np.repeat(np.asarray(arrays["energy"], dtype=np.float32)[:1], 2, axis=0)
    pa = np.repeat(np.asarray(arrays["pitch_angle"], dtype=np.float32)[:1], max(T, 1), axis=0)
    write_cdf(path, [
        {"name": "time_unix", "data": np.asarray(arrays["time_unix"], dtype=np.float64), "gzip": gzip, "records_per_block": 256},
        {"name": "data", "data": np.asarray(arrays["data"]), "gzip": gzip, "records_per_block": records_per_block},
        {"name": "energy", "data": energy, "gzip": gzip},
        {"name": "pitch_angle", "data": pa, "gzip": gzip, "records_per_block": records_per_block},
    ], encoding=encoding, file_gzip=file_gzip)

"""PNG hand-off for the RGBA rasters (the on-disk product of the reference:
``fast/process_orbit.py:98-117``, ``generic_batch.py:108-113``).

Filter 0 + zlib DEFLATE on the host; ``zlib.compress`` releases the GIL, so a batch of figures
is encoded on a thread pool while the GPU works on the next shard.  A GPU DEFLATE stage is the
next item on the scope list.
"""

from __future__ import annotations

import struct
import zlib
from concurrent.futures import ThreadPoolExecutor

import numpy as np

_SIGNATURE = b"\x89PNG\r\n\x1a\n"


def _chunk(tag: bytes, payload: bytes) -> bytes:
    return struct.pack(">I", len(payload)) + tag + payload + struct.pack(">I", zlib.crc32(tag + payload) & 0xFFFFFFFF)


def encode_rgba(image: np.ndarray, compress_level: int = 6) -> bytes:
    """(H, W, 4) uint8 -> PNG bytes (8-bit RGBA, no interlace, filter type 0 on every row)."""
    img = np.ascontiguousarray(image, dtype=np.uint8)
    if img.ndim != 3 or img.shape[2] != 4:
        raise ValueError(f"expected an (H, W, 4) uint8 image, got {img.shape}")
    h, w, _ = img.shape
    raw = np.zeros((h, 1 + 4 * w), dtype=np.uint8)  # leading filter byte 0 per scanline
    raw[:, 1:] = img.reshape(h, 4 * w)
    ihdr = struct.pack(">IIBBBBB", w, h, 8, 6, 0, 0, 0)
    return _SIGNATURE + _chunk(b"IHDR", ihdr) + _chunk(b"IDAT", zlib.compress(raw.tobytes(), compress_level)) + _chunk(b"IEND", b"")


def write_rgba(path, image: np.ndarray, compress_level: int = 6) -> None:
    data = encode_rgba(image, compress_level)
    with open(path, "wb") as f:
        f.write(data)


def decode_rgba(data: bytes) -> np.ndarray:
    """Inverse of :func:`encode_rgba` (filter-0 RGBA only) -- used by the round-trip tests."""
    if data[:8] != _SIGNATURE:
        raise ValueError("not a PNG")
    pos, idat, shape = 8, b"", None
    while pos < len(data):
        (n,) = struct.unpack(">I", data[pos : pos + 4])
        tag, payload = data[pos + 4 : pos + 8], data[pos + 8 : pos + 8 + n]
        (crc,) = struct.unpack(">I", data[pos + 8 + n : pos + 12 + n])
        if zlib.crc32(tag + payload) & 0xFFFFFFFF != crc:
            raise ValueError("PNG chunk CRC mismatch")
        if tag == b"IHDR":
            w, h, depth, ctype = struct.unpack(">IIBB", payload[:10])
            if (depth, ctype) != (8, 6):
                raise ValueError("only 8-bit RGBA is supported")
            shape = (h, w)
        elif tag == b"IDAT":
            idat += payload
        pos += 12 + n
    h, w = shape
    raw = np.frombuffer(zlib.decompress(idat), dtype=np.uint8).reshape(h, 1 + 4 * w)
    if raw[:, 0].any():
        raise ValueError("only filter type 0 is supported")
    return raw[:, 1:].reshape(h, w, 4).copy()


def write_many(jobs, compress_level: int = 6, max_workers: int = 8) -> None:
    """``jobs``: iterable of (path, image).  Encodes and writes on a thread pool."""
    jobs = list(jobs)
    if not jobs:
        return
    with ThreadPoolExecutor(max_workers=max(1, min(max_workers, len(jobs)))) as pool:
        list(pool.map(lambda j: write_rgba(j[0], j[1], compress_level), jobs))

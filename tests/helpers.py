"""Shared test helpers: golden-fixture access and bit-level comparisons."""

import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_npz(name):
    with np.load(os.path.join(GOLDEN, name), allow_pickle=False) as z:
        return {k: z[k] for k in z.files}


def load_json(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def panels(store, prefix):
    out = []
    for i in range(int(store[f"{prefix}_n"])):
        mode, vmin, vmax = store[f"{prefix}_{i}_meta"]
        out.append(
            {
                "matrix": store[f"{prefix}_{i}_matrix"],
                "mode": "log" if mode == 1.0 else "linear",
                "vmin": float(vmin),
                "vmax": float(vmax),
                "extent": store[f"{prefix}_{i}_extent"],
            }
        )
    return out


def bits(a):
    a = np.ascontiguousarray(a)
    return a.view(np.uint32 if a.dtype == np.float32 else np.uint64)


def same_bits(a, b, zero_sign_insensitive=True):
    """Bit-exact float equality; NaN payloads equal as NaN, optional -0.0 == +0.0."""
    a = np.asarray(a)
    b = np.asarray(b)
    if a.shape != b.shape or a.dtype != b.dtype:
        return False
    nan = np.isnan(a) & np.isnan(b)
    eq = bits(a) == bits(b)
    if zero_sign_insensitive:
        eq |= (a == 0) & (b == 0)
    return bool(np.all(eq | nan))


def same_float(a, b):
    a = float(a)
    b = float(b)
    return a == b or (np.isnan(a) and np.isnan(b))


def dataset_from_arrays(arrays):
    """What ``load_fast_cdf_dataset`` returns (``cdf_utils.py:247-256``) for raw variables."""
    times = np.asarray(arrays["time_unix"])
    data = np.asarray(arrays["data"])
    energy_full = np.asarray(arrays["energy"])
    pa_full = np.asarray(arrays["pitch_angle"])
    energy = energy_full[0, 0, :] if energy_full.ndim == 3 else energy_full
    pa = pa_full[0, :, 0] if pa_full.ndim == 3 else pa_full
    if data.shape[1] == len(energy) and data.shape[2] == len(pa):
        data = np.transpose(data, (0, 2, 1))
    return {"times": times, "data": data, "energy": energy, "pitch_angle": pa}

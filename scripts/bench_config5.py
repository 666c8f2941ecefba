#!/usr/bin/env python
"""BASELINE config 5: the collapse alone on a generic stress cube (100 000 time x 96 x 64, float32,
2.46 GB): layout A = C-contiguous (T, P, E) collapsed over axis 1 (stream kernel), layout B = the
stored (T, E, P) view (pitch contiguous).  Prints achieved GB/s of algorithmic bytes (cube + sums)
per layout, CUDA-event timed, and checks a slab of the result against numpy bit for bit."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch

    from configurable_spectrograms_b200 import _lib
    from configurable_spectrograms_b200.engine import Batch

    T, P, E = 100_000, 96, 64
    dev = torch.device("cuda", 0)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ctx = _lib.Context(0, stream=stream.cuda_stream)
    gen = torch.Generator(device=dev)
    gen.manual_seed(5)
    cube = torch.rand((T, P, E), device=dev, generator=gen) * 1000.0
    cube[torch.rand((T, P, E), device=dev, generator=gen) < 0.005] = float("nan")
    out = {}
    for name, layout, shape in (("A_tpe", None, (T, P, E)), ("B_tep_view", _lib.LAYOUT_TEP, (T, P, E))):
        src = cube if layout is None else cube.transpose(1, 2).contiguous()  # (T, E, P) stored
        torch.cuda.synchronize()
        b = Batch(ctx, np.float32, n_groups=0)
        f = b.add_file(None, None, shape=shape, layout=layout, device_ptr=src.data_ptr())
        b.upload_cubes()
        for _ in range(3):
            b.collapse()
        ctx.sync()
        ms = []
        for _ in range(5):
            ctx.timer_start(0)
            b.collapse()
            ctx.timer_stop(0)
            ms.append(ctx.timer_ms(0))
        got = b.sums(f)[:2000]
        ref_src = cube[:2000].cpu().numpy()
        with np.errstate(invalid="ignore"):
            ref = np.nansum(ref_src if layout is None else np.ascontiguousarray(ref_src.transpose(0, 2, 1)).transpose(0, 2, 1), axis=1)
        exact = bool(np.array_equal(got.view(np.uint32), ref.view(np.uint32)))
        bytes_ = 4 * (T * P * E + T * E)
        out[name] = {"ms": float(np.mean(ms)), "gb_per_s": bytes_ / (np.mean(ms) * 1e-3) / 1e9, "bit_exact_vs_numpy_first_2000_rows": exact,
                     "kernel": "stream" if any(k[1] == _lib.K1_STREAM for k in b.d_files) else "generic/tep"}
        del b, src
    # ---- the same cubes through the public API: generic_batch_plot (reference generic_batch.py:15-129), the
    # items of a group sharing one K2a / K3 pass, figures at 150 dpi written as PNG files
    import shutil
    import tempfile
    import time

    from configurable_spectrograms_b200.generic_batch import generic_batch_plot

    n_items = int(os.environ.get("CONFIG5_ITEMS", "4"))
    host_cubes = []
    for k in range(n_items):  # host arrays, as a user's build_datasets_fn would hand them over
        c = cube.cpu().numpy().copy()
        c[k :: n_items] += np.float32(k)  # the items differ
        host_cubes.append(c)
    del cube
    torch.cuda.empty_cache()
    times = 946684800.0 + 0.02 * np.arange(T)
    energy = np.geomspace(4.0, 30000.0, E)[::-1].astype(np.float32)

    def build(item):
        return [{"x": times, "y": energy, "data": host_cubes[item], "label": f"cube {item}", "y_max": 4000}]

    work = tempfile.mkdtemp(prefix="config5_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    try:
        api = {}
        for label in ("cold", "warm"):
            shutil.rmtree(os.path.join(work, "out"), ignore_errors=True)
            t0 = time.perf_counter()
            res = generic_batch_plot(list(range(n_items)), os.path.join(work, "out"), build, y_scale="linear", z_scale="log",
                                     colormap="viridis", max_workers=4, progress_json_path=os.path.join(work, f"p_{label}.json"),
                                     install_signal_handlers=False)
            sec = time.perf_counter() - t0
            pngs = sum(len(fs) for _d, _s, fs in os.walk(os.path.join(work, "out")))
            api[label] = {"seconds": sec, "items_per_s": n_items / sec, "input_gb_per_s": n_items * 4 * T * P * E / sec / 1e9,
                          "statuses": sorted({st for _i, st in res}), "pngs": pngs}
        out["generic_batch_plot"] = {"items": n_items, "cube_gb": 4 * T * P * E / 1e9, **api,
                                     "note": "host cubes (pageable numpy, as the reference's build_datasets_fn returns them) -> "
                                             "H2D + K1 per item, one K2a + K3 + K4 pass per group, generic.png per item; wall clock"}
    finally:
        shutil.rmtree(work, ignore_errors=True)
    print(json.dumps({"workload": "config5 generic stress cube (100000, 96, 64) float32, 2.46 GB", **out}))


if __name__ == "__main__":
    main()

"""CPU: the native CDF v3 reader (``csrc/cdf.cpp`` through the C ABI) against files written by
``tests/cdf_writer.py`` -- encodings, gzip variables, index shapes, sparse records, whole-file
compression -- and ``load_fast_cdf_dataset`` on a real (synthetic-content) ``.cdf`` file."""

import numpy as np
import pytest

from tests import cdf_writer as W


def _vars(rng, n_rec=37):
    return [
        {"name": "time_unix", "data": 946684800.0 + 2.5 * np.arange(n_rec)},
        {"name": "data", "data": rng.gamma(2.0, 3.0, (n_rec, 5, 7)).astype(np.float32)},
        {"name": "counts", "data": rng.integers(-5, 1000, (n_rec, 3), dtype=np.int32)},
        {"name": "scalar64", "data": rng.integers(0, 2**40, (n_rec,), dtype=np.int64)},
    ]


@pytest.mark.parametrize("encoding", [W.NETWORK, W.IBMPC])
@pytest.mark.parametrize("gzip", [None, 6])
@pytest.mark.parametrize("file_gzip", [False, True])
def test_reader_round_trip(tmp_path, encoding, gzip, file_gzip):
    from configurable_spectrograms_b200.cdf_reader import CdfFile, is_cdf_v3

    rng = np.random.default_rng(5)
    variables = _vars(rng)
    for v in variables:
        v["gzip"] = gzip
        v["records_per_block"] = 8  # several VVR / CVVR per variable, chained VXRs of 3 entries
    variables[1]["two_level"] = True
    path = tmp_path / "t.cdf"
    W.write_cdf(path, variables, encoding=encoding, file_gzip=file_gzip)
    assert is_cdf_v3(str(path))
    with CdfFile(str(path)) as cdf:
        assert cdf.variables() == [v["name"] for v in variables]
        for v in variables:
            assert cdf.shape(v["name"]) == v["data"].shape
            got = cdf.read(v["name"])
            assert got.dtype == v["data"].dtype and np.array_equal(got, v["data"])
            # a window that starts and ends inside blocks
            part = cdf.read(v["name"], 5, 20)
            assert np.array_equal(part, v["data"][5:25])
        # decode into caller memory (the pinned-slot path)
        out = np.empty(37 * 5 * 7 + 11, dtype=np.float32)
        view = cdf.read("data", out=out)
        assert view.base is not None and np.array_equal(view, variables[1]["data"])
        with pytest.raises(KeyError):
            cdf.shape("nope")
        with pytest.raises(Exception):
            cdf.read("data", 30, 20)


def test_reader_sparse_records_and_pad(tmp_path):
    from configurable_spectrograms_b200.cdf_reader import CdfFile

    rng = np.random.default_rng(6)
    a = rng.normal(size=(20, 4)).astype(np.float32)
    b = rng.normal(size=(20, 4))
    path = tmp_path / "s.cdf"
    W.write_cdf(path, [
        {"name": "a", "data": a, "sparse": {3, 4, 11}, "pad": np.float32(-1e31), "records_per_block": 5},
        {"name": "b", "data": b, "sparse": {0, 19}, "gzip": 1, "records_per_block": 4},
    ], encoding=W.NETWORK)
    with CdfFile(str(path)) as cdf:
        want = a.copy()
        want[[3, 4, 11]] = np.float32(-1e31)
        assert np.array_equal(cdf.read("a"), want)
        want = b.copy()
        want[[0, 19]] = 0.0  # no pad value in the file: zeros
        assert np.array_equal(cdf.read("b"), want)


def test_reader_rejects_non_cdf(tmp_path):
    from configurable_spectrograms_b200 import _lib
    from configurable_spectrograms_b200.cdf_reader import CdfFile, is_cdf_v3

    p = tmp_path / "x.cdf"
    p.write_bytes(b"")
    assert not is_cdf_v3(str(p))
    with pytest.raises(_lib.CsgError):
        CdfFile(str(p))
    p.write_bytes(b"\xcd\xf3\x00\x01\x00\x00\xff\xff" + b"\0" * 100)
    with pytest.raises(_lib.CsgError):
        CdfFile(str(p))


@pytest.mark.parametrize("stored_layout", ["tpe", "tep"])
def test_load_fast_cdf_dataset_from_a_cdf_file(tmp_path, stored_layout):
    """``load_fast_cdf_dataset`` on a binary CDF equals the side-car path on the same arrays
    (``CS/cdf_utils.py:247-256``: 1-D bins from record 0, conditional transpose to a (T,P,E) view)."""
    from configurable_spectrograms_b200 import synth
    from configurable_spectrograms_b200.cdf_utils import load_fast_cdf_dataset
    from tests.helpers import dataset_from_arrays

    rng = np.random.default_rng(8)
    arrays = synth.make_file_arrays(rng, "ees", n_time=50, quirks=True, stored_layout=stored_layout)
    path = tmp_path / synth.fast_filename("ees", arrays["time_unix"][0], 777)
    W.write_fast_cdf(path, arrays, encoding=W.NETWORK, gzip=6, records_per_block=16)
    ds = load_fast_cdf_dataset(str(path))
    ref = dataset_from_arrays(arrays)
    for key in ("times", "energy", "pitch_angle"):
        assert np.array_equal(ds[key], ref[key], equal_nan=True), key
    assert ds["data"].shape == ref["data"].shape and ds["data"].dtype == ref["data"].dtype
    assert ds["data"].flags.c_contiguous == ref["data"].flags.c_contiguous  # the transposed view is kept a view
    assert np.array_equal(ds["data"], ref["data"], equal_nan=True)


@pytest.mark.parametrize("gzip", [None, 6])
def test_reader_against_cdflib_when_it_is_installed(tmp_path, gzip):
    """The native reader is "parity unpinned" because ``cdflib`` (the reference's reader,
    ``CS/cdf_utils.py:247-251``) is not installable offline.  The day it is, this test pins both halves by
    itself: a file ``cdflib`` WRITES must read the same through ``csrc/cdf.cpp`` as through ``cdflib``, and
    a file of ``tests/cdf_writer.py`` must read the same through ``cdflib`` (skipped until then)."""
    cdflib = pytest.importorskip("cdflib")
    if not hasattr(cdflib, "CDF") or "oracle" in (getattr(cdflib, "__file__", "") or ""):
        pytest.skip("a cdflib stand-in is on the path, not cdflib")
    from configurable_spectrograms_b200.cdf_reader import CdfFile

    rng = np.random.default_rng(11)
    variables = _vars(rng)
    for v in variables:
        v["gzip"] = gzip
        v["records_per_block"] = 8
    ours = tmp_path / "ours.cdf"
    W.write_cdf(ours, variables, encoding=W.IBMPC)
    theirs_reader = cdflib.CDF(str(ours))
    for v in variables:  # our writer -> cdflib
        assert np.array_equal(np.asarray(theirs_reader.varget(v["name"])), v["data"]), v["name"]
    writer = getattr(getattr(cdflib, "cdfwrite", None), "CDF", None)
    if writer is None:
        return
    theirs = tmp_path / "theirs.cdf"
    out = writer(str(theirs), cdf_spec={"Compressed": 6 if gzip else 0}, delete=True)
    for k, v in enumerate(variables):
        spec = {"Variable": v["name"], "Data_Type": W.CDF_TYPES[v["data"].dtype], "Num_Elements": 1, "Rec_Vary": True,
                "Var_Type": "zVariable", "Dim_Sizes": list(v["data"].shape[1:]), "Compress": 6 if gzip else 0}
        out.write_var(spec, var_data=v["data"])
    out.close()
    with CdfFile(str(theirs)) as cdf:  # cdflib's writer -> the native reader
        for v in variables:
            assert np.array_equal(cdf.read(v["name"]), v["data"]), v["name"]


_FUZZ = r'''
import os, sys, tempfile
import numpy as np
sys.path.insert(0, sys.argv[1])
from tests import cdf_writer as W
from configurable_spectrograms_b200.cdf_reader import CdfFile

seed = int(sys.argv[2])
rng = np.random.default_rng(seed)
variables = [
    {"name": "time_unix", "data": 946684800.0 + 2.5 * np.arange(37), "gzip": 6 if seed % 3 == 0 else None, "records_per_block": 5},
    {"name": "data", "data": rng.gamma(2.0, 3.0, (37, 5, 7)).astype(np.float32), "gzip": 6 if seed % 2 else None,
     "records_per_block": 4, "two_level": seed % 4 == 0, "sparse": {3, 4, 20} if seed % 5 == 0 else None,
     "pad": -1.0 if seed % 5 == 0 else None},
    {"name": "counts", "data": rng.integers(-5, 1000, (37, 3), dtype=np.int32)},
]
tmp = tempfile.mkdtemp()
good = os.path.join(tmp, "good.cdf")
W.write_cdf(good, variables, encoding=W.NETWORK if seed % 2 else W.IBMPC, file_gzip=seed % 7 == 0)
blob = bytearray(open(good, "rb").read())
clean = errors = 0
for trial in range(int(sys.argv[3])):
    b = bytearray(blob)
    mode = trial % 4
    if mode == 0:
        b = b[: int(rng.integers(0, len(b)))]
    elif mode == 1:
        for _ in range(int(rng.integers(1, 6))):
            b[int(rng.integers(8, len(b)))] = int(rng.integers(0, 256))
    elif mode == 2:
        at = int(rng.integers(8, len(b) - 8))
        b[at : at + 8] = rng.integers(0, 256, 8, dtype=np.uint8).tobytes()
    else:
        at = int(rng.integers(8, min(len(b) - 8, 900)))
        b[at : at + 4] = int(rng.integers(0, 1 << 31)).to_bytes(4, "big")
    path = os.path.join(tmp, "damaged.cdf")
    open(path, "wb").write(bytes(b))
    try:
        with CdfFile(path) as cdf:
            for name in cdf.variables():
                if np.prod(cdf.shape(name), dtype=np.float64) > 1e8:
                    raise MemoryError("implausible shape")
                cdf.read(name)
        clean += 1
    except Exception:
        errors += 1
print("FUZZ_DONE", clean, errors)
'''


@pytest.mark.parametrize("seed", [2, 5, 7, 12])
def test_reader_survives_damaged_files(seed):
    """Truncated files, flipped bytes, overwritten offsets and sizes: the reader answers with an error (or
    with data, when the damage missed what it reads) -- never with a crash, an exception across the C boundary
    (a damaged size once ended in ``std::bad_alloc`` -> abort) or an unbounded allocation.  Run in a child
    process so that a crash is a failed test."""
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", _FUZZ, root, str(seed), "240"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "FUZZ_DONE" in r.stdout, (r.returncode, r.stdout[-500:], r.stderr[-1500:])
    clean, errors = (int(x) for x in r.stdout.split("FUZZ_DONE")[1].split()[:2])
    assert clean + errors == 240 and errors > 0

"""Data ingest and predictions writer of the host layer (parsmurf_b200/host/ingest.cpp) against the reference's own
Importer::import (src/fileImport.cpp:5-100) and saveToFile (src/HyperSMURFUtils.cpp:66-86), called through oracle/_ref.  CPU only."""
import os

import numpy as np
import pytest

from oracle import pyoracle as O

needs_ref = pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref (the compiled reference) is not present")


def _write(tmp_path, X, y, f=None, sep="\t", fmt="%.17g", trailing_newline=True):
    lines = [sep.join(fmt % v for v in row) for row in X]
    (tmp_path / "data.txt").write_text("\n".join(lines) + ("\n" if trailing_newline else ""))
    (tmp_path / "labels.txt").write_text("\n".join(str(int(v)) for v in y) + "\n")
    if f is not None:
        (tmp_path / "folds.txt").write_text(" ".join(str(int(v)) for v in f) + "\n")
    return str(tmp_path / "data.txt"), str(tmp_path / "labels.txt"), (str(tmp_path / "folds.txt") if f is not None else None)


@needs_ref
@pytest.mark.parametrize("sep,fmt,nl", [("\t", "%.17g", True), (" ", "%.6e", True), ("  \t ", "%g", False), ("\t", "%+.9f", True)])
def test_import_equals_the_reference_importer(tmp_path, sep, fmt, nl):
    from parsmurf_b200 import host_api as H
    rng = np.random.default_rng(3)
    n, m = 700, 13
    X = rng.standard_normal((n, m)) * rng.choice([1e-8, 1.0, 1e6], size=(n, m))
    X[5, 3] = 0.0; X[6, 0] = -0.0; X[7, 1] = 123456789.0; X[8, 2] = 5e-324; X[9, 4] = 1.7976931348623157e308
    y = (rng.random(n) < 0.1).astype(np.uint32); y[0] = 3              # any label > 0 is a positive (src/fileImport.cpp:88)
    f = rng.integers(0, 7, n)
    paths = _write(tmp_path, X, y, f, sep=sep, fmt=fmt, trailing_newline=nl)
    rx, ry, rf, rnf = O.ref_import(*paths)
    for threads in (1, 4):
        got = H.import_files(*paths, threads=threads)
        assert got["m"] == m and got["nFolds"] == rnf == int(f.max()) + 1
        assert np.array_equal(got["x"].ravel().view(np.uint64), rx.view(np.uint64))     # every parsed double, bit for bit
        assert np.array_equal(got["y"], ry) and np.array_equal(got["f"], rf)
    assert np.array_equal(got["x"][:, m], np.where(y > 0, 1.0, 2.0))


@needs_ref
def test_import_large_file_uses_several_threads_and_matches(tmp_path):
    from parsmurf_b200 import host_api as H
    rng = np.random.default_rng(4)
    n, m = 40_000, 26                                                   # ~20 MB of text: the parser splits it between threads
    X = rng.standard_normal((n, m))
    y = (rng.random(n) < 0.01).astype(np.uint32)
    paths = _write(tmp_path, X, y, None)
    rx, ry, rf, _ = O.ref_import(*paths)
    got = H.import_files(*paths)
    assert np.array_equal(got["x"].ravel().view(np.uint64), rx.view(np.uint64)) and np.array_equal(got["y"], ry) and len(got["f"]) == 0
    assert np.array_equal(got["x"][:, :m].view(np.uint64), X.view(np.uint64))           # %.17g round-trips


def test_import_errors_and_binary_side_channel(tmp_path):
    from parsmurf_b200 import host_api as H
    from parsmurf_b200._capi import PsmError
    rng = np.random.default_rng(5)
    n, m = 300, 7
    X = rng.standard_normal((n, m)); y = (rng.random(n) < 0.2).astype(np.uint32); f = rng.integers(0, 4, n).astype(np.uint32)
    d, l, ff = _write(tmp_path, X, y, f)
    with pytest.raises(PsmError):
        H.import_files(str(tmp_path / "missing.txt"), l, ff)
    (tmp_path / "short_labels.txt").write_text("\n".join(str(int(v)) for v in y[:-1]) + "\n")
    with pytest.raises(PsmError):                                        # the reference only warns and reads out of bounds
        H.import_files(d, str(tmp_path / "short_labels.txt"), None)
    got = H.import_files(d, l, ff)
    H.export_binary(tmp_path / "data.psmb", got["x"], y, f)
    back = H.import_files(tmp_path / "data.psmb")
    assert np.array_equal(back["x"].view(np.uint64), got["x"].view(np.uint64)) and np.array_equal(back["y"], y) and np.array_equal(back["f"], f)
    assert back["nFolds"] == 4 and back["m"] == m


@needs_ref
def test_predictions_writer_is_byte_identical_to_the_reference(tmp_path):
    from parsmurf_b200 import host_api as H
    rng = np.random.default_rng(6)
    n = 60_000
    c1 = rng.random(n) * rng.choice([1e-12, 1e-5, 1.0], size=n); c1[:6] = [0.0, 1.0, 0.5, 1e-300, 0.123456789, 0.1]
    c2 = 1.0 - c1
    y = (rng.random(n) < 0.5).astype(np.uint32)
    for labels in (None, y):
        O.ref_save(c1, c2, labels, tmp_path / "ref.txt")
        want = (tmp_path / "ref.txt").read_bytes()
        for threads in (1, 5):
            H.save_predictions(tmp_path / "ours.txt", c1, c2, labels, threads=threads)
            assert (tmp_path / "ours.txt").read_bytes() == want
    H.save_predictions(tmp_path / "ours.bin", c1, c2)
    raw = (tmp_path / "ours.bin").read_bytes()
    assert raw[:5] == b"PSMP1" and np.array_equal(np.frombuffer(raw, np.float64, n, 16), c1) and np.array_equal(np.frombuffer(raw, np.float64, n, 16 + 8 * n), c2)

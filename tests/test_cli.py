"""The C++ command-line driver (parsmurf_b200_cv --cfg file.json): parSMURF1's configuration surface end to end."""
import json
import os
import re
import subprocess

import numpy as np
import pytest

from helpers import make_case

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "parsmurf_b200", "_build", "parsmurf_b200_cv")


def _write_case(tmp_path, case, name):
    np.savetxt(tmp_path / "data.txt", case["X"], fmt="%.17g")
    np.savetxt(tmp_path / "labels.txt", case["y"], fmt="%d")
    np.savetxt(tmp_path / "folds.txt", case["f"], fmt="%d")
    cfg = {"name": name,
           "exec": {"name": "parSMURF1", "ensThrd": 4, "rfThrd": 1, "seed": 7, "verboseLevel": 0, "saveTime": True, "timeFile": str(tmp_path / "time.txt"),
                    "printCfg": False, "mode": "cv", "optimizer": "no"},
           "data": {"dataFile": str(tmp_path / "data.txt"), "foldFile": str(tmp_path / "folds.txt"), "labelFile": str(tmp_path / "labels.txt"),
                    "outFile": str(tmp_path / "predictions.txt")},
           "folds": {"nFolds": case["nF"]},
           "params": {"nParts": [case["nParts"]], "fp": [case["fp"]], "ratio": [case["ratio"]], "k": [case["k"]], "nTrees": [case["T"]], "mtry": [case["mtry"]]}}
    (tmp_path / "cfg.json").write_text(json.dumps(cfg, indent=1))
    return str(tmp_path / "cfg.json")


def test_cli_usage_and_no_device_failure(tmp_path):
    assert os.path.exists(BIN), "run __graft_entry__.build() first"
    r = subprocess.run([BIN, "--help"], capture_output=True, text=True)
    assert r.returncode == 0 and "--cfg" in r.stdout
    r = subprocess.run([BIN], capture_output=True, text=True)
    assert r.returncode == 2
    import torch
    if not torch.cuda.is_available():
        cfg = _write_case(tmp_path, make_case("fy_26"), "nogpu")
        r = subprocess.run([BIN, "--cfg", cfg], capture_output=True, text=True)
        assert r.returncode == 1 and "no usable CUDA device" in r.stderr     # fails loudly: there is no CPU fallback


@pytest.mark.gpu
def test_cli_cv_matches_host_api(tmp_path):
    from oracle import pyoracle as O
    case = make_case("fy_26")
    cfg = _write_case(tmp_path, case, "cli_fy26")
    r = subprocess.run([BIN, "--cfg", cfg], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    mt = re.search(r"AUROC: ([0-9.eE+-]+) - AUPRC: ([0-9.eE+-]+)", r.stdout)
    assert mt, r.stdout
    x = O.make_x(case["X"], case["y"])
    rc, c1, c2 = O.o_cv(x, case["y"], case["f"], case["nF"], case["nParts"], case["fp"], case["ratio"], case["k"], case["T"], case["mtry"], 7)
    want = O.o_curves(case["y"], c1)
    assert abs(float(mt.group(1)) - want[0]) < 1e-8 and abs(float(mt.group(2)) - want[1]) < 1e-8
    pred = np.loadtxt(tmp_path / "predictions.txt")          # P(pos) \t P(neg); the label column only for simulated data (src/parSMURF.cpp:128-137)
    assert pred.shape == (case["n"], 2)
    assert np.allclose(pred[:, 0], c1, rtol=1e-5, atol=1e-12) and np.allclose(pred[:, 1], c2, rtol=1e-5, atol=1e-12)
    assert os.path.exists(tmp_path / "time.txt")


@pytest.mark.gpu
def test_cli_two_gpus_through_the_library_communicator(tmp_path):
    """--gpus 2: one rank per GPU inside one process, units sharded, scores summed by psm_comm_allreduce_sum_f64 (NCCL bound by
    the C ABI -- no torch.distributed anywhere); rank 0 writes the same predictions as the single-GPU run up to the fp64
    addition order across ranks"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    case = make_case("fy_26")
    cfg = _write_case(tmp_path, case, "cli_2gpu")
    r1 = subprocess.run([BIN, "--cfg", cfg], capture_output=True, text=True)
    assert r1.returncode == 0, r1.stderr
    p1 = np.loadtxt(tmp_path / "predictions.txt")
    os.remove(tmp_path / "predictions.txt")
    r2 = subprocess.run([BIN, "--cfg", cfg, "--gpus", "2"], capture_output=True, text=True)
    assert r2.returncode == 0, r2.stderr
    assert "2 GPUs" in r2.stdout
    p2 = np.loadtxt(tmp_path / "predictions.txt")
    assert p1.shape == p2.shape and np.allclose(p1, p2, rtol=1e-5, atol=1e-12)
    m1 = re.search(r"AUROC: ([0-9.eE+-]+) - AUPRC: ([0-9.eE+-]+)", r1.stdout)
    m2 = re.search(r"AUROC: ([0-9.eE+-]+) - AUPRC: ([0-9.eE+-]+)", r2.stdout)
    assert abs(float(m1.group(1)) - float(m2.group(1))) < 1e-4 and abs(float(m1.group(2)) - float(m2.group(2))) < 1e-4


@pytest.mark.gpu
def test_cli_train_then_predict_modes(tmp_path):
    """exec.mode "train" writes one ranger forest file per partition into data.forestDir; exec.mode "predict" scores every
    example with them -- the same numbers as the host API's predict mode"""
    from oracle import pyoracle as O
    from parsmurf_b200 import host_api as H
    case = make_case("fy_26")
    cfg_path = _write_case(tmp_path, case, "cli_modes")
    cfg = json.loads(open(cfg_path).read())
    fdir = tmp_path / "forests"
    fdir.mkdir()
    cfg["data"]["forestDir"] = str(fdir)
    del cfg["data"]["foldFile"]                               # train / predict use one fold holding every example
    for mode in ("train", "predict"):
        cfg["exec"]["mode"] = mode
        (tmp_path / ("cfg_%s.json" % mode)).write_text(json.dumps(cfg, indent=1))
    r = subprocess.run([BIN, "--cfg", str(tmp_path / "cfg_train.json")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "forests saved" in r.stdout and not os.path.exists(tmp_path / "predictions.txt")
    assert sorted(os.listdir(fdir)) == ["%d.out.forest" % p for p in range(case["nParts"])]
    r = subprocess.run([BIN, "--cfg", str(tmp_path / "cfg_predict.json")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    mt = re.search(r"AUROC: ([0-9.eE+-]+) - AUPRC: ([0-9.eE+-]+)", r.stdout)
    assert mt, r.stdout
    x = O.make_x(case["X"], case["y"])
    out = H.predict_run(x, case["y"], fdir, case["nParts"], case["T"], case["mtry"], 7)
    assert abs(float(mt.group(1)) - out["auroc"]) < 1e-8 and abs(float(mt.group(2)) - out["auprc"]) < 1e-8
    pred = np.loadtxt(tmp_path / "predictions.txt")
    assert np.allclose(pred[:, 0], out["class1"], rtol=1e-5, atol=1e-12)
    assert out["auroc"] > 0.9                                  # resubstitution: the forests have seen these rows

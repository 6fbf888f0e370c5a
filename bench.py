#!/usr/bin/env python
"""bench.py -- hyperSMURF partition pipeline (kNN + SMOTE + RF train + predict + ensemble + AUROC/AUPRC) on B200.

A "step" is one complete cross-validation of BASELINE.json configs[1]:
    synthetic 1M x 26, imbalanced 1:1000, nParts=100, fp=2, ratio=3, k=5, nTrees=10, mtry=5, 5-fold CV  (500 partitions)
metric = partitions/sec (whole job, all GPUs); ms_per_step = CV wall time.

  python bench.py --gpus N --steps K --warmup W                   our arm (CUDA, sm_100a)
  python bench.py --impl reference --gpus N --steps K --warmup W  the reference's own CPU code on the host cores
  torchrun ... bench.py --gpus N ...                              N ranks, partitions sharded, one NCCL all-reduce per fold

`value`  : dataset already resident in HBM when the timed region starts (device pointer attached).
`e2e`    : the same CV through the public host API with the dataset in pinned HOST memory: the H2D copy of x, the
           index/draw uploads and the D2H of the scores are inside the timed region.
`roofline`: split_search kernel -- algorithmic bytes (sum over searched nodes of n_node*(mtry*8+8)) divided by
           its accumulated CUDA-event time inside the timed region, against MEASURED_PEAKS.json hbm_gbs.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs[1]
    "cfg2": dict(n=1_000_000, m=26, npos=1000, nFolds=5, nParts=100, fp=2, ratio=3, k=5, nTrees=10, mtry=5, seed=1, shift=5),
    # BASELINE.json configs[2] (Mendelian scale; fp/ratio/k assumed as configs[1]) and configs[3] (high-dimensional)
    "cfg3": dict(n=14_000_000, m=26, npos=400, nFolds=10, nParts=1000, fp=2, ratio=3, k=5, nTrees=10, mtry=5, seed=1, shift=5),
    "cfg4": dict(n=200_000, m=1000, npos=10_000, nFolds=5, nParts=10, fp=2, ratio=3, k=10, nTrees=10, mtry=31, seed=1, shift=50),
    # configs[2] at a tenth of the rows: same positives / partition shape / tree sizes, 140k test rows per fold (profiling aid, not a bench line)
    "cfg3s": dict(n=1_400_000, m=26, npos=400, nFolds=10, nParts=100, fp=2, ratio=3, k=5, nTrees=10, mtry=5, seed=1, shift=5),
    # small smoke-sized variant for quick checks (not a bench line)
    "small": dict(n=50_000, m=26, npos=200, nFolds=5, nParts=10, fp=2, ratio=3, k=5, nTrees=10, mtry=5, seed=1, shift=5),
}


def make_data(w):
    """Seeded synthetic imbalanced data of the named shape: iid N(0,1) features, positives shifted by +1 on
    the first `shift` features so AUROC < 1; exactly `npos` positives; stratified folds."""
    rng = np.random.default_rng(1)
    n, m = w["n"], w["m"]
    x = np.empty((n, m + 1), dtype=np.float64)
    x[:, :m] = rng.standard_normal((n, m))
    y = np.zeros(n, np.uint32)
    y[rng.choice(n, w["npos"], replace=False)] = 1
    x[y == 1, : w["shift"]] += 1.0
    x[:, m] = np.where(y > 0, 1.0, 2.0)
    f = np.zeros(n, np.uint32)
    for c in (0, 1):
        idx = np.flatnonzero(y == c)
        rng.shuffle(idx)
        f[idx] = np.arange(len(idx)) % w["nFolds"]
    return x, y, f


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) >= 9:
                for i, nm in enumerate(names):
                    if r[5 + i].lower().startswith("active"):
                        reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)).get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(launches_per_step):
    """DRAM bytes per split-search launch from the committed ncu capture of the same command (profiles/<tag>_summary.json,
    written by scripts/summarize_profiles.py: dram__bytes_read.sum + dram__bytes_write.sum over every split-search launch of
    one CV), expressed per launch of this run."""
    import glob
    best = None
    for p in sorted(glob.glob(os.path.join(ROOT, "profiles", "*_summary.json"))):
        try:
            d = json.load(open(p)).get("split_search_dram")
        except Exception:
            d = None
        if d:
            best = (d, os.path.basename(p))
    if not best or launches_per_step <= 0:
        return None, None
    d, name = best
    return (d["dram_read_bytes"] + d["dram_write_bytes"]) / launches_per_step, "profiles/%s (ncu dram__bytes_read.sum + dram__bytes_write.sum over the %d split-search launches of one CV)" % (name, d["launches"])


# ------------------------------------------------------------------------------------------------ reference arm
def cpu_sample(w, x, y, f, parts_sample, threads, export=False):
    """Time the reference's own CPU code (oracle/_ref, unmodified classes) on a bounded sample: `parts_sample` partitions of
    fold 0 with `threads` OpenMP threads over partitions (ensThrd), rfThrd 1 -- falls back to the oracle port if the
    compiled reference did not travel.  export=True also returns what the parity check needs (the draws every partition
    consumed, the reference's raw per-partition predictions, the fold's index arrays)."""
    from oracle import pyoracle as O
    if O.have_ref():
        ctx = O.RefCtx(x, y, f, w["nFolds"])
        a = (0, 0, parts_sample, w["nParts"], w["fp"], w["ratio"], w["k"], w["nTrees"], w["mtry"], w["seed"], threads)
        if export:
            secs, c1, _, draws, preds, fd = ctx.run_partitions_ex(*a)
            return secs, "reference", dict(draws=draws, preds=preds, fd=fd, class1=c1)
        secs, _, _ = ctx.run_partitions(*a)
        return secs, "reference", None
    t0 = time.time()
    # the port has no partition sub-range: scale a full fold by the sample fraction
    O.o_cv(x, y, f, w["nFolds"], w["nParts"], w["fp"], w["ratio"], w["k"], w["nTrees"], w["mtry"], w["seed"], 0, 1, threads)
    return (time.time() - t0) * parts_sample / w["nParts"], "port", None


def mpi_probe():
    """BASELINE.md 3.4 / SURVEY 8(d): the parSMURFn (MPI) CPU baseline needs an MPI toolchain on the box"""
    import shutil
    exe = shutil.which("mpirun") or shutil.which("mpiexec")
    return "not runnable: no MPI (mpirun / mpiexec not found on this box)" if not exe else "mpirun present at %s, but parSMURFn is not built (no MPI headers where oracle/_ref was compiled)" % exe


def run_reference(args, w):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    x, y, f = make_data(w)
    cores = os.cpu_count() or 1
    sample = max(1, min(w["nParts"], cores))
    times = []
    kind = "reference"
    for i in range(args.warmup + args.steps):
        secs, kind, _ = cpu_sample(w, x, y, f, sample, cores)
        if i >= args.warmup:
            times.append(secs)
    total = sum(times)
    value = sample * len(times) / total
    desc = "%d of %d partitions of fold 0 per step (Nte=%d test rows each), %d OpenMP threads over partitions, rfThrd 1" % (sample, w["nParts"], w["n"] // w["nFolds"], cores)
    line = {"impl": "reference", "metric": "partitions/sec", "value": value, "unit": "partitions/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1000.0 * total / len(times), "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, w),
            "cpu_baseline": {"value": value, "unit": "partitions/s", "cores": cores, "kind": kind, "sample": desc},
            "mpi_baseline": mpi_probe(),
            "e2e": {"value": value, "unit": "partitions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


WL_NAMES = {"cfg2": "configs[1]", "cfg3": "configs[2]", "cfg4": "configs[3]", "cfg5_grid": "configs[4]"}


def workload_config(args, w, name=None):
    """identical in both arms (the driver compares the two `config` objects)"""
    name = name or args.workload
    return {"workload": "BASELINE.json %s: synthetic %dx%d imbalanced 1:%d, nParts=%d, fp=%d, ratio=%d, k=%d, nTrees=%d, mtry=%d, %d-fold CV"
            % (WL_NAMES.get(name, name), w["n"], w["m"], w["n"] // w["npos"], w["nParts"], w["fp"], w["ratio"], w["k"], w["nTrees"], w["mtry"], w["nFolds"]),
            "partitions_per_step": w["nParts"] * w["nFolds"], "parallelism": ("single GPU" if args.gpus == 1 else
                            "%d GPUs, dataset replicated, %s" % (args.gpus, "contiguous ranges of the nFolds x nParts (fold, partition) units per rank, ONE NCCL all-reduce of the score arrays [2n] f64 at the end (+ one of the per-fold metrics), issued by the library (psm_comm_allreduce_sum_f64); only rank 0 fetches the scores"
                                                                if args.shard != "fold" else "every rank takes its chunk of every fold's partitions (the reference's MPI rule), one NCCL all-reduce of the fold scores per fold")),
            "l2_policy": "inputs larger than L2 (dataset %d MB + per-fold training batches of GBs >> 126 MB L2); no explicit flush" % (w["n"] * (w["m"] + 1) * 8 // 1_000_000),
            "curves_tie_mode": "auto: device radix sort; exact unless a run of tied scores holds both labels, then bracketed by the positives-first / negatives-first tie orders (midpoint if they agree to 1e-6, else the reference's host std::sort)"}


GRID = dict(nParts=[10, 20, 30], fp=[1, 2], ratio=[1], k=[5], nTrees=[10], mtry=[5, 7, 10])   # cfgEx/gridTune.json "params"


# ------------------------------------------------------------------------------------------------ our arm
class Rig:
    """what every measurement needs: torch, the process group, the stream, one persistent GPU context"""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        from parsmurf_b200 import host_api as H
        self.torch, self.dist, self.H, self.args = torch, dist, H, args
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device -- parsmurf_b200 has no CPU fallback")
        torch.cuda.set_device(self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
        self.stream = torch.cuda.current_stream()
        self.hdev = H.HostDevice(self.local, self.stream.cuda_stream)   # one long-lived GPU context, as in a long-lived host process
        # the cross-rank sum runs inside the library (psm_comm_*: NCCL bound by the C ABI); torch.distributed only carries the
        # 128-byte communicator id from rank 0 to the others.  --reduce torch keeps the old callback around dist.all_reduce.
        self.lib_reduce = self.world > 1 and args.reduce == "library"
        if self.lib_reduce:
            idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
            if self.rank == 0:
                idt.copy_(torch.frombuffer(bytearray(H.HostDevice.comm_unique_id()), dtype=torch.uint8))
            dist.broadcast(idt, 0)
            self.hdev.comm_init(bytes(idt.cpu().numpy().tobytes()), self.rank, self.world)
            self.comm = self.hdev.comm_info()

    def reduce(self, ptr, count):
        torch = self.torch

        class _Arr:   # zero-copy view of a raw device pointer as a torch tensor
            __cuda_array_interface__ = {"shape": (count,), "typestr": "<f8", "data": (ptr, False), "version": 2}
        t = torch.as_tensor(_Arr(), device="cuda")
        self.dist.all_reduce(t)
        torch.cuda.current_stream().synchronize()

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def timed(self, fn, steps):
        """K steps bracketed by barrier + synchronize, CUDA events on the launching stream, max over ranks"""
        torch = self.torch
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        marks = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]   # one more event per step: the spread of the steps (this rank)
        e0.record(self.stream)
        outs = []
        for i in range(steps):
            outs.append(fn())
            marks[i].record(self.stream)
        e1.record(self.stream)
        self.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
        if self.world > 1:
            self.dist.all_reduce(ms, op=self.dist.ReduceOp.MAX)
        self.last_step_ms = [round(([e0] + marks)[i].elapsed_time(marks[i]), 3) for i in range(steps)]
        return float(ms.item()), outs

    def stage(self, x):
        torch = self.torch
        x_pinned = torch.from_numpy(x).pin_memory()
        x_dev = x_pinned.to("cuda", non_blocking=False)
        torch.cuda.synchronize()
        return x_pinned, x_dev


def measure_cv(rig, w, x_pinned, x_dev, y, f, steps, warmup, roof_steps, with_breakdown):
    """the measurements of one workload: HBM-resident CV, e2e CV from pinned host memory, split-search roofline region"""
    args = rig.args
    n, m = w["n"], w["m"]
    x_host = x_pinned.numpy()
    common = dict(labels=y, folds=f, nFolds=w["nFolds"], nParts=w["nParts"], fp=w["fp"], ratio=w["ratio"], k=w["k"], nTrees=w["nTrees"], mtry=w["mtry"],
                  seed=w["seed"], device=rig.local, stream=rig.stream.cuda_stream, tie_mode=2, eval_fold_curves=True, rank=rig.rank, world=rig.world,
                  reduce=rig.reduce if (rig.world > 1 and not rig.lib_reduce) else None, profile=0, fold_lanes=args.fold_lanes, shard=args.shard,
                  scores_root_only=rig.world > 1)
    scores = (np.zeros(n), np.zeros(n))               # class1Prob / class2Prob host arrays, reused by every step
    run_dev = lambda: rig.hdev.cv_run(None, x_device_ptr=x_dev.data_ptr(), n=n, m=m, out=scores, **common)
    run_e2e = lambda: rig.hdev.cv_run(x_host, out=scores, **common)
    for _ in range(warmup):
        run_dev()
    sampler = ClockSampler(rig.local)
    if rig.rank == 0:
        sampler.start()
    ms_dev, outs = rig.timed(run_dev, steps)
    each_dev = list(rig.last_step_ms)
    clocks = sampler.stop() if rig.rank == 0 else None
    for _ in range(min(warmup, 1)):
        run_e2e()
    ms_e2e, outs_e2e = rig.timed(run_e2e, steps)
    # roofline kernel: its own timed region, CUDA events around every split-search launch (on the stream it is launched on),
    # folds one after the other -- with several fold lanes in flight the kernels of different folds share the SMs and an
    # event pair around one launch also measures the others' work
    run_roof = lambda: rig.hdev.cv_run(None, x_device_ptr=x_dev.data_ptr(), n=n, m=m, out=scores, **dict(common, profile=2, fold_lanes=1))
    run_roof()
    ms_roof, outs_roof = rig.timed(run_roof, roof_steps)
    # one more, untimed, step with every kernel class under events: the per-kernel breakdown reported next to the roofline
    brk = rig.hdev.cv_run(None, x_device_ptr=x_dev.data_ptr(), n=n, m=m, out=scores, **dict(common, profile=1, fold_lanes=1)) if with_breakdown else None

    parts = w["nParts"] * w["nFolds"]
    lanes = outs_e2e[0].get("fold_lanes") or 1
    # per-step traffic of the e2e path (this rank), counted from what is copied: x once, and -- because a new upload of x drops
    # the resident labels and fold ids -- the 0/1 labels (4n) and the fold ids (4n) once each; the per-class fold lists, every
    # fold's index arrays and the SMOTE / shuffle draws are built on the device (no H2D).  Back: class1Prob/class2Prob once
    # (2 x n f64; rank 0 only in multi-rank runs), AUROC/AUPRC per fold and global, the fold sizes, a 32-byte status word per tree level.
    # Whole job (all ranks): with the library communicator the ranks share the upload of x -- a 1/world slice each over PCIe, one
    # NCCL all-gather over NVLink (psm_dataset_upload_sharded) -- otherwise every rank uploads its own copy.
    x_total = x_host.nbytes if (rig.world == 1 or rig.lib_reduce) else x_host.nbytes * rig.world
    h2d = x_total + rig.world * (4 * n + 4 * n)
    d2h = 16 * n + rig.world * ((w["nFolds"] + 1) * 16 + 8 * w["nFolds"] + 32 * int(sum(o["stats"]["levels"] for o in outs_e2e) / max(steps, 1)))
    km = {k_: sum(o["kernel_ms"][k_] for o in outs_roof) for k_ in outs_roof[0]["kernel_ms"]}
    kl = {k_: sum(o["kernel_launches"][k_] for o in outs) for k_ in outs[0]["kernel_launches"]}
    kl_roof = sum(o["kernel_launches"]["split_search"] for o in outs_roof)
    split_bytes = sum(o["stats"]["split_bytes"] for o in outs_roof)
    assert split_bytes * steps == sum(o["stats"]["split_bytes"] for o in outs) * roof_steps, "the roofline region did not do the work of the timed region"
    peak, peak_src = peaks()
    ss_ms, ss_launches = km["split_search"], max(kl_roof, 1)
    achieved = split_bytes / (ss_ms / 1000.0) / 1e9 if ss_ms > 0 else 0.0
    roofline = {"kernel": "split_search_merged_kernel", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None, "peak_source": peak_src, "algorithmic_bytes_per_launch": split_bytes / ss_launches,
                "avg_launch_ms": ss_ms / ss_launches, "launches": kl_roof, "kernel_ms_per_step_in_region": ss_ms / roof_steps,
                "timed_region": "%d steps with the folds one after the other (fold lanes = 1), %.2f ms per step; events around every split-search launch" % (roof_steps, ms_roof / roof_steps)}
    if brk is not None:
        roofline["kernel_ms_per_step"] = dict(brk["kernel_ms"])
        roofline["kernel_ms_per_step_source"] = "one extra untimed step (fold lanes = 1) with all kernel classes under CUDA events"
    return dict(ms_dev=ms_dev, ms_e2e=ms_e2e, each_dev=each_dev, outs=outs, outs_e2e=outs_e2e, clocks=clocks, parts=parts, h2d=int(h2d), d2h=int(d2h), kl=kl,
                roofline=roofline, scores=scores, lanes=outs[0].get("fold_lanes"), kl_roof=kl_roof,
                curves={"mode": "auto", "mixed_tie_pairs_per_step": sum(o["stats"]["mixed_tie_pairs"] for o in outs) // max(steps, 1),
                        "host_sorts_per_step": sum(o["stats"]["curves_host_sorts"] for o in outs) / max(steps, 1),
                        "max_tie_spread": max(o["stats"]["curves_max_tie_spread"] for o in outs)})


def parity_block(rig, w, name, x_dev, y, exported, sample):
    """the CUDA path against the reference's own classes at the benchmarked shape: the partitions the cpu_baseline leg ran,
    with the draws those partitions consumed injected, through the C ABI (psm_fold_begin / psm_fold_knn /
    psm_fold_run_partitions / psm_fold_scores_get).  max_abs_diff must be 0.0."""
    import parsmurf_b200 as psm
    from oracle import pyoracle as O   # the checker
    ctx = psm.Context(rig.local)
    try:
        ctx.dataset_attach_device(x_dev.data_ptr(), w["n"], w["m"])
        r = O.scale_parity(ctx, exported["fd"], exported["draws"], exported["preds"], 0, sample, w["nParts"], w["fp"], w["ratio"], w["k"], w["nTrees"],
                           w["mtry"], w["seed"], 0, ref_class1=exported["class1"])
    finally:
        ctx.close()
    r["config"] = WL_NAMES.get(name, name)
    r["what"] = ("partitions [0,%d) of fold 0: reference = unmodified classes (oracle/_ref), %d test rows; GPU = C ABI with the reference's SMOTE draws "
                 "injected; compared bit for bit with sum_p pred_p * (1/nParts) in ascending p" % (sample, r["rows"]))
    return r


def tie_check(rig, w, y, class1):
    """AUROC / AUPRC of the final scores under the three tie modes (0 = the reference's own std::sort call on the host)"""
    import parsmurf_b200 as psm
    ctx = psm.Context(rig.local)
    try:
        a0 = ctx.curves(y, class1, tie_mode=0)
        a1 = ctx.curves(y, class1, tie_mode=1)
        a2 = ctx.curves(y, class1, tie_mode=2)
        mixed, host, spread = ctx.curves_info()
    finally:
        ctx.close()
    return {"auroc_auto_minus_tie0": a2[0] - a0[0], "auprc_auto_minus_tie0": a2[1] - a0[1], "auroc_tie1_minus_tie0": a1[0] - a0[0],
            "auprc_tie1_minus_tie0": a1[1] - a0[1], "mixed_tie_pairs_global": mixed, "auto_used_host_sort_global": bool(host), "tie_spread_global": list(spread)}


def also_block(rig, name, w, x_pinned, x_dev, y, f, steps, warmup, with_parity):
    r = measure_cv(rig, w, x_pinned, x_dev, y, f, steps, warmup, 1, False)
    blk = {"workload": workload_config(rig.args, w, name)["workload"], "steps": steps, "warmup": warmup, "ms_per_cv": r["ms_dev"] / steps,
           "partitions_per_sec": r["parts"] * steps / (r["ms_dev"] / 1000.0), "examples_per_sec": w["n"] * steps / (r["ms_dev"] / 1000.0),
           "e2e": {"ms_per_cv": r["ms_e2e"] / steps, "partitions_per_sec": r["parts"] * steps / (r["ms_e2e"] / 1000.0),
                   "h2d_bytes_per_step": r["h2d"], "d2h_bytes_per_step": r["d2h"]},
           "roofline": r["roofline"], "clocks": r["clocks"], "fold_lanes": r["lanes"], "gpu_launches": int(sum(r["kl"].values())),
           "quality": {"auroc": r["outs"][-1]["auroc"], "auprc": r["outs"][-1]["auprc"]}, "curves": r["curves"]}
    if with_parity and rig.world == 1 and rig.rank == 0:
        cores = os.cpu_count() or 1
        sample = max(1, min(w["nParts"], cores, 16))
        try:
            secs, kind, exported = cpu_sample(w, x_pinned.numpy(), y, f, sample, sample, export=True)
            blk["cpu_baseline"] = {"value": sample / secs, "unit": "partitions/s", "cores": sample, "kind": kind,
                                   "sample": "%d of %d partitions of fold 0 (Nte=%d), one OpenMP thread each, rfThrd 1, %.1f s" % (sample, w["nParts"], w["n"] // w["nFolds"], secs)}
            blk["parity"] = parity_block(rig, w, name, x_dev, y, exported, sample) if exported else {"unavailable": "oracle/_ref did not travel"}
        except Exception as e:   # the checker must not take the measurement down
            blk["parity"] = {"error": repr(e)[:300]}
    return blk


def grid_block(rig, w, x_dev, y, f):
    """BASELINE.json configs[4]: cfgEx/gridTune.json's 18-point grid (nParts x fp x mtry) on the configs[1] data -- every outer fold
    runs an inner (nFolds-1)-fold CV per grid point, then the outer fold with the arg-max-AUPRC point (src/hyperSMURF.cpp:176-228)"""
    import itertools
    grid = [(a, b, c, d, e, g) for a, b, c, d, e, g in itertools.product(GRID["nParts"], GRID["fp"], GRID["ratio"], GRID["k"], GRID["nTrees"], GRID["mtry"])]
    run = lambda: rig.hdev.cv_grid_run(None, y, f, w["nFolds"], grid, w["seed"], device=rig.local, tie_mode=2, stream=rig.stream.cuda_stream, rank=rig.rank,
                                       world=rig.world, reduce=rig.reduce if (rig.world > 1 and not rig.lib_reduce) else None, x_device_ptr=x_dev.data_ptr(), n=w["n"], m=w["m"])
    run()
    sampler = ClockSampler(rig.local)
    if rig.rank == 0:
        sampler.start()
    ms, outs = rig.timed(run, 1)
    clocks = sampler.stop() if rig.rank == 0 else None
    inner = w["nFolds"] * (w["nFolds"] - 1) * sum(g[0] for g in grid)
    return {"workload": "BASELINE.json configs[4]: cfgEx/gridTune.json grid (nParts %s x fp %s x mtry %s = %d points, ratio 1, k 5, nTrees 10) on the configs[1] data, %d-fold CV with inner %d-fold CVs"
                        % (GRID["nParts"], GRID["fp"], GRID["mtry"], len(grid), w["nFolds"], w["nFolds"] - 1),
            "steps": 1, "warmup": 1, "ms_per_grid_cv": ms, "inner_partitions": inner, "inner_partitions_per_sec": inner / (ms / 1000.0), "clocks": clocks,
            "gpu_launches": int(sum(outs[0]["kernel_launches"].values())), "quality": {"auroc": outs[0]["auroc"], "auprc": outs[0]["auprc"]},
            "sharding": "single GPU" if rig.world == 1 else "every rank runs whole inner CVs -- its cost-balanced share of the 18 grid points of each outer fold -- and the [18][2] metric table is summed once per outer fold; the outer folds' partitions are chunked over the ranks",
            "curves": {"mode": "auto", "mixed_tie_pairs": outs[0]["stats"]["mixed_tie_pairs"], "host_sorts": outs[0]["stats"]["curves_host_sorts"],
                       "max_tie_spread": outs[0]["stats"]["curves_max_tie_spread"]}}


def run_ours(args, w):
    rig = Rig(args)
    torch, dist, rank, world = rig.torch, rig.dist, rig.rank, rig.world
    x, y, f = make_data(w)
    n, m = w["n"], w["m"]
    x_pinned, x_dev = rig.stage(x)
    del x

    if args.single:   # profiling aid (scripts/collect_profiles.sh): just the CVs, no timing regions, no JSON line
        common = dict(labels=y, folds=f, nFolds=w["nFolds"], nParts=w["nParts"], fp=w["fp"], ratio=w["ratio"], k=w["k"], nTrees=w["nTrees"], mtry=w["mtry"],
                      seed=w["seed"], device=rig.local, stream=rig.stream.cuda_stream, tie_mode=2, eval_fold_curves=True, rank=rank, world=world,
                      reduce=rig.reduce if (world > 1 and not rig.lib_reduce) else None, profile=0, fold_lanes=args.fold_lanes, shard=args.shard,
                      scores_root_only=world > 1)
        scores = (np.zeros(n), np.zeros(n))
        for _ in range(max(args.steps, 1)):
            rig.hdev.cv_run(None, x_device_ptr=x_dev.data_ptr(), n=n, m=m, out=scores, **common)
        torch.cuda.synchronize()
        if world > 1:
            dist.destroy_process_group()
        return

    r = measure_cv(rig, w, x_pinned, x_dev, y, f, args.steps, args.warmup, args.steps, True)
    ms_dev, ms_e2e, outs, outs_e2e = r["ms_dev"], r["ms_e2e"], r["outs"], r["outs_e2e"]
    parts = r["parts"]
    value = parts * args.steps / (ms_dev / 1000.0)
    e2e_value = parts * args.steps / (ms_e2e / 1000.0)
    roofline = r["roofline"]
    if args.workload == "cfg2" and world == 1:
        roofline["traffic"], roofline["traffic_source"] = ncu_traffic(r["kl_roof"] / max(args.steps, 1))
    line = None
    if rank == 0:
        line = {"metric": "partitions/sec", "value": value, "unit": "partitions/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_dev / args.steps, "ms_each_step_rank0": r["each_dev"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": workload_config(args, w), "fold_lanes": r["lanes"], "clocks": r["clocks"],
                "e2e": {"value": e2e_value, "unit": "partitions/s", "ms_per_step": ms_e2e / args.steps, "h2d_bytes_per_step": r["h2d"], "d2h_bytes_per_step": r["d2h"]},
                "gpu_launches": int(sum(r["kl"].values())), "roofline": roofline,
                "examples_per_sec": n * args.steps / (ms_dev / 1000.0),
                "quality": {"auroc": outs[-1]["auroc"], "auprc": outs[-1]["auprc"]}, "curves": r["curves"],
                "host_phases_ms_per_step": {k_: 1000.0 * sum(o["host_phases_s"][k_] for o in outs) / args.steps for k_ in outs[0]["host_phases_s"]},
                "host_phases_ms_per_step_e2e": {k_: 1000.0 * sum(o["host_phases_s"][k_] for o in outs_e2e) / args.steps for k_ in outs_e2e[0]["host_phases_s"]},
                "mpi_baseline": mpi_probe()}
        if rig.lib_reduce:
            line["comm"] = dict(rig.comm, where="inside libparsmurf_b200.so: psm_comm_init / psm_comm_allreduce_sum_f64 (ncclAllReduce on the context's stream)")
        if world == 1:
            try:
                line["curves"].update(tie_check(rig, w, y, r["scores"][0]))
            except Exception as e:
                line["curves"]["tie_check_error"] = repr(e)[:300]
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            sample = max(1, min(w["nParts"], cores))
            secs, kind, exported = cpu_sample(w, x_pinned.numpy(), y, f, sample, cores, export=True)
            line["cpu_baseline"] = {"value": sample / secs, "unit": "partitions/s", "cores": cores, "kind": kind,
                                    "sample": "%d of %d partitions of fold 0 (Nte=%d), %d OpenMP threads over partitions, rfThrd 1, %.1f s" % (sample, w["nParts"], n // w["nFolds"], cores, secs)}
            try:
                line["parity"] = parity_block(rig, w, args.workload, x_dev, y, exported, sample) if exported else {"unavailable": "oracle/_ref did not travel"}
            except Exception as e:
                line["parity"] = {"error": repr(e)[:300]}
    # the other BASELINE configs, a few steps each, in the same driver-visible line: configs[4] on the
    # configs[1] data, then configs[2] (the north-star target shape) and configs[3]
    if args.workload == "cfg2" and not args.no_also:
        also = {}
        for name in [v for v in args.also.split(",") if v]:
            try:
                if name == "cfg5_grid":
                    blk = grid_block(rig, w, x_dev, y, f)
                else:
                    if x_dev is not None:
                        del x_pinned, x_dev
                        x_pinned = x_dev = None
                        torch.cuda.empty_cache()
                    w2 = WORKLOADS[name]
                    x2, y2, f2 = make_data(w2)
                    xp2, xd2 = rig.stage(x2)
                    del x2
                    blk = also_block(rig, name, w2, xp2, xd2, y2, f2, 2, 1, with_parity=(name == "cfg3" and not args.no_cpu_baseline))
                    del xp2, xd2
                    torch.cuda.empty_cache()
            except Exception as e:
                blk = {"error": repr(e)[:400]}
            also[name] = blk
        if line is not None:
            line["also"] = also
    if rank == 0:
        emit(line)
    if world > 1:
        dist.destroy_process_group()


_REAL_STDOUT = None


def emit(line):
    """the ONE JSON line goes to the real stdout; everything else any library prints (NCCL's version banner, ...) goes to stderr"""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=list(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-also", action="store_true", help="skip the `also` blocks (the other BASELINE configs) of the default run")
    ap.add_argument("--also", default="cfg5_grid,cfg3,cfg4", help="comma list of the `also` blocks to run after the configs[1] measurement")
    ap.add_argument("--fold-lanes", type=int, default=None, help="folds in flight at once (default: the library's)")
    ap.add_argument("--single", action="store_true", help="run only --steps CVs and exit (for ncu passes)")
    ap.add_argument("--shard", default="unit", choices=["unit", "fold"], help="multi-GPU sharding rule (see host_api.cv_run)")
    ap.add_argument("--reduce", default="library", choices=["library", "torch"], help="multi-GPU sum: the library's own NCCL communicator (psm_comm_*) or a callback around torch.distributed.all_reduce")
    args = ap.parse_args()
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, w)
    else:
        run_ours(args, w)


if __name__ == "__main__":
    main()

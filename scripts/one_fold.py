"""One fold of a workload through the C ABI (profiling aid for ncu passes): python scripts/one_fold.py [workload] [fold]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
import parsmurf_b200 as psm
from parsmurf_b200 import host_api as H
name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
fold = int(sys.argv[2]) if len(sys.argv) > 2 else 0
w = dict(bench.WORKLOADS[name])
x, y, f = bench.make_data(w)
fd = H.ttd_fold(y, f, w["nFolds"], fold)
ctx = psm.Context(0)
ctx.dataset_upload(x)
ctx.fold_begin(fd["trngPos"], fd["trngNeg"], fd["testPos"], fd["testNeg"])
ctx.fold_knn(w["k"])
P = len(fd["trngPos"])
draws = H.rand_fill(1, P * w["fp"] * (w["m"] + 1) * w["nParts"])
ctx.fold_run_partitions(w["nParts"], w["fp"], w["ratio"], w["nTrees"], w["mtry"], w["seed"], fold, 0, w["nParts"], draws)
ctx.sync()
print("done", ctx.stats_get())

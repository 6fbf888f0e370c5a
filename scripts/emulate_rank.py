"""One rank's share of a multi-GPU CV on a single GPU (profiling aid): the sharding rule, the lanes and the host phases are the
real ones, the cross-rank reduce is a no-op.   PSM_TIMELINE=1 python scripts/emulate_rank.py cfg3 8 3 [steps] [fold_lanes]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from parsmurf_b200 import host_api as H
name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
world = int(sys.argv[2]) if len(sys.argv) > 2 else 8
rank = int(sys.argv[3]) if len(sys.argv) > 3 else 0
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
lanes = int(sys.argv[5]) if len(sys.argv) > 5 else None
w = dict(bench.WORKLOADS[name])
x, y, f = bench.make_data(w)
xd = torch.from_numpy(x).pin_memory().to("cuda")
n, m = w["n"], w["m"]
del x
stream = torch.cuda.current_stream()
hdev = H.HostDevice(0, stream.cuda_stream)
scores = (np.zeros(n), np.zeros(n))
for i in range(steps):
    torch.cuda.synchronize()
    t0 = time.time()
    o = hdev.cv_run(None, y, f, w["nFolds"], w["nParts"], w["fp"], w["ratio"], w["k"], w["nTrees"], w["mtry"], w["seed"], device=0, stream=stream.cuda_stream,
                    tie_mode=2, rank=rank, world=world, reduce=(lambda p, c: None) if world > 1 else None, x_device_ptr=xd.data_ptr(), n=n, m=m, out=scores,
                    fold_lanes=lanes, shard="unit")
    torch.cuda.synchronize()
    print("step %d: %.1f ms  phases(ms) %s" % (i, 1e3 * (time.time() - t0), {k: round(1e3 * v, 1) for k, v in o["host_phases_s"].items()}), file=sys.stderr)

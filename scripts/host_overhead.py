"""Where does host time go in one cfg-2 CV call?  wall time of the Python call vs the library's own phase timers."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from parsmurf_b200 import host_api as H
w = dict(bench.WORKLOADS["cfg2"])
x, y, f = bench.make_data(w)
n, m = x.shape[0], x.shape[1] - 1
xd = torch.from_numpy(x).cuda()
hdev = H.HostDevice(0, None)
scores = (np.zeros(n), np.zeros(n))
kw = dict(labels=y, folds=f, nFolds=w["nFolds"], nParts=w["nParts"], fp=w["fp"], ratio=w["ratio"], k=w["k"], nTrees=w["nTrees"], mtry=w["mtry"],
          seed=w["seed"], tie_mode=1, eval_fold_curves=True, profile=("--profile" in sys.argv))
for it in range(6):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    o = hdev.cv_run(None, x_device_ptr=xd.data_ptr(), n=n, m=m, out=scores, **kw)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    ph = o["host_phases_s"]
    print("wall %.2f ms  phases sum %.2f ms  " % ((t1 - t0) * 1e3, sum(ph.values()) * 1e3) + " ".join("%s=%.2f" % (k, v * 1e3) for k, v in ph.items() if v > 1e-5))

"""Per-level timing trace of one fold of BASELINE configs[1] (PSM_TRACE_LEVELS=1)."""
import os, sys
os.environ["PSM_TRACE_LEVELS"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from parsmurf_b200 import host_api as H
w = dict(bench.WORKLOADS["cfg2"]); w["nFolds"] = 5
x, y, f = bench.make_data(w)
# only fold 0: mark minFold/maxFold through a 1-fold trick is not exposed; run the whole CV but stop tracing after fold 0
out = H.cv_run(x, y, f, w["nFolds"], w["nParts"], w["fp"], w["ratio"], w["k"], w["nTrees"], w["mtry"], w["seed"], tie_mode=1)

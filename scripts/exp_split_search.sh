#!/bin/bash
# A/B measurement of the split-search launch forms on one B200 (run through gpurun from the repo root):
#   gpurun --timeout 900 -- 'bash scripts/exp_split_search.sh'
# Every line of gpurun_out/exp_ss/summary.txt: variant, ms per CV, split-search ms per CV (own timed region), roofline.frac.
set -u
OUT=gpurun_out/exp_ss
mkdir -p "$OUT"
B="python bench.py --steps 6 --warmup 4 --no-also --no-cpu-baseline"
if [ "${1:-}" = "tests" ]; then
	timeout 600 python -m pytest tests -m gpu -x -q  > "$OUT/pytest.log" 2>&1
	tail -3 "$OUT/pytest.log"
fi
run() {   # name, env assignments...
	local name=$1; shift
	timeout 180 env "$@" $B ${BARGS:-} > "$OUT/$name.json" 2> "$OUT/$name.err" || { echo "$name FAILED"; tail -5 "$OUT/$name.err"; return; }
	python - "$name" "$OUT/$name.json" <<'EOF' | tee -a "$OUT/summary.txt"
import json, sys
d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
r = d["roofline"]; k = r.get("kernel_ms_per_step", {})
print("%-14s %7.2f ms/CV  e2e %7.2f  ss %6.2f ms (%d launches)  frac %.3f  partition %.2f  predict %.2f  curves %.2f  gather %.2f  auroc %.8f" % (
    sys.argv[1], d["ms_per_step"], d["e2e"]["ms_per_step"], r["kernel_ms_per_step_in_region"], r["launches"], r["frac"],
    k.get("partition_lists", -1), k.get("predict", -1), k.get("curves", -1), k.get("gather", -1), d.get("quality", {}).get("auroc", -1)))
EOF
}
: > "$OUT/summary.txt"
V=$PWD/parsmurf_b200
run tps1 PSM_PREDICT_TPS=1
run tps2 PSM_PREDICT_TPS=2
BARGS="--workload cfg3s" run tps1_cfg3s PSM_PREDICT_TPS=1
BARGS="--workload cfg3s" run tps2_cfg3s PSM_PREDICT_TPS=2

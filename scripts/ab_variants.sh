#!/bin/bash
# A/B measurement of kernel variants on one B200 (run through gpurun from the repo root; the r02 kernel comments quote its output):
#   scripts/build_variant.sh m9 "-DSS_MINB_MERGED32=9"; gpurun --timeout 900 -- 'AB_VARIANTS="m9" bash scripts/ab_variants.sh [tests]'
# Every line of gpurun_out/exp_ss/summary.txt: variant, ms per CV, e2e, split-search ms per CV (own timed region), roofline.frac,
# per-kernel ms of the list partition, prediction, curves and fold set-up.  Environment switches (PSM_SS_BLOCK, PSM_SS_SEPARATE,
# PSM_PREDICT_TPS, PSM_CURVES_TWO_SORTS, ...) can be measured the same way: run NAME VAR=value.
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
run base
# variants: scripts/build_variant.sh NAME "-DFOO=1" first, then list the names (AB_VARIANTS="a b c")
for v in ${AB_VARIANTS:-}; do run $v PSM_BUILD_DIR=$V/_build_$v; done

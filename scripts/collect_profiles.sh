#!/bin/bash
# Collect the round's measurement set on a B200 box (run through gpurun from the repo root):
#   gpurun --timeout 1500 -- 'bash scripts/collect_profiles.sh r02'
# Every ncu pass only runs after the same command exited 0 without ncu; numbers printed under ncu are never bench values.
# gpurun copies back at most 64 MiB, so the .ncu-rep files are reduced to CSV on the box (scripts/summarize_profiles.py reads
# the CSVs).  The ncu passes run the CV with the folds one after the other (--fold-lanes 1): ncu serialises kernels anyway and
# the launch order then reads level by level.
set -u
TAG=${1:-r02}
OUT=gpurun_out/profiles_$TAG
mkdir -p "$OUT"
B="python bench.py --steps 5 --warmup 3 --no-also"
S="python bench.py --steps 1 --warmup 0 --single --fold-lanes 1"

$B > "$OUT/${TAG}_bench_cfg2_1gpu.json" 2> "$OUT/bench.log" || { echo "bench failed"; tail -20 "$OUT/bench.log"; exit 1; }
python bench.py --impl reference --steps 2 --warmup 1 > "$OUT/${TAG}_bench_cfg2_reference_cpu.json" 2>> "$OUT/bench.log" || echo "reference arm failed"

# launch list (per-launch durations, serialised and cold-cache: only the kernels' SHARES are comparable with the live timing)
$S > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv \
	--log-file "$OUT/${TAG}_launches_cfg2_1gpu.csv" $S > "$OUT/ncu_launches.log" 2>&1

# DRAM bytes of every split-search launch of one CV (roofline.traffic)
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:split_search --csv \
	--log-file "$OUT/${TAG}_dram_split_search_all_levels.csv" $S > "$OUT/ncu_dram.log" 2>&1

# full-section captures: a window of launches from the first fold for the per-level kernels, one launch for the per-fold ones
# (KERNELS="a b" restricts the list: a capture costs one profiled CV each)
for K in ${KERNELS:-split_search_merged_kernel partition_lists_kernel predict_compact_kernel feature_sort_radix_kernel build_lists_multi_kernel tree_level_kernel split_select_kernel}; do
	case $K in
		split_search_merged_kernel|partition_lists_kernel|tree_level_kernel|split_select_kernel) SKIP=0; CNT=12 ;;
		*) SKIP=0; CNT=1 ;;
	esac
	ncu --set full --clock-control none --import-source on -k regex:$K --launch-skip $SKIP --launch-count $CNT \
		-o "$OUT/${TAG}_$K" -f $S > "$OUT/ncu_$K.log" 2>&1
	ncu -i "$OUT/${TAG}_$K.ncu-rep" --page raw --csv > "$OUT/${TAG}_ncu_raw_$K.csv" 2>/dev/null
	rm -f "$OUT/${TAG}_$K.ncu-rep"
done
du -sh "$OUT"; ls -la "$OUT"

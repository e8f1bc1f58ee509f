#!/bin/bash
# Run ON THE GPU BOX (through gpurun): launch list of the default bench command and --set full
# captures of a few aggregation and dense-layer launches of one cfgC step.  Only the text
# summaries and the launch-list CSV are kept (the .ncu-rep files exceed gpurun's 64 MiB return
# limit and are deleted after they have been summarised).
# Usage: tools/capture_profiles.sh <tag> [all|launches|combine|dense]...
set -u
TAG=${1:-r01}
shift || true
PARTS=${*:-all}
want() { [[ " $PARTS " == *" all "* || " $PARTS " == *" $1 "* ]]; }
MATH=${MATH:-tf32}                 # MATH=bf16 tools/capture_profiles.sh r02_bf16 ...: the bf16 storage mode
COMBINE_PER_STEP=${COMBINE_PER_STEP:-21}   # hgin_gin_combine launches per Cfg-C step (r01: 23)
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra --math $MATH"
$CMD > gpurun_out/${TAG}_plain.json 2> gpurun_out/${TAG}_plain.err || { tail -5 gpurun_out/${TAG}_plain.err; exit 1; }
# launch list (cold-cache, serialised): ~2 steps after the warm-up steps
want launches && ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 700 -c 300 --csv \
    --log-file gpurun_out/${TAG}_launches_cfgC_${MATH}.csv $CMD > gpurun_out/${TAG}_ncu_launches.log 2>&1
# full sets on the launches of ONE step (after 4 complete steps): the aggregation kernel, then the dense layers
if want combine; then
ncu --set full --clock-control none --import-source on -k regex:gin_combine --launch-skip $((4 * COMBINE_PER_STEP)) -c $COMBINE_PER_STEP \
    -o /tmp/${TAG}_combine -f $CMD > gpurun_out/${TAG}_ncu_combine.log 2>&1
python tools/ncu_summary.py /tmp/${TAG}_combine.ncu-rep > gpurun_out/${TAG}_prof_combine_cfgC_summary.txt
ncu -i /tmp/${TAG}_combine.ncu-rep --page details --csv 2>/dev/null | grep -E "gin_combine_kernel<4, (16|32)" | grep -E "Memory Throughput|DRAM Throughput|L2 Cache Throughput|Achieved Occupancy|Registers Per|Warp Cycles Per Issued|Theoretical Occupancy" | head -60 > gpurun_out/${TAG}_prof_combine_cfgC_details.csv
fi
if want dense; then
ncu --set full --clock-control none --import-source on -k 'regex:gemm_nt|gemm_tn|thin_bwd|thin_fwd|dz_prepare' --launch-skip 120 -c 16 \
    -o /tmp/${TAG}_dense -f $CMD > gpurun_out/${TAG}_ncu_dense.log 2>&1
python tools/ncu_summary.py /tmp/${TAG}_dense.ncu-rep > gpurun_out/${TAG}_prof_dense_cfgC_summary.txt
fi
ls -la /tmp/*.ncu-rep
rm -f /tmp/*.ncu-rep
ls -la gpurun_out/ | tail -12

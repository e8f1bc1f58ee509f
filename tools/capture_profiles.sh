#!/bin/bash
# Run ON THE GPU BOX (through gpurun): launch list of the default bench command and --set full
# captures of the aggregation and dense-layer kernels of one cfgC step.  Outputs -> gpurun_out/.
# Usage: tools/capture_profiles.sh <tag>
set -u
TAG=${1:-r01}
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/${TAG}_plain.json 2> gpurun_out/${TAG}_plain.err || { tail -5 gpurun_out/${TAG}_plain.err; exit 1; }
# launch list (cold-cache, serialised): ~2 steps after the warm-up steps
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 700 -c 300 --csv \
    --log-file gpurun_out/${TAG}_launches_cfgC.csv $CMD > gpurun_out/${TAG}_ncu_launches.log 2>&1
# full sets: one step's aggregation launches, then its dense-layer launches
ncu --set full --clock-control none --import-source on -k regex:gin_combine --launch-skip 110 -c 22 \
    -o gpurun_out/${TAG}_combine_cfgC -f $CMD > gpurun_out/${TAG}_ncu_combine.log 2>&1
ncu --set full --clock-control none --import-source on -k 'regex:gemm_|dz_prepare|thin_|head_' --launch-skip 250 -c 50 \
    -o gpurun_out/${TAG}_dense_cfgC -f $CMD > gpurun_out/${TAG}_ncu_dense.log 2>&1
python tools/ncu_summary.py gpurun_out/${TAG}_combine_cfgC.ncu-rep > gpurun_out/${TAG}_prof_combine_cfgC_summary.txt
python tools/ncu_summary.py gpurun_out/${TAG}_dense_cfgC.ncu-rep > gpurun_out/${TAG}_prof_dense_cfgC_summary.txt
ls -la gpurun_out/ | tail -12

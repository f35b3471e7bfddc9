#!/bin/bash
# The profiling / measurement set committed under profiles/ (run on a B200 box through gpurun; outputs land in gpurun_out/).
#   tools/evidence.sh <tag>            e.g. tools/evidence.sh r2_final
# 1. ncu --set full of EVERY kernel of one forward + backward (tools/ncu_full_table.py -> per-class table)
# 2. ncu launch list (durations) of one whole eager train step (tools/launch_summary.py, tools/per_layer_table.py)
# 3. ncu DRAM / L2 byte counters of the same step (tools/step_metrics_table.py: whole-step DRAM traffic)
# 4. the default bench line and the reference arm
set -x
TAG=${1:-run}
OUT=gpurun_out
mkdir -p $OUT
SIMQ_GRAPH=0 ncu --profile-from-start off --set full --clock-control none -f -o $OUT/${TAG}_full_fwdbwd python tools/one_step.py --mode fwdbwd > $OUT/${TAG}_ncu_full.log 2>&1
SIMQ_GRAPH=0 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/${TAG}_launches.csv python tools/one_step.py > $OUT/${TAG}_ncu_launches.log 2>&1
SIMQ_GRAPH=0 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__t_bytes.sum --clock-control none --csv --log-file $OUT/${TAG}_step_metrics.csv python tools/one_step.py > $OUT/${TAG}_ncu_metrics.log 2>&1
python bench.py --steps 20 --warmup 5 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
python bench.py --impl reference --steps 20 --warmup 5 > $OUT/${TAG}_bench_reference_arm.json 2> $OUT/${TAG}_bench_reference_arm.err
tail -c 300 $OUT/${TAG}_bench.err; ls -la $OUT | grep ${TAG}

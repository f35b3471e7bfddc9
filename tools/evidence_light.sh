#!/bin/bash
# The cheap half of tools/evidence.sh: bench lines, launch list and byte counters of one step (no section / full-set captures).
set -x
TAG=${1:-run}
OUT=gpurun_out
mkdir -p $OUT
python bench.py --steps 20 --warmup 5 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
python bench.py --impl reference --steps 20 --warmup 5 > $OUT/${TAG}_bench_reference_arm.json 2> $OUT/${TAG}_bench_reference_arm.err
SIMQ_GRAPH=0 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/${TAG}_launches.csv python tools/one_step.py > $OUT/${TAG}_ncu_launches.log 2>&1
SIMQ_GRAPH=0 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__t_bytes.sum --clock-control none --csv --log-file $OUT/${TAG}_step_metrics.csv python tools/one_step.py > $OUT/${TAG}_ncu_metrics.log 2>&1
tail -c 300 $OUT/${TAG}_bench.err; ls -la $OUT | grep ${TAG}; du -sh $OUT

#!/bin/bash
# The profiling / measurement set committed under profiles/ (run on a B200 box through gpurun; outputs land in gpurun_out/).
#   tools/evidence.sh <tag>            e.g. tools/evidence.sh r2_final
# 1. the default bench line and the reference arm
# 2. ncu launch list (durations) of one whole eager train step (tools/launch_summary.py, tools/per_layer_table.py)
# 3. ncu DRAM / L2 byte counters of the same step (tools/step_metrics_table.py: whole-step DRAM traffic)
# 4. ncu SpeedOfLight / MemoryWorkloadAnalysis / ComputeWorkloadAnalysis / LaunchStats / Occupancy sections of EVERY kernel of one
#    forward + backward, and `--set full` of one launch of the dominant kernels; the .ncu-rep files are converted to raw CSV on the
#    box and deleted (gpurun_out/ only travels back below 64 MiB)
set -x
TAG=${1:-run}
OUT=gpurun_out
mkdir -p $OUT
python bench.py --steps 20 --warmup 5 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
python bench.py --impl reference --steps 20 --warmup 5 > $OUT/${TAG}_bench_reference_arm.json 2> $OUT/${TAG}_bench_reference_arm.err
SIMQ_GRAPH=0 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/${TAG}_launches.csv python tools/one_step.py > $OUT/${TAG}_ncu_launches.log 2>&1
SIMQ_GRAPH=0 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__t_bytes.sum --clock-control none --csv --log-file $OUT/${TAG}_step_metrics.csv python tools/one_step.py > $OUT/${TAG}_ncu_metrics.log 2>&1
SIMQ_GRAPH=0 ncu --profile-from-start off --section SpeedOfLight --section MemoryWorkloadAnalysis --section ComputeWorkloadAnalysis --section LaunchStats --section Occupancy \
    --metrics l1tex__m_xbar2l1tex_read_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -f -o /tmp/${TAG}_sections python tools/one_step.py --mode fwdbwd > $OUT/${TAG}_ncu_sections.log 2>&1
ncu -i /tmp/${TAG}_sections.ncu-rep --page raw --csv > $OUT/${TAG}_sections_raw.csv 2>> $OUT/${TAG}_ncu_sections.log
SIMQ_GRAPH=0 ncu --profile-from-start off --set full --clock-control none -k 'regex:conv2w_umma_kernel|wgrad2_umma_kernel|bn_bwd_apply_kernel|bn_apply_kernel|bn_bwd_reduce_kernel' -s 20 -c 14 -f -o /tmp/${TAG}_full python tools/one_step.py --mode fwdbwd > $OUT/${TAG}_ncu_full.log 2>&1
ncu -i /tmp/${TAG}_full.ncu-rep --page raw --csv > $OUT/${TAG}_full_raw.csv 2>> $OUT/${TAG}_ncu_full.log
rm -f /tmp/${TAG}_sections.ncu-rep /tmp/${TAG}_full.ncu-rep
tail -c 300 $OUT/${TAG}_bench.err; ls -la $OUT | grep ${TAG}; du -sh $OUT

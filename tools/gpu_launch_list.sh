#!/bin/bash
# ncu launch list of one timed cycle of the default bench workload (frozen regime): per-launch gpu__time_duration of the
# ~31.5k kernels of a cycle (150 rollout steps + one PPO update) after the warm-up cycles are skipped.
#   gpurun --timeout 2400 -- 'bash tools/gpu_launch_list.sh'
# then here:  python tools/launch_summary.py gpurun_out/launches_r02.csv > profiles/r02_launch_summary.txt
set -x
mkdir -p gpurun_out
args="--steps 1 --warmup 3 --no-cpu --no-eager --no-shares --no-e2e --regime frozen"
timeout 600 python bench.py $args > gpurun_out/launch_list_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/launch_list_plain.log; exit 1; }
grep '{"metric' gpurun_out/launch_list_plain.log | cut -c1-200
skip=${1:-97500}
# (ncu profiles ~9 launches/s here: 31.6k launches do not fit a 2000 s window - the round-2 list stopped at 17.7k)
timeout 2000 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip $skip -c ${2:-16000} --csv --log-file gpurun_out/launches_r02.csv python bench.py $args > gpurun_out/launch_list_ncu.log 2>&1
echo "ncu rc=$?"; wc -l gpurun_out/launches_r02.csv; tail -3 gpurun_out/launch_list_ncu.log | cut -c1-300
gzip -kf gpurun_out/launches_r02.csv

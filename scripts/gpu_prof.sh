# usage: bash scripts/gpu_prof.sh <tag>     (tape_cache=0 so that every launch logs its tape shape, in launch order)
TAG=${1:-r1k}
set -x
export FMC_OPTIONS=tape_cache=0
FMC_LOG_TAPES=1 timeout -s KILL 300 python benchmarks/profile_cases.py > gpurun_out/prof_plain.log 2> gpurun_out/prof_tapes.log && \
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$TAG.csv python benchmarks/profile_cases.py > gpurun_out/ncu_launch.log 2>&1 && \
timeout -s KILL 1200 ncu --set full --clock-control none --import-source on -k regex:"tape_kernel|reduce_kernel" -s 0 -c 12 -o gpurun_out/prof_micro_$TAG -f python benchmarks/profile_cases.py > gpurun_out/ncu_full1.log 2>&1 && \
timeout -s KILL 1200 ncu --set full --clock-control none --import-source on -k regex:tape_kernel -s 150 -c 3 -o gpurun_out/prof_swaption_$TAG -f python benchmarks/profile_cases.py > gpurun_out/ncu_full2.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu_full2.log

# usage: bash scripts/gpu_evidence_r3.sh <tag>   — the ncu passes of round 2's final state (windows of time steps): launch list of
# bench.py and full captures of window kernels and swaption kernels; each pass after the same command has exited 0 without ncu.
TAG=${1:-r3}
BENCH="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-calibration --no-extras"
set -x
timeout -s KILL 600 $BENCH > gpurun_out/evidence_plain_$TAG.json 2> gpurun_out/evidence_plain_$TAG.err || { echo "plain run failed"; tail -5 gpurun_out/evidence_plain_$TAG.err; exit 1; }
timeout -s KILL 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_$TAG.csv $BENCH > gpurun_out/ncu_launches_$TAG.log 2>&1
echo "launches rc=$?"
timeout -s KILL 1200 ncu --set full --clock-control none --import-source on -k regex:tape_kernel -s 173 -c 3 -o gpurun_out/prof_sim_$TAG -f $BENCH > gpurun_out/ncu_sim_$TAG.log 2>&1
echo "sim rc=$?"
timeout -s KILL 1200 ncu --set full --clock-control none --import-source on -k regex:tape_kernel -s 212 -c 6 -o gpurun_out/prof_swaption_$TAG -f $BENCH > gpurun_out/ncu_swaption_$TAG.log 2>&1
echo "swaption rc=$?"

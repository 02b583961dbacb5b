# usage: bash scripts/gpu_evidence.sh <launches|sim|swaption|brownian> <tag>
# One ncu pass per gpurun call, each after the same command has exited 0 without ncu.
WHAT=$1; TAG=${2:-r2}
BENCH="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-calibration --no-extras"
set -x
timeout -s KILL 600 $BENCH > gpurun_out/evidence_plain_$TAG.json 2> gpurun_out/evidence_plain_$TAG.err || { echo "plain run failed"; tail -5 gpurun_out/evidence_plain_$TAG.err; exit 1; }
case $WHAT in
launches) timeout -s KILL 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$TAG.csv $BENCH > gpurun_out/ncu_launches_$TAG.log 2>&1 ;;
sim)      timeout -s KILL 1200 ncu --set full --clock-control none --import-source on -k regex:tape_kernel -s 234 -c 3 -o gpurun_out/prof_sim_$TAG -f $BENCH > gpurun_out/ncu_sim_$TAG.log 2>&1 ;;
swaption) timeout -s KILL 1200 ncu --set full --clock-control none --import-source on -k regex:tape_kernel -s 304 -c 6 -o gpurun_out/prof_swaption_$TAG -f $BENCH > gpurun_out/ncu_swaption_$TAG.log 2>&1 ;;
brownian) timeout -s KILL 1200 ncu --set full --clock-control none --import-source on -k regex:brownian_kernel -c 1 -o gpurun_out/prof_brownian_$TAG -f $BENCH > gpurun_out/ncu_brownian_$TAG.log 2>&1 ;;
esac
echo "ncu rc=$?"

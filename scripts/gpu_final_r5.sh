# final state of round 2 (session 3): smoke, the default bench line and the reference arm as the driver runs them, the launch list of the
# bench under ncu (after the plain command has exited 0), the calibration and configs benchmarks
set -x
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout -s KILL 1200 python bench.py > gpurun_out/bench_r5_final.json 2> gpurun_out/bench_r5_final.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/bench_r5_final.json
timeout -s KILL 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r5_reference.json 2>/dev/null; cut -c1-300 gpurun_out/bench_r5_reference.json
BENCH="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-calibration --no-extras"
timeout -s KILL 600 $BENCH > gpurun_out/evidence_plain_r5.json 2> gpurun_out/evidence_plain_r5.err || { echo "plain run failed"; tail -5 gpurun_out/evidence_plain_r5.err; exit 1; }
timeout -s KILL 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r5.csv $BENCH > gpurun_out/ncu_launches_r5.log 2>&1
echo "launches rc=$?"
timeout -s KILL 300 python benchmarks/brownian_rate.py 2>&1 | tail -1
timeout -s KILL 600 python benchmarks/configs.py 2>&1 | tail -6 | cut -c1-300

set -x
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout -s KILL 1200 python bench.py > gpurun_out/bench_r4_final.json 2> gpurun_out/bench_r4_final.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/bench_r4_final.json
timeout -s KILL 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r4_reference.json 2>/dev/null; cut -c1-300 gpurun_out/bench_r4_reference.json
timeout -s KILL 300 python benchmarks/latency.py > gpurun_out/latency_r4_final.log 2>&1; tail -9 gpurun_out/latency_r4_final.log

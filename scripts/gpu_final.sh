set -x
timeout -s KILL 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout -s KILL 900 python bench.py > gpurun_out/bench_r1_final.json 2> gpurun_out/bench_r1_final.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/bench_r1_final.json
timeout -s KILL 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r1_reference.json 2>/dev/null; cut -c1-200 gpurun_out/bench_r1_reference.json
timeout -s KILL 1200 python benchmarks/raw_ops.py --sizes 10000,100000,1000000,10000000,100000000 --out gpurun_out/raw_ops_r1_final.json > gpurun_out/raw_ops_r1_final.log 2>&1; echo "raw rc=$?"; grep -c "B" gpurun_out/raw_ops_r1_final.log
timeout -s KILL 300 python benchmarks/latency.py > gpurun_out/latency_r1_final.log 2>&1; tail -9 gpurun_out/latency_r1_final.log

# HEAD of the round: the whole GPU suite, smoke, the default bench line
timeout -s KILL 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout -s KILL 1200 python bench.py > gpurun_out/bench_r5_final.json 2> gpurun_out/bench_r5_final.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/bench_r5_final.json

nproc
timeout -s KILL 900 python bench.py > gpurun_out/bench_r3c.json 2> gpurun_out/bench_r3c.err; echo "bench rc=$?"
timeout -s KILL 600 python bench.py --impl reference > gpurun_out/bench_r3c_ref.json 2> gpurun_out/bench_r3c_ref.err; echo "ref rc=$?"
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2

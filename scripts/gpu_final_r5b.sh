# last build of the round: full ncu captures of the window and swaption kernels (tag r5), then the default bench line and the reference arm
bash scripts/gpu_evidence_r3.sh r5 2>&1 | grep -v "^+" | tail -8
timeout -s KILL 1200 python bench.py > gpurun_out/bench_r5_final.json 2> gpurun_out/bench_r5_final.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/bench_r5_final.json
timeout -s KILL 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r5_reference.json 2>/dev/null; cut -c1-200 gpurun_out/bench_r5_reference.json

# round 2, first GPU call: the three chunk geometries of the interpreter (parity + where the LMM step spends its time)
set -x
nvidia-smi -L
timeout -s KILL 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
for e in 16 8 4; do
  FMC_TEST_OPTIONS=tape_elems=$e timeout -s KILL 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_workloads.py -m gpu -x -q 2>&1 | tail -4
done
for e in 16 8 4; do
  FMC_OPTIONS=tape_elems=$e timeout -s KILL 300 python benchmarks/lmm_phases.py > gpurun_out/phases_e$e.log 2>&1; tail -9 gpurun_out/phases_e$e.log
done
timeout -s KILL 300 python benchmarks/lmm_phases.py > gpurun_out/phases_auto.log 2>&1; tail -9 gpurun_out/phases_auto.log
timeout -s KILL 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err; echo "bench rc=$?"; cut -c1-600 gpurun_out/bench_r2a.json

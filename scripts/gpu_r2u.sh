for p in 10000 100000; do
  timeout -s KILL 300 python benchmarks/lmm_phases.py $p 2>&1 | grep -E "kernels:|full step|simulate:|swaption" | tail -6
done
FMC_PROFILE_DUMP=gpurun_out/launch_dump_10k.txt timeout -s KILL 300 python benchmarks/lmm_phases.py 10000 > /dev/null 2>&1

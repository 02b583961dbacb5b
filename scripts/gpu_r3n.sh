nproc
for t in 8 4 12 16; do echo "FMC_HOST_THREADS=$t"; FMC_HOST_THREADS=$t timeout -s KILL 300 python benchmarks/e2e_phases.py 2>&1 | tail -1; done

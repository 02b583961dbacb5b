# host worker pool of the upload path: blocks per thread and spinning workers, end-to-end step (host double[] every step)
run() { timeout -s KILL 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-calibration --no-extras 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('ms_per_step',round(d['ms_per_step'],3),'e2e',round(d['e2e']['ms_per_step'],2),'e2e_pinned',round(d['e2e_pinned']['ms_per_step'],2))"; }
nproc
for rep in 1 2; do
for cfg in "1 0 8" "4 0 8" "4 200 8" "4 1000 8" "8 200 8" "4 200 12" "4 200 16" "4 200 6"; do set -- $cfg
echo "== blocks_per_thread=$1 spin_us=$2 threads=$3"; FMC_HOST_BLOCKS_PER_THREAD=$1 FMC_HOST_SPIN_US=$2 FMC_HOST_THREADS=$3 run; done; done

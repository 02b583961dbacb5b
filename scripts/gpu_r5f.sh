# upload path: 128-bit against 256-bit conversion loop, end-to-end step (host double[] every step); then the GPU tests that upload and download
run() { timeout -s KILL 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-calibration --no-extras 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('ms_per_step',round(d['ms_per_step'],3),'e2e',round(d['e2e']['ms_per_step'],2),'e2e_pinned',round(d['e2e_pinned']['ms_per_step'],2))"; }
nproc; grep -m1 "model name" /proc/cpuinfo
for rep in 1 2; do
for cfg in "0 200" "1 200" "1 50" "1 500"; do set -- $cfg
echo "== avx=$1 spin_us=$2"; FMC_HOST_AVX=$1 FMC_HOST_SPIN_US=$2 run; done; done
timeout -s KILL 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2

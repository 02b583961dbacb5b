set -x
timeout -s KILL 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
for o in "" "target_ctas=6" "target_ctas=3" "cta_warps=2" "fuse_ops=0" "ring_min=2,target_ctas=6"; do
  echo "== FMC_OPTIONS=$o"
  FMC_OPTIONS="$o" timeout -s KILL 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('ms_per_step',round(d['ms_per_step'],2),'kernel_ms',round(d['roofline']['kernel_ms_per_step'],2),'launches',d['gpu_launches_per_step'],'e2e_ms',round(d['e2e']['ms_per_step'],1))"
done
timeout -s KILL 600 python benchmarks/raw_ops.py --sizes 1048576,67108864 --out gpurun_out/raw_ops_r1d.json > gpurun_out/raw_ops_r1d.log 2>&1; echo "raw rc=$?"; grep "67108864" gpurun_out/raw_ops_r1d.log

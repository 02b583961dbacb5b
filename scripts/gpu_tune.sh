# tuning sweep of interpreter scheduling knobs on the headline bench (device-resident arm only)
for o in "" "target_ctas=8,ring_min=4" "target_ctas=8,ring_min=3" "target_ctas=8,ring_min=2" "target_ctas=6,ring_min=4" "target_ctas=2" "target_ctas=8,ring_min=2,pipeline=0"; do
  echo "== FMC_OPTIONS=$o"
  FMC_OPTIONS="$o" timeout -s KILL 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('ms_per_step',round(d['ms_per_step'],2),'kernel_ms',round(d['roofline']['kernel_ms_per_step'],2),'launches',d['gpu_launches_per_step'],'e2e_ms',round(d['e2e']['ms_per_step'],1))"
done
for o in "" "max_sets=1" "max_sets=2"; do
  echo "== raw FMC_OPTIONS=$o"
  FMC_OPTIONS="$o" timeout -s KILL 300 python benchmarks/raw_ops.py --sizes 67108864 --cases b1,b2,b3 --out gpurun_out/raw_tune.json 2>&1 | grep -E "add\(scalar\)|add\(vec\)|BS-Euler|getAverage  |getVariance|payoff" 
done

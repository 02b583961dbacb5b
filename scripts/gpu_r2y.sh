for rep in 1 2; do
for opts in "flush_threshold=2048" "flush_threshold=3072" "flush_threshold=4096" "flush_threshold=5120"; do
  FMC_OPTIONS=$opts timeout -s KILL 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-calibration --no-extras 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('$opts', 'ms_per_step',round(d['ms_per_step'],3),'kernel',round(d['roofline']['kernel_ms_per_step'],3),'launches',d['roofline']['launches'],'e2e',round(d['e2e']['ms_per_step'],2), 'algGB', round(d['roofline']['algorithmic_bytes_per_step']/1e9,2))"
done
done

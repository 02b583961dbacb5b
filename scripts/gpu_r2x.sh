for rep in 1 2 3; do
for opts in "flush_threshold=4096" "flush_threshold=6144"; do
  FMC_OPTIONS=$opts timeout -s KILL 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-calibration --no-extras 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('$opts', 'ms_per_step',round(d['ms_per_step'],3),'kernel',round(d['roofline']['kernel_ms_per_step'],3),'e2e',round(d['e2e']['ms_per_step'],2), d['host_profile'])"
done
done

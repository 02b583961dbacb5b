for tp in 0 1; do
  if [ $tp = 1 ]; then export FMD_TWO_PASS=1; fi
  timeout -s KILL 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-calibration --no-extras 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('two_pass=$tp', 'ms_per_step',round(d['ms_per_step'],3),'kernel',round(d['roofline']['kernel_ms_per_step'],3),'launches',d['roofline']['launches'],'e2e',round(d['e2e']['ms_per_step'],2), 'algGB', round(d['roofline']['algorithmic_bytes_per_step']/1e9,2), d['host_profile'], d.get('parity'))"
done

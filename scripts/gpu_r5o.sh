# flush threshold after the host mirror change (the simulation phase is host-bound: an earlier first flush starts the GPU earlier)
for o in flush_threshold=4096 flush_threshold=3072 flush_threshold=5120 flush_threshold=4096; do echo "== $o"
FMC_OPTIONS=$o timeout -s KILL 100 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-calibration --no-extras 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('ms_per_step',round(d['ms_per_step'],3),'kernel',round(d['roofline']['kernel_ms_per_step'],3),'launches',d.get('gpu_launches'))"; done

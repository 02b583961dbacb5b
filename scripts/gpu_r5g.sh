# host mirror: recycled object blocks and borrowed operand references — bench step, path sweep floor, calibration, the workload / AAD tests
timeout -s KILL 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('ms_per_step',round(d['ms_per_step'],3),'kernel',round(d['roofline']['kernel_ms_per_step'],3),'frac',round(d['roofline']['frac'],4),'e2e',round(d['e2e']['ms_per_step'],2),'pp',round(d['price_products_step']['ms_per_step'],2),'launches',d.get('gpu_launches')); print(d['host_profile']); print(d['extras']['path_sweep_ms_per_step']); print(d['extras']['bermudan_1m_paths_sharded']); c=d['calibration']; print(c['seconds_per_evaluation'], c['t_atm_10k_paths_2_iterations'], c['t_atm_10k_paths_to_convergence'])"
for p in 10000 1048576; do timeout -s KILL 300 python benchmarks/lmm_phases.py $p 2>&1 | grep -A3 "timing off" | tail -3; done
timeout -s KILL 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2

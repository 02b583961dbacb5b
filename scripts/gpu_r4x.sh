for o in window_levels=3; do echo "== $o"; FMC_OPTIONS=$o timeout -s KILL 600 python benchmarks/aad_footprint.py 2>&1 | tail -4 | cut -c1-300; done
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-calibration 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('ms_per_step',round(d['ms_per_step'],3),'kernel',round(d['roofline']['kernel_ms_per_step'],3),'frac',round(d['roofline']['frac'],4), d['extras']['bermudan_1m_paths_sharded'], d['extras']['path_sweep_ms_per_step'])"
python -m pytest tests/test_gpu_aad.py tests/test_gpu_workloads.py -m gpu -x -q 2>&1 | tail -2

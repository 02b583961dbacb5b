for o in "max_regs=8" "max_regs=6" "max_regs=4"; do
echo "== $o"
FMC_OPTIONS="$o" timeout -s KILL 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('ms/step', round(d['ms_per_step'],2), 'e2e ms', round(d['e2e']['ms_per_step'],2), 'kernel ms', round(d['roofline']['kernel_ms_per_step'],2), 'launches/step', d['gpu_launches_per_step'])"
FMC_OPTIONS="$o" timeout -s KILL 300 python benchmarks/configs.py 2>&1 | grep -E "config 3|config 4" | cut -c1-170
FMC_OPTIONS="$o" timeout -s KILL 600 python benchmarks/raw_ops.py --sizes 100000000 --cases b2 --out gpurun_out/raw_tmp.json 2>&1 | grep -E "chain of 16|chain of 64|LMM comp" | grep -v REF | cut -c1-120
done

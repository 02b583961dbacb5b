nproc; lscpu | grep -E "Model name|Socket|NUMA node\(s\)" 
for t in 8 16 32 64; do
echo "== FMC_HOST_THREADS=$t"
FMC_HOST_THREADS=$t timeout -s KILL 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('ms_per_step',round(d['ms_per_step'],2),'e2e_ms',round(d['e2e']['ms_per_step'],1))"
done
timeout -s KILL 600 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | cut -c1-400

timeout -s KILL 300 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
timeout -s KILL 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 benchmarks/multi_gpu_check.py 2>&1 | grep -v "^\*\|OMP_NUM" | tail -16
for o in "p2p_reduce=1" "p2p_reduce=0"; do
FMC_OPTIONS=$o timeout -s KILL 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "$o rc=$?"; python -c "
import json
s=open('gpurun_out/bench_2gpu.json').read(); d=json.loads(s[s.index('{'):]); print('2gpu ms_per_step',round(d['ms_per_step'],2),'value',d['value'],d['host_profile'])"
done
timeout -s KILL 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('1gpu ms_per_step',round(d['ms_per_step'],2),'value',d['value'])"

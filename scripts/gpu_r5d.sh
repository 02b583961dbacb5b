# multi-GPU sanity of the round's final code, as the driver launches it: bench.py under torchrun with default flags, the reference arm, parity check
N=$1
nproc
timeout -s KILL 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 benchmarks/multi_gpu_check.py 2>&1 | grep -v "^\*\|OMP_NUM" | tail -3
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_${N}gpu_r5.json 2> gpurun_out/bench_${N}gpu_r5.err; echo "rc=$?"; python -c "
import json,sys
s=open('gpurun_out/bench_${N}gpu_r5.json').read(); d=json.loads(s[s.index('{'):]); print('${N}gpu ms_per_step',round(d['ms_per_step'],2),'value',d['value'],'e2e',round(d['e2e']['ms_per_step'],1),d['host_profile'],'clocks',d['clocks'], 'parity', d.get('multi_gpu_parity'))"
timeout -s KILL 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-calibration --no-extras 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('1gpu on this box: ms_per_step',round(d['ms_per_step'],2),'value',d['value'],'e2e',round(d['e2e']['ms_per_step'],1))"

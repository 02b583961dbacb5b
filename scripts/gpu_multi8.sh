N=$1
nproc
timeout -s KILL 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 benchmarks/multi_gpu_check.py 2>&1 | grep -v "^\*\|OMP_NUM" | tail -2
timeout -s KILL 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; echo "rc=$?"; python -c "
import json,sys
s=open('gpurun_out/bench_${N}gpu.json').read(); d=json.loads(s[s.index('{'):]); print('${N}gpu ms_per_step',round(d['ms_per_step'],2),'value',d['value'],'e2e',round(d['e2e']['ms_per_step'],1),d['host_profile'],'clocks',d['clocks'])"
timeout -s KILL 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus $N --steps 2 --warmup 1 2>/dev/null | cut -c1-200

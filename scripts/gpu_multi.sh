timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 benchmarks/multi_gpu_check.py 2>&1 | grep -v "^\*\|OMP_NUM" | tail -10
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/bench_2gpu.json').read()); print('2gpu ms_per_step',d['ms_per_step'],'value',d['value'],d['host_profile'])"
timeout -s KILL 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('1gpu ms_per_step',round(d['ms_per_step'],2),'value',d['value'],'kernel_ms',round(d['roofline']['kernel_ms_per_step'],2), d['host_profile'])"

# round 2, two GPUs: the three exchange modes of a sharded reduction (2 shared host memory [default], 1 in-kernel peer memory, 0 NCCL)
for o in "exchange=2" "exchange=1" "exchange=0"; do
  echo "== $o"
  FMC_OPTIONS=$o timeout -s KILL 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 benchmarks/multi_gpu_check.py 2>&1 | grep -v "^\*\|OMP_NUM" | tail -3
  FMC_OPTIONS=$o timeout -s KILL 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 --no-calibration > gpurun_out/bench_2gpu_$o.json 2> gpurun_out/bench_2gpu_$o.err; echo "$o rc=$?"; tail -2 gpurun_out/bench_2gpu_$o.err; python -c "
import json,sys
s=open('gpurun_out/bench_2gpu_$o.json').read(); d=json.loads(s[s.index('{'):]); print('2gpu ms_per_step',round(d['ms_per_step'],2),'e2e',round(d['e2e']['ms_per_step'],2),'kernel',round(d['roofline']['kernel_ms_per_step'],2),d['host_profile'], d.get('multi_gpu_parity'))"
done
timeout -s KILL 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-calibration 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('1gpu ms_per_step',round(d['ms_per_step'],2),'e2e',round(d['e2e']['ms_per_step'],2),'kernel',round(d['roofline']['kernel_ms_per_step'],2),d['host_profile'])"

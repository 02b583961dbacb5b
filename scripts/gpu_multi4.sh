N=$1
for o in "p2p_reduce=1" "p2p_reduce=0" "p2p_reduce=1,zero_copy_reduce=1"; do
FMC_HOST_THREADS=2 FMC_OPTIONS=$o timeout -s KILL 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; echo "$o rc=$?"; python -c "
import json,sys
s=open('gpurun_out/bench_${N}gpu.json').read(); d=json.loads(s[s.index('{'):]); print('${N}gpu ms_per_step',round(d['ms_per_step'],2),'value',d['value'],'e2e',round(d['e2e']['ms_per_step'],1),d['host_profile'],'kernel_ms',round(d['roofline']['kernel_ms_per_step'],2))"
done
nproc; uptime

# 2-GPU device-timed step, three repetitions (variance between runs), then one GPU on the same box
nproc
for rep in 1 2 3; do
timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$rep bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline --no-calibration --no-extras 2>/dev/null | python -c "
import sys,json
s=sys.stdin.read(); d=json.loads(s[s.index('{'):]); print('2gpu ms_per_step',round(d['ms_per_step'],2),'e2e',round(d['e2e']['ms_per_step'],1),d['host_profile'])"
done
for rep in 1 2; do timeout -s KILL 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-calibration --no-extras 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('1gpu ms_per_step',round(d['ms_per_step'],2),'e2e',round(d['e2e']['ms_per_step'],1),d['host_profile'])"; done

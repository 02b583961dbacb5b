timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 3 --no-calibration --no-extras 2>/dev/null | python -c "
import sys,json
s=sys.stdin.read(); d=json.loads(s[s.index('{'):]); print('2gpu ms_per_step',round(d['ms_per_step'],2),'e2e',round(d['e2e']['ms_per_step'],1),'parity',d.get('multi_gpu_parity'))"

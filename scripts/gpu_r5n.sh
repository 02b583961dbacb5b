# 4 ranks on one host with the final host pool defaults: device-timed and end-to-end step, parity against the oracle
nproc
timeout -s KILL 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 4 --steps 10 --warmup 3 --no-calibration --no-extras 2>/dev/null | python -c "
import sys,json
s=sys.stdin.read(); d=json.loads(s[s.index('{'):]); print('4gpu ms_per_step',round(d['ms_per_step'],2),'value',d['value'],'e2e',round(d['e2e']['ms_per_step'],1),'parity',(d.get('multi_gpu_parity') or {}).get('max_rel_diff'))"

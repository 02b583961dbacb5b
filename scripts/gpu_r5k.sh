# last build: smoke, 1-GPU and 2-GPU step with default pool settings
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout -s KILL 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-calibration --no-extras 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('1gpu ms_per_step',round(d['ms_per_step'],2),'e2e',round(d['e2e']['ms_per_step'],1))"
timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline --no-calibration --no-extras 2>/dev/null | python -c "
import sys,json
s=sys.stdin.read(); d=json.loads(s[s.index('{'):]); print('2gpu ms_per_step',round(d['ms_per_step'],2),'e2e',round(d['e2e']['ms_per_step'],1),'parity',d.get('multi_gpu_parity'))"
timeout -s KILL 300 python -m pytest tests -m gpu -x -q -k "upload or roundtrip or pool or staging or realizations" 2>&1 | tail -1

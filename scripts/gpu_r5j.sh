# 2 ranks on one host: host pool settings for the end-to-end step
nproc
for cfg in "8 0" "12 0" "12 500" "8 500"; do set -- $cfg
echo "== threads=$1 spin_us=$2"
FMC_HOST_THREADS=$1 FMC_HOST_SPIN_US=$2 timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline --no-calibration --no-extras 2>/dev/null | python -c "
import sys,json
s=sys.stdin.read(); d=json.loads(s[s.index('{'):]); print('2gpu ms_per_step',round(d['ms_per_step'],2),'e2e',round(d['e2e']['ms_per_step'],1),'pinned',round(d['e2e_pinned']['ms_per_step'],1))"
done

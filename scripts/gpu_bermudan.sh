N=$1
timeout -s KILL 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 benchmarks/multi_gpu_check.py 2>&1 | grep -v "^\*\|OMP_NUM" | tail -8
timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 benchmarks/bermudan_sharded.py 2>&1 | grep -v "^\*\|OMP_NUM" | tail -2
timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 benchmarks/bermudan_sharded.py 8000000 2>&1 | grep -v "^\*\|OMP_NUM" | tail -1

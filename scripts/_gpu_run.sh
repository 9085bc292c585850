set -x
N=$1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 30 --warmup 5 2>&1 | tail -1 > gpurun_out/bench$N.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 30 --warmup 5 --shard rows 2>&1 | tail -1 > gpurun_out/bench${N}r.log
timeout 300 python bench.py --no-cpu 2>&1 | tail -1 > gpurun_out/bench1q.log

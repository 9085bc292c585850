# 8-GPU runs: the driver's bench line (weak-scaled configs[1] + the north-star target), then reduce-parts variants of the target
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
timeout 420 $TR bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_8gpu.json 2> gpurun_out/bench_8gpu.err
for parts in 2 4; do
  LMDEC_REDUCE_PARTS=$parts timeout 200 $TR bench.py --gpus 8 --target-only > gpurun_out/target_8gpu_parts$parts.json 2> gpurun_out/target_8gpu_parts$parts.err
done
tail -c 1500 gpurun_out/bench_8gpu.json; tail -c 400 gpurun_out/bench_8gpu.err

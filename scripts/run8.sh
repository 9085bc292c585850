# 8-GPU runs: the driver's bench line (weak-scaled configs[1] + the north-star target with the fused reduction), then the
# target alone over NCCL for comparison
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_8gpu.json 2> gpurun_out/bench_8gpu.err; echo "bench rc=$?"
LMDEC_FUSED_REDUCE=0 timeout 150 $TR bench.py --gpus 8 --target-only > gpurun_out/target_8gpu_nccl.json 2> gpurun_out/target_8gpu_nccl.err; echo "nccl rc=$?"
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/bench_8gpu.json").read().strip().splitlines()[-1]); t=d["target"]
    print("bench", d["value"], d["ms_per_step"], "target fused: it", t["iterations"], "ms/iter", t["ms_per_iteration"], "s", t["seconds_to_converged"], "tflops", t["tflops"], "err", t["oracle_slab_check"]["max_err_AtU_product_vs_oracle"])
except Exception as e: print("bench failed", e)
try:
    t=json.loads(open("gpurun_out/target_8gpu_nccl.json").read().strip().splitlines()[-1])["target"]
    print("target nccl: it", t["iterations"], "ms/iter", t["ms_per_iteration"], "s", t["seconds_to_converged"], "tflops", t["tflops"])
except Exception as e: print("nccl failed", e)
PY
tail -c 600 gpurun_out/bench_8gpu.err

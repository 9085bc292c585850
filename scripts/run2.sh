TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
LMDEC_FUSED_REDUCE=require timeout 250 $TR tests/dist_gpu_check.py > gpurun_out/dist2_fused.log 2>&1; echo "dist check rc=$?"; grep -v "^\*\|OMP_NUM" gpurun_out/dist2_fused.log | tail -12
timeout 200 $TR bench.py --gpus 2 --target-only > gpurun_out/target_2gpu_fused.json 2> gpurun_out/target_2gpu_fused.err; echo "fused rc=$?"
LMDEC_FUSED_REDUCE=0 timeout 200 $TR bench.py --gpus 2 --target-only > gpurun_out/target_2gpu_nccl.json 2> gpurun_out/target_2gpu_nccl.err; echo "nccl rc=$?"
python - <<'PY'
import json
for f in ("fused","nccl"):
    try:
        t=json.loads(open(f"gpurun_out/target_2gpu_{f}.json").read().strip().splitlines()[-1])["target"]
        print(f, t["iterations"], "ms/iter", t["ms_per_iteration"], "s", t["seconds_to_converged"], "prod err", t["oracle_slab_check"]["max_err_AtU_product_vs_oracle"], "S0", t["S"][0])
    except Exception as e:
        print(f, "failed", e)
PY
tail -5 gpurun_out/target_2gpu_fused.err

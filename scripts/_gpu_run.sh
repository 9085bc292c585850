set -x
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -5 > gpurun_out/t1.log
LMDEC_LIB=scratch/lib_trace.so timeout 300 python tools/trace_cqr.py > gpurun_out/trace_cqr.txt 2>&1
for i in 1 2; do
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu 2>&1 | tail -1 >> gpurun_out/ab_new.log
done

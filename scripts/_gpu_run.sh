set -x
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -40 > gpurun_out/t1.log
LMDEC_ISSUER_GROUPS=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/l_g1.csv python bench.py --steps 3 --warmup 3 --no-cpu --profile-range > gpurun_out/ncu1.log 2>&1
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/l_g2.csv python bench.py --steps 3 --warmup 3 --no-cpu --profile-range > gpurun_out/ncu2.log 2>&1
for i in 1 2; do
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu 2>&1 | tail -1 >> gpurun_out/ab_new.log
done

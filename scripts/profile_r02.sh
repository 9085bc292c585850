# Round-2 profiles: bench line (no profiler), ncu launch list of the same command, --set full captures of the two
# product kernels.  Run under gpurun; everything lands in gpurun_out/.
set -x
timeout 280 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02_final.json 2> gpurun_out/bench_r02_final.err || exit 1
timeout 250 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file gpurun_out/launches_r02.csv python bench.py --steps 2 --warmup 3 --no-target --no-cpu --profile-range \
  > gpurun_out/ncu_launches.log 2>&1
timeout 280 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:geno_mma -c 2 \
  -f -o gpurun_out/geno_mma_r02 python bench.py --steps 2 --warmup 3 --no-target --no-cpu --profile-range \
  > gpurun_out/ncu_geno.log 2>&1
CUBLAS=0 timeout 200 ncu --set full --clock-control none --import-source on -k regex:dense_mma -s 4 -c 2 \
  -f -o gpurun_out/dense_mma_r02 python tools/time_dense.py 200000 4000 110 1 > gpurun_out/ncu_dense.log 2>&1
ls -la gpurun_out/*.ncu-rep

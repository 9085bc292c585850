# usage: bash scripts/ab_run.sh name1 name2 ...   (libraries ab/lib_<name>.so; results in gpurun_out/ab_<name>.log)
for v in "$@"; do
  LMDEC_LIB=ab/lib_$v.so timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu 2>&1 | tail -1 > gpurun_out/ab_$v.log
done

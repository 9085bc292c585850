# A/B of producer-warp counts / producer groups: bash scripts/ab_groups.sh  (libraries under ab/, output gpurun_out/ab_groups.log)
out=gpurun_out/ab_groups.log; : > $out
run() { # lib gm g shape...
  lib=$1; gm=$2; g=$3; shift 3
  echo "lib=$lib groups_missing=$gm groups=$g" >> $out
  LMDEC_LIB=ab/lib_$lib.so LMDEC_PRODUCER_GROUPS_MISSING=$gm LMDEC_PRODUCER_GROUPS=$g timeout 120 python tools/time_products.py "$@" 2>&1 | tail -1 >> $out
}
for cfg in "base 1 2" "base 2 2" "e4 1 2" "p16e4 1 2" "p16e4 2 2" "p16e4 4 4" "p16e8 1 2" "p16e8 2 2" "p16e8 4 4"; do
  set -- $cfg
  run $1 $2 $3
  run $1 $2 $3 12500 500000 20 0.0 10
done

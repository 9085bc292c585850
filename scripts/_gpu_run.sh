set -x
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -15 > gpurun_out/t1.log

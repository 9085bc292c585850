"""bench.py contract pieces that can be checked without a GPU: the reference arm (CPU oracle port) prints one JSON line
with the keys the driver reads, and the workload description names the decomposition that was run."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_workload_config_names_the_decomposition():
    sys.path.insert(0, ROOT)
    import bench
    one = bench.workload_config(1)
    assert one["global_shape"] == [bench.N_ROWS, bench.N_COLS] and one["parallelism"] == "single GPU"
    snp = bench.workload_config(8, "snps")
    assert snp["global_shape"] == [bench.N_ROWS, 8 * bench.N_COLS] and "SNPs sharded x8" in snp["parallelism"]
    rows = bench.workload_config(8, "rows")
    assert rows["global_shape"] == [8 * bench.N_ROWS, bench.N_COLS] and "row-sharded x8" in rows["parallelism"]
    assert "model" not in one and "workload" in one


@pytest.mark.timeout(600)
def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=580, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "TFLOP/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["ms_per_step"] > 0 and line["n_gpus"] == 1
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and "sample" in cb
    e2e = line["e2e"]
    assert e2e["value"] == line["value"] and e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0
    assert line["config"]["workload"].startswith("configs[1]")

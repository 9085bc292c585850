"""Build A/B variants of the library with extra -D macros: python tools/ab_build.py name MACRO=VALUE ...  -> ab/lib_<name>.so
(run a variant with LMDEC_LIB=ab/lib_<name>.so; ab/ is git-ignored but travels to the GPU box)."""
import os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lmdec_b200 import _build

name, macros = sys.argv[1], sys.argv[2:]
out_dir = os.path.join(os.path.dirname(_build.HERE), "ab")
os.makedirs(out_dir, exist_ok=True)
out = os.path.join(out_dir, f"lib_{name}.so")
cmd = [_build._nvcc()] + _build.NVCC_FLAGS + [f"-D{m}" for m in macros] + ["-o", out] + _build.SOURCES
res = subprocess.run(cmd, capture_output=True, text=True)
if res.returncode:
    raise SystemExit(res.stdout + res.stderr)
print(out)

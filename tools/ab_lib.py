"""Run bench.py against an alternative build of the library (A/B of compile-time variants):
    python tools/ab_lib.py mtgs_b200/libb200splat_cg3.so --steps 30 --warmup 5 --no-cpu-baseline"""
import os
import runpy
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mtgs_b200._lib as L  # noqa: E402

L.LIB_PATH = os.path.abspath(sys.argv[1])
sys.argv = [os.path.join(ROOT, "bench.py")] + sys.argv[2:]
runpy.run_path(sys.argv[0], run_name="__main__")

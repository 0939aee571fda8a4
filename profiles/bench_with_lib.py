"""run bench.py against another build of the CUDA library (same-box A/B of a kernel change):
    python profiles/bench_with_lib.py profiles/ab/lib_gemm_before.so --steps 10 --warmup 3 --no-cpu-baseline ..."""
import os, runpy, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dualdiff_b200._lib as L
L.LIB_PATH = os.path.abspath(sys.argv[1])
sys.argv = [os.path.join(ROOT, "bench.py")] + sys.argv[2:]
runpy.run_path(sys.argv[0], run_name="__main__")

"""One process, one torch import: the tf32 GEMM microbenchmark (tensor-core tier only), then its -m gpu unit tests."""
import contextlib
import io
import os
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
os.chdir(ROOT)
os.makedirs("gpurun_out", exist_ok=True)
sys.argv = ["bench_tf32_gemm.py", "8", "tf32"]
import tools.bench_tf32_gemm as MB  # noqa: E402

buf = io.StringIO()
with contextlib.redirect_stdout(buf):
    MB.main()
open("gpurun_out/i_micro.log", "w").write(buf.getvalue())
import pytest  # noqa: E402

rc = pytest.main(["tests/test_gpu_tf32_gemm.py", "-q", "-x", "-p", "no:cacheprovider"])
open("gpurun_out/i_rc.txt", "w").write(f"pytest rc {int(rc)}\n")
print("pytest rc", int(rc))

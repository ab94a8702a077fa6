"""Build timing-experiment variants of the library (-DCEVB_STUB=<mask>, see csrc/fused.cu) into
csrc/libcevb200_stub<mask>.so; select one with CEVB_LIB=<path>.
usage: python tools/stub_build.py 1 2 4 ..."""
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from chromevol_enhanced_b200 import _build  # noqa: E402

for mask in sys.argv[1:]:
    out = os.path.join(_build.CSRC, f"libcevb200_stub{mask}.so")
    srcs = [os.path.join(_build.CSRC, s) for s in _build.SOURCES]
    cmd = [_build._nvcc()] + _build.NVCC_FLAGS + [f"-DCEVB_STUB={mask}", "-o", out] + srcs
    subprocess.run(cmd, check=True)
    print(out)

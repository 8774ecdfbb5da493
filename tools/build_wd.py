"""Build csrc/libllamax_b200_wd.so: the same library with -DLX_WATCHDOG (mbarrier waits trap after 2^26 failed polls), used
for the FIRST run of a new kernel on the GPU box: a protocol bug then surfaces as a launch failure instead of a hung GPU.
    python tools/build_wd.py ;  LLAMAX_B200_LIB=llamax_b200/csrc/libllamax_b200_wd.so python -m pytest tests -m gpu ..."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from llamax_b200.build import CSRC, FLAGS, NVCC, SOURCES

out = os.path.join(CSRC, "libllamax_b200_wd.so")


def cc(src):
    obj = os.path.join(CSRC, src.replace(".cu", "_wd.o"))
    subprocess.check_call([NVCC, *FLAGS, "-DLX_WATCHDOG", "-c", os.path.join(CSRC, src), "-o", obj])
    return obj


with ThreadPoolExecutor(4) as ex:
    objs = list(ex.map(cc, SOURCES))
subprocess.check_call([NVCC, "-shared", "-o", out, *objs, "-gencode", "arch=compute_100a,code=sm_100a"])
print("built", out)

"""Build the C-ABI CUDA library in-tree: llamax_b200/csrc/libllamax_b200.so (sm_100a only).

    python -m llamax_b200.build [--force]

nvcc cross-compiles without a GPU. Objects are rebuilt only when a source or header is newer.
"""

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")
LIB = os.path.join(CSRC, "libllamax_b200.so")
SOURCES = ["host_utils.cu", "elementwise.cu", "gemm.cu", "attention.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-std=c++17", "-O3", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "-I", CSRC, "-I", INCLUDE,
]


def _newest_header_mtime():
    paths = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    paths += [os.path.join(INCLUDE, f) for f in os.listdir(INCLUDE) if f.endswith(".h")]
    return max(os.path.getmtime(p) for p in paths)


def _compile(src, force):
    obj = os.path.join(CSRC, src.replace(".cu", ".o"))
    srcp = os.path.join(CSRC, src)
    if not force and os.path.exists(obj):
        if os.path.getmtime(obj) >= max(os.path.getmtime(srcp), _newest_header_mtime()):
            return obj, False
    cmd = [NVCC, *FLAGS, *os.environ.get("LLAMAX_NVCC_FLAGS", "").split(), "-c", srcp, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    return obj, True


def build(force: bool = False, verbose: bool = True) -> str:
    srcs = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    with ThreadPoolExecutor(max_workers=len(srcs)) as ex:
        results = list(ex.map(lambda s: _compile(s, force), srcs))
    objs = [o for o, _ in results]
    rebuilt = any(ch for _, ch in results)
    if rebuilt or not os.path.exists(LIB):
        cmd = [NVCC, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
        if verbose:
            print(f"[llamax_b200] built {LIB}")
    elif verbose:
        print(f"[llamax_b200] up to date: {LIB}")
    return LIB


REFERENCE_SRC = os.environ.get("LLAMAX_REFERENCE", "/root/reference")
REFERENCE_DST = os.path.join(os.path.dirname(HERE), "baseline", "_ref")


def install_reference(verbose: bool = True) -> str | None:
    """Install the UNMODIFIED reference into the git-ignored `baseline/_ref` (it travels to the GPU box with the
    snapshot) so that `bench.py --impl reference` and tools/incumbents.py can run the reference's own code there.
    `pip install --target baseline/_ref /root/reference` fails (the reference has no package metadata: a flat layout
    with two top-level packages and a pyproject that only configures formatters), so the install is what pip would
    have done for a pure-Python project: its two packages, byte for byte. No-op where /root/reference is absent."""
    import filecmp
    import shutil

    if not os.path.isdir(REFERENCE_SRC):
        return REFERENCE_DST if os.path.isdir(os.path.join(REFERENCE_DST, "modelling")) else None
    for pkg in ("modelling", "subclasses"):
        src, dst = os.path.join(REFERENCE_SRC, pkg), os.path.join(REFERENCE_DST, pkg)
        same = os.path.isdir(dst) and not filecmp.dircmp(src, dst, ignore=["__pycache__"]).diff_files \
            and not filecmp.dircmp(src, dst, ignore=["__pycache__"]).left_only
        if not same:
            shutil.rmtree(dst, ignore_errors=True)
            shutil.copytree(src, dst, ignore=shutil.ignore_patterns("__pycache__"))
            if verbose:
                print(f"[llamax_b200] installed reference package {pkg} -> {dst}")
    return REFERENCE_DST


if __name__ == "__main__":
    build(force="--force" in sys.argv)
    install_reference()

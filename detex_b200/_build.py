"""Build the in-tree CUDA shared library (sm_100a) with nvcc.

    python -m detex_b200._build [--force]

The .so lands in detex_b200/_C/ (git-ignored, but shipped to the GPU box by gpurun).
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "_C")
LIB = os.path.join(OUT_DIR, "libdetex_b200.so")
SOURCES = ["dtx_api.cu", "k0_prep.cu", "k1_project.cu", "k_direct.cu", "k3_post.cu", "k4_ccx.cu", "k6_stalta.cu", "k7_mag.cu", "k8_preproc.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared", "-Xlinker", "--no-undefined", "--use_fast_math=false",
]


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    for f in os.listdir(CSRC):
        if os.path.getmtime(os.path.join(CSRC, f)) > t:
            return True
    inc = os.path.join(HERE, "..", "include", "detex_b200.h")
    return os.path.getmtime(inc) > t


def build(force=False, verbose=False, extra_flags=(), out=None):
    """Compile the library.  `extra_flags` / `out` build experiment variants next to it."""
    if out is None and not force and not _stale():
        return LIB
    os.makedirs(OUT_DIR, exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    flags = [f for f in NVCC_FLAGS if not f.startswith("--use_fast_math")]
    if verbose:
        flags += ["-Xptxas", "-v"]
    target = out or LIB
    cmd = [nvcc] + flags + list(extra_flags) + [os.path.join(CSRC, s) for s in SOURCES] + ["-o", target]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose:
        sys.stderr.write(r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    return target


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

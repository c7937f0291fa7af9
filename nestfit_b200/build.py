"""Build libnestfit_b200.so (hand-written sm_100a CUDA + C ABI) in-tree with nvcc."""
import os
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB = PKG / "libnestfit_b200.so"
SOURCES = ["nf_model.cu", "nf_nh3.cu", "nf_priors.cu", "nf_sampler.cu", "nf_peaks.cu", "nf_capi.cu"]
NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-shared", "-Xcompiler", "-fPIC", "-DNF_BUILD",
]


def _stale():
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    deps = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list((PKG.parent / "include").glob("*.h"))
    return any(p.stat().st_mtime > t for p in deps)


def build(force=False, verbose=False):
    """Compile every CUDA source for sm_100a; returns the library path."""
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    srcs = [str(CSRC / s) for s in SOURCES if (CSRC / s).exists()]
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", str(LIB)] + srcs
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building libnestfit_b200.so")
    if verbose:
        sys.stderr.write(res.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

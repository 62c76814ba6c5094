"""Builds csrc/*.cu into lib/libqcgraft.so for sm_100a with nvcc (cross-compiles without a GPU)."""
from __future__ import annotations

import os
import subprocess
import shutil

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_DIR = os.path.join(PKG_DIR, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libqcgraft.so")

SOURCES = ["api.cu", "k1_preprocess.cu", "k2_yolo.cu", "k3_roi.cu", "k3b_resize.cu", "k4_segment.cu", "k5_spc.cu", "k6_rules.cu", "k10_jpeg.cu"]
# -fmad=false: the parity-critical float paths must not be contracted into FMAs (the reference's
# CPU libraries are not); these kernels are memory-bound so nothing is lost.
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-fmad=false", "-Xcompiler", "-fPIC"] + os.environ.get("QC_NVCC_EXTRA", "").split()


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _torch_cudart_dir():
    try:
        import importlib.util
        spec = importlib.util.find_spec("nvidia.cuda_runtime")
        if spec and spec.submodule_search_locations:
            d = os.path.join(list(spec.submodule_search_locations)[0], "lib")
            if os.path.exists(os.path.join(d, "libcudart.so.12")):
                return d
    except Exception:
        pass
    return None


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps.append(os.path.join(os.path.dirname(PKG_DIR), "include", "qcgraft.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(LIB_DIR, src.replace(".cu", ".o"))
        cmd = [_nvcc(), *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{out}")
        if verbose and out:
            print(out)
    # One CUDA runtime per process: link the SAME libcudart.so.12 torch loads, so the library
    # shares torch's current-device state, primary context and stream handles.
    cmd = [_nvcc(), "-shared", "-o", LIB_PATH, *objs, "-gencode", "arch=compute_100a,code=sm_100a",
           "-Xcompiler", "-fPIC", "-cudart", "shared"]
    rt = _torch_cudart_dir()
    if rt:
        cmd += ["-L", rt, "-Xlinker", f"-rpath={rt}"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}")
    return LIB_PATH


if __name__ == "__main__":
    print(build(force=True, verbose=True))

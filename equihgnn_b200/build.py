"""In-tree build of libeqh_b200.so (sm_100a only) and of the oracle's native helpers.

    python -m equihgnn_b200.build [--force]

nvcc cross-compiles without a GPU; the built library travels to the GPU box with the
repository snapshot (it is git-ignored, not gpurun-ignored).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(ROOT, "include")
LIB = os.path.join(CSRC, "libeqh_b200.so")
OBJ = os.path.join(CSRC, "build")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr",
    "-I", INCLUDE, "-I", CSRC,
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libeqh_b200.so cannot be built")


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every .cu under csrc/ for sm_100a and link libeqh_b200.so.  Returns its path."""
    nvcc = _nvcc()
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(INCLUDE, "eqh_b200.h"))
    srcs = sources()
    objs = [os.path.join(OBJ, os.path.basename(s)[:-3] + ".o") for s in srcs]

    def compile_one(pair):
        src, obj = pair
        if not force and not _stale(obj, [src] + headers):
            return None
        extra = os.environ.get("EQH_NVCC_EXTRA", "").split()          # e.g. -DEQH_TRACE for the phase-timing experiments
        cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        return r.stderr

    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        logs = list(ex.map(compile_one, zip(srcs, objs)))
    if verbose:
        for lg in logs:
            if lg:
                sys.stderr.write(lg)
    if force or _stale(LIB, objs):
        cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs + ["-lcudart"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


TORCH_EXT_SRC = os.path.join(CSRC, "torch_ext", "eqh_torch.cpp")
TORCH_EXT_LIB = os.path.join(CSRC, "libeqh_b200_torch.so")


def build_torch_ext(force: bool = False) -> str:
    """Compile the thin torch C++ extension over the C ABI (csrc/torch_ext/eqh_torch.cpp -> libeqh_b200_torch.so,
    `torch.ops.load_library`).  Plain g++ against torch's headers: the file holds no kernels."""
    import torch
    from torch.utils import cpp_extension as ce
    build()                                   # the C-ABI library it links against
    if not force and not _stale(TORCH_EXT_LIB, [TORCH_EXT_SRC, LIB, os.path.join(INCLUDE, "eqh_b200.h")]):
        return TORCH_EXT_LIB
    tlib = os.path.join(os.path.dirname(torch.__file__), "lib")
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    cmd = ["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-fvisibility=hidden",
           f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}", "-DTORCH_API_INCLUDE_EXTENSION_H"]
    for inc in ce.include_paths() + [cuda_inc, INCLUDE]:
        cmd += ["-I", inc]
    import sysconfig
    cmd += ["-I", sysconfig.get_paths()["include"]]
    cmd += [TORCH_EXT_SRC, "-o", TORCH_EXT_LIB, "-L", tlib, "-L", CSRC, "-ltorch", "-ltorch_cpu", "-lc10", "-ltorch_cuda",
            "-lc10_cuda", "-leqh_b200", f"-Wl,-rpath,{tlib}", "-Wl,-rpath,$ORIGIN"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"torch extension build failed:\n{r.stdout}\n{r.stderr[-4000:]}")
    return TORCH_EXT_LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(path)
    if "--torch-ext" in sys.argv:
        print(build_torch_ext(force="--force" in sys.argv))

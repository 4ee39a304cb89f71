"""Builds libosd_b200.so (the C-ABI library with the sm_100a kernels) in-tree with nvcc.

    python -m oneshotdet_b200.build [-f] [-v]

nvcc cross-compiles for sm_100a without a GPU.  The .so lands in oneshotdet_b200/lib/ (git-ignored,
shipped to the GPU box by gpurun).  There is no other backend and no CPU fallback: if this library is
missing the package raises on first use.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
OBJ = os.path.join(PKG, "lib", "obj")
LIB = os.path.join(PKG, "lib", "libosd_b200.so")
SOURCES = ["common.cu", "nms.cu", "fcos_post.cu", "match.cu", "fusion_conv.cu", "fusion_fused.cu", "support_pool.cu", "box_post.cu", "roi_pool.cu", "coco_writer.cu", "peer_comm.cu", "box_head.cu"]
HEADERS = ["osd_common.cuh", "osd_device_utils.cuh", os.path.join(ROOT, "include", "osd_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "-I", os.path.join(ROOT, "include"),
    "-I", CSRC,
]


def nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; libosd_b200.so cannot be built")
    return exe


def _stale(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    sources = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    hdrs = [h if os.path.isabs(h) else os.path.join(CSRC, h) for h in HEADERS]
    hdrs += [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.abspath(__file__))
    jobs = []
    objs = []
    for s in sources:
        src = os.path.join(CSRC, s)
        obj = os.path.join(OBJ, s.replace(".cu", ".o"))
        objs.append(obj)
        if force or _stale(obj, [src] + hdrs):
            jobs.append([nvcc()] + NVCC_FLAGS + ["-c", src, "-o", obj])

    def run(cmd):
        if verbose:
            print(" ".join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
        return r

    with ThreadPoolExecutor(max_workers=min(4, max(1, len(jobs)))) as ex:
        list(ex.map(run, jobs))
    if jobs or force or _stale(LIB, objs):
        run([nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs +
            ["-cudart", "static"])
    return LIB


EXT = os.path.join(PKG, "_C_torch.so")


def build_torch_extension(force: bool = False, verbose: bool = False) -> str:
    """Thin pybind11/torch extension `oneshotdet_b200._C_torch` over the C ABI (csrc/torch_ext.cpp): g++ directly, linked
    against libosd_b200.so (rpath $ORIGIN/lib) and torch's libraries."""
    import sysconfig

    import torch
    from torch.utils import cpp_extension as ce

    src = os.path.join(CSRC, "torch_ext.cpp")
    lib = build(force=False, verbose=verbose)
    if not force and not _stale(EXT, [src, lib, os.path.join(ROOT, "include", "osd_b200.h")]):
        return EXT
    inc = []
    for p in ce.include_paths() + [sysconfig.get_paths()["include"], "/usr/local/cuda/include"]:
        inc += ["-isystem", p]
    tlib = os.path.join(os.path.dirname(torch.__file__), "lib")
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-w", "-DTORCH_EXTENSION_NAME=_C_torch",
           "-DTORCH_API_INCLUDE_EXTENSION_H", "-D_GLIBCXX_USE_CXX11_ABI=%d" % int(torch._C._GLIBCXX_USE_CXX11_ABI),
           "-I", os.path.join(ROOT, "include")] + inc + [src, "-o", EXT,
           "-L", os.path.join(PKG, "lib"), "-losd_b200", "-Wl,-rpath,$ORIGIN/lib",
           "-L", tlib, "-Wl,-rpath," + tlib, "-lc10", "-lc10_cuda", "-ltorch", "-ltorch_cpu", "-ltorch_cuda", "-ltorch_python"]
    if verbose:
        print(" ".join(cmd), flush=True)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("g++ failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    return EXT


if __name__ == "__main__":
    print(build(force="-f" in sys.argv, verbose="-v" in sys.argv))
    if "--ext" in sys.argv:
        print(build_torch_extension(force="-f" in sys.argv, verbose="-v" in sys.argv))

"""TEST INFRASTRUCTURE ONLY -- builds the checkers, is never part of the product path.

Two recipes:

* ``build_c_oracle()``  -> ``oracle/_build/libosd_oracle.so``: the plain-C restatement
  (``oracle/nms_oracle.c``), gcc only, no torch.
* ``build_ref()``       -> ``oracle/_ref/osd_ref_C.so``: the reference's *own* CPU operators
  (``maskrcnn_benchmark/csrc/cpu/nms_cpu.cpp``, ``cpu/ROIAlign_cpu.cpp`` and the ``nms.h`` /
  ``ROIAlign.h`` dispatchers) compiled from the sources where they lie under ``/root/reference``
  through ``oracle/ref_wrap/ref_ops_wrap.cpp``.  g++ is invoked directly on those files; the
  reference's ``setup.py`` is not run.  Needs ``/root/reference``, so it only happens in the
  build container; the GPU box uses the prebuilt ``.so`` that gpurun ships.

Flags follow what the reference's ``setup.py:17-56`` would give a ``CppExtension`` on this
interpreter (sysconfig CFLAGS = ``-O2``), no ``-march``/``-mfma`` (so no FMA contraction in
``nms_cpu_kernel``), no ``WITH_CUDA``.
"""
from __future__ import annotations

import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("OSD_REFERENCE_ROOT", "/root/reference")
REF_CSRC = os.path.join(REF_ROOT, "maskrcnn_benchmark", "csrc")
REF_OUT_DIR = os.path.join(HERE, "_ref")
C_OUT_DIR = os.path.join(HERE, "_build")
REF_SO = os.path.join(REF_OUT_DIR, "osd_ref_C.so")
C_SO = os.path.join(C_OUT_DIR, "libosd_oracle.so")


def _newer(target: str, sources: list[str]) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources if os.path.exists(s))


def build_c_oracle(force: bool = False, verbose: bool = False) -> str:
    src = os.path.join(HERE, "nms_oracle.c")
    if not force and _newer(C_SO, [src]):
        return C_SO
    os.makedirs(C_OUT_DIR, exist_ok=True)
    # -ffp-contract=off: the reference translation unit is built for baseline x86-64 where
    # g++ cannot contract a*b+c; make that explicit so the restatement rounds identically
    # on any host.
    cmd = ["gcc", "-O2", "-std=c11", "-fPIC", "-shared", "-ffp-contract=off",
           "-fno-fast-math", src, "-o", C_SO, "-lm"]
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return C_SO


def ref_available() -> bool:
    return os.path.exists(os.path.join(REF_CSRC, "cpu", "nms_cpu.cpp"))


def build_ref(force: bool = False, verbose: bool = False) -> str | None:
    """Compile the reference CPU ops. Returns the .so path, or None if neither the
    reference tree nor a prebuilt .so is present."""
    wrap = os.path.join(HERE, "ref_wrap", "ref_ops_wrap.cpp")
    if not ref_available():
        return REF_SO if os.path.exists(REF_SO) else None
    srcs = [wrap,
            os.path.join(REF_CSRC, "cpu", "nms_cpu.cpp"),
            os.path.join(REF_CSRC, "cpu", "ROIAlign_cpu.cpp"),
            os.path.join(REF_CSRC, "nms.h"),
            os.path.join(REF_CSRC, "ROIAlign.h")]
    if not force and _newer(REF_SO, srcs):
        return REF_SO
    import torch  # noqa: F401  (only for include / lib paths)
    from torch.utils import cpp_extension as ce

    os.makedirs(REF_OUT_DIR, exist_ok=True)
    inc = []
    for p in ce.include_paths():
        inc += ["-isystem", p]
    inc += ["-isystem", sysconfig.get_paths()["include"]]
    libdir = os.path.join(os.path.dirname(torch.__file__), "lib")

    def q(path):
        return '"%s"' % path

    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-w",
           "-DTORCH_EXTENSION_NAME=osd_ref_C", "-DTORCH_API_INCLUDE_EXTENSION_H",
           "-D_GLIBCXX_USE_CXX11_ABI=%d" % int(torch._C._GLIBCXX_USE_CXX11_ABI),
           "-DOSD_REF_NMS_CPU_CPP=" + q(srcs[1]),
           "-DOSD_REF_ROIALIGN_CPU_CPP=" + q(srcs[2]),
           "-DOSD_REF_NMS_H=" + q(srcs[3]),
           "-DOSD_REF_ROIALIGN_H=" + q(srcs[4]),
           "-I", REF_CSRC] + inc + [wrap, "-o", REF_SO,
           "-L", libdir, "-Wl,-rpath," + libdir,
           "-lc10", "-ltorch", "-ltorch_cpu", "-ltorch_python"]
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return REF_SO


PYREF_DIR = os.path.join(REF_OUT_DIR, "pyref")


def vendor_reference_python(force: bool = False) -> str | None:
    """Copy the reference's Python package (maskrcnn_benchmark/**/*.py, unmodified, no csrc) next to the compiled ops,
    into oracle/_ref/pyref -- git-ignored build output that gpurun ships to the GPU box, so that `bench.py --impl
    reference` can run the reference's own FCOSPostProcessor there.  Needs /root/reference; returns the directory, or
    None when neither the reference tree nor a previous copy exists."""
    import shutil

    src_pkg = os.path.join(REF_ROOT, "maskrcnn_benchmark")
    dst_pkg = os.path.join(PYREF_DIR, "maskrcnn_benchmark")
    marker = os.path.join(dst_pkg, "modeling", "rpn", "fcos", "inference.py")
    if not os.path.isdir(src_pkg):
        return PYREF_DIR if os.path.exists(marker) else None
    if os.path.exists(marker) and not force:
        return PYREF_DIR
    if os.path.isdir(dst_pkg):
        shutil.rmtree(dst_pkg)
    for root, dirs, files in os.walk(src_pkg):
        rel = os.path.relpath(root, src_pkg)
        if rel.split(os.sep)[0] == "csrc":
            dirs[:] = []
            continue
        os.makedirs(os.path.join(dst_pkg, rel), exist_ok=True)
        for f in files:
            if f.endswith(".py"):
                shutil.copyfile(os.path.join(root, f), os.path.join(dst_pkg, rel, f))
    return PYREF_DIR


def load_reference_python():
    """Make the vendored, unmodified reference package importable with the compiled CPU ops injected as
    ``maskrcnn_benchmark._C``.  Returns the ``_C`` module, or None if unavailable."""
    ref_c = load_ref()
    pyref = vendor_reference_python()
    if ref_c is None or pyref is None:
        return None
    if pyref not in sys.path:
        sys.path.insert(0, pyref)
    import maskrcnn_benchmark  # noqa: PLC0415

    maskrcnn_benchmark._C = ref_c
    sys.modules["maskrcnn_benchmark._C"] = ref_c
    return ref_c


def load_ref():
    """Import oracle/_ref/osd_ref_C.so as a module exposing nms / roi_align_forward.
    Returns None when it has not been built (and cannot be)."""
    so = build_ref()
    if so is None or not os.path.exists(so):
        return None
    import importlib.util
    import torch  # noqa: F401  must be imported before the extension

    spec = importlib.util.spec_from_file_location("osd_ref_C", so)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    v = "-v" in sys.argv
    print("C oracle  :", build_c_oracle(force="-f" in sys.argv, verbose=v))
    print("reference :", build_ref(force="-f" in sys.argv, verbose=v))
    print("ref python:", vendor_reference_python(force="-f" in sys.argv))

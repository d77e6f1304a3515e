"""Builds pl_yolo_b200/libplyolo.so (hand-written sm_100a kernels + the C ABI of include/plyolo.h).

In-tree on purpose: the .so is git-ignored but travels with the repo snapshot to the GPU box.
Flags: sm_100a only, -fmad=false (one rounding per reference op; the single deliberate fma is
explicit), no fast-math, -lineinfo for ncu source pages, static cudart (no runtime dependency on
torch's CUDA libraries: the library is plain C ABI).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libplyolo.so")
SOURCES = ["api.cu", "decode.cu", "postprocess.cu", "simota.cu", "simota_ops.cu", "loss.cu", "evaluator.cu"]
# postprocess.cu launches nms_general_kernel from the device (CUDA dynamic parallelism): relocatable device code
RDC = {"postprocess.cu": ["-rdc=true"]}
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-fmad=false", "-std=c++17",
    "-Xcompiler", "-fPIC", "--cudart", "static",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "plyolo.h"), __file__]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, defines=(), out: str = OUT) -> str:
    """`defines` / `out`: kernel variants for tuning (e.g. defines=["PLYOLO_SCORE_REGS=56"], out=".../variant.so")."""
    if not force and out == OUT and not needs_build():
        return OUT
    nvcc = _nvcc()
    objdir = os.path.join(HERE, "build" if out == OUT else "build_" + os.path.basename(out).replace(".", "_"))
    os.makedirs(objdir, exist_ok=True)
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, *RDC.get(src, []), *["-D" + d for d in defines], "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for src, pr in procs:
        log, _ = pr.communicate()
        if verbose or pr.returncode != 0:
            sys.stderr.write(log)
        if pr.returncode != 0:
            raise RuntimeError("nvcc failed on %s" % src)
    link = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "--cudart", "static", "-o", out, *objs, "-lcudadevrt"]
    subprocess.run(link, check=True)
    return out


if __name__ == "__main__":
    defs = [a[2:] for a in sys.argv[1:] if a.startswith("-D")]
    outs = [a[6:] for a in sys.argv[1:] if a.startswith("--out=")]
    print(build(force="--force" in sys.argv or bool(outs), verbose="-v" in sys.argv, defines=defs, out=os.path.abspath(outs[0]) if outs else OUT))

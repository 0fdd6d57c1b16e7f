"""Builds gta_b200/libgta_b200.so in-tree with nvcc for sm_100a (no torch, no JIT cache).

    python -m gta_b200.build [--force] [--verbose]
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libgta_b200.so")
SOURCES = ["gta_abi.cu", "gta_reps.cu", "gta_rotate_kv.cu", "gta_generic.cu", "gta_attn_fwd2.cu", "gta_attn_fwd3.cu", "gta_attn_fwd4.cu", "gta_attn_fwd4_ct.cu", "gta_attn_fwd5a.cu", "gta_attn_fwd5b.cu", "gta_attn_fwd6.cu", "gta_attn_fwd_hp.cu", "gta_attn_bwd.cu", "gta_attn_bwd2.cu"]
# development library (probes, micro-benchmarks, the first-generation kernel): include/gta_b200_dev.h
DEV_SOURCES = ["gta_dev_abi.cu", "gta_attn_fwd.cu"]
DEV_OUT = os.path.join(HERE, "libgta_b200_dev.so")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _stale() -> bool:
    if not os.path.exists(OUT) or not os.path.exists(DEV_OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "gta_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, defines=(), out: str = OUT) -> str:
    if not force and not defines and out == OUT and not _stale():
        return OUT
    objdir = os.path.join(HERE, "build" + ("_" + "_".join(d.replace("=", "") for d in defines) if defines else ""))
    os.makedirs(objdir, exist_ok=True)
    nvcc = _nvcc()
    common = [nvcc, *ARCH, "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
              "-I", os.path.join(HERE, "..", "include")]
    common += ["-D" + d for d in defines]
    if verbose:
        common += ["-Xptxas", "-v"]
    objs = []
    procs = []
    dev_objs = [os.path.join(objdir, src.replace(".cu", ".o")) for src in DEV_SOURCES]
    for src in SOURCES + DEV_SOURCES:
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        if src in SOURCES:
            objs.append(obj)
        procs.append((src, subprocess.Popen(common + ["-c", os.path.join(CSRC, src), "-o", obj],
                                            stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for src, p in procs:
        log, _ = p.communicate()
        if verbose or p.returncode:
            sys.stderr.write(log)
        if p.returncode:
            raise RuntimeError("nvcc failed on %s" % src)
    subprocess.check_call([nvcc, *ARCH, "-shared", "-o", out, *objs, "-lcudart"])
    if out == OUT:
        subprocess.check_call([nvcc, *ARCH, "-shared", "-o", DEV_OUT, *dev_objs, "-lcudart"])
    return out


if __name__ == "__main__":
    # tuning variants: python -m gta_b200.build -DGTA_POLY_NUM=1 --out gta_b200/libgta_b200_p14.so
    defs = [a[2:] for a in sys.argv if a.startswith("-D")]
    out = sys.argv[sys.argv.index("--out") + 1] if "--out" in sys.argv else OUT
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, defines=defs, out=os.path.abspath(out)))

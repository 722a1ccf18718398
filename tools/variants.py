"""Builds tuning variants of libmassb200.so (same sources, different -D flags for cells.cu) into
mass_b200/csrc/variants/<name>.so; MASSB200_LIB=<path> makes mass_b200 load one of them.
    python tools/variants.py name1="-DMB_TG_MINB=5 -DMB_WHASH=256" name2="..."
On the GPU box:  for v in mass_b200/csrc/variants/*.so; do MASSB200_LIB=$v python bench.py ...; done"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "mass_b200", "csrc")
OUT = os.path.join(CSRC, "variants")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-fmad=false", "-Xcompiler",
         "-fPIC,-fvisibility=hidden", "-Wno-deprecated-gpu-targets"]


def main():
    os.makedirs(OUT, exist_ok=True)
    subprocess.check_call(["make", "-C", CSRC, "-j8"], stdout=subprocess.DEVNULL)
    others = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".o") and f != "cells.o"]
    procs = []
    for spec in sys.argv[1:]:
        name, extra = spec.split("=", 1)
        obj = os.path.join(OUT, name + ".o")
        procs.append((name, obj, subprocess.Popen(["nvcc"] + FLAGS + extra.split() + ["-c", os.path.join(CSRC, "cells.cu"), "-o", obj],
                                                  stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True)))
    for name, obj, p in procs:
        err = p.communicate()[1]
        if p.returncode:
            print(name, "FAILED\n", err[-2000:])
            continue
        so = os.path.join(OUT, name + ".so")
        subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", so, obj] + others)
        os.remove(obj)
        print("built", so)


if __name__ == "__main__":
    main()

"""Builds tests/_build/libbj_emul.so: bj_factor.cu and bj_solve.cu compiled for the CPU emulation of tests/emul/cuda_emul.h.
The only source transformation is the launch syntax of bj_factor.cu: `k<<<grid, block, smem, stream>>>(args);` becomes
`emul_launch(grid, block, smem, [&] { k(args); });` (bj_solve.cu launches through cudaLaunchKernelEx, which the shim provides)."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
BUILD = os.path.join(ROOT, "tests", "_build")
CSRC = os.path.join(ROOT, "prealps_b200", "csrc")
EMUL = os.path.join(ROOT, "tests", "emul")
METIS_A = "/usr/local/cuda/targets/x86_64-linux/lib/libmetis_static.a"


def split_top(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "(<[":
            depth += 1
        elif ch in ")>]":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip()); cur = ""
        else:
            cur += ch
    out.append(cur.strip())
    return out


def rewrite_launches(text):
    pat = re.compile(r"(\b\w+)<<<(.*?)>>>\((.*?)\);", re.S)

    def sub(m):
        cfg = split_top(m.group(2))
        assert len(cfg) == 4, cfg
        return "emul_launch(%s, %s, %s, [&] { %s(%s); });" % (cfg[0], cfg[1], cfg[2], m.group(1), m.group(3))
    out, n = pat.subn(sub, text)
    return out, n


def build(asan=False):
    """asan=True: the same library under AddressSanitizer (load it with LD_PRELOAD=libasan): out-of-bounds accesses of the
    kernels to "device" memory, which is host memory with red zones here"""
    os.makedirs(BUILD, exist_ok=True)
    so = os.path.join(BUILD, "libbj_emul_asan.so" if asan else "libbj_emul.so")
    deps = [os.path.join(CSRC, f) for f in ("bj_factor.cu", "bj_solve.cu", "bj_symbolic.cpp", "bj.h", "bj_symbolic.h", "common.cuh")]
    deps += [os.path.join(EMUL, f) for f in ("cuda_emul.h", "cuda_emul.cpp", "bj_emul_glue.cpp", "build_bj_emul.py")]
    if os.path.exists(so) and all(os.path.getmtime(d) <= os.path.getmtime(so) for d in deps):
        return so
    text, n = rewrite_launches(open(os.path.join(CSRC, "bj_factor.cu")).read())
    assert n >= 10 and "<<<" not in text, n
    gen = os.path.join(BUILD, "bj_factor_emul.cpp")
    open(gen, "w").write('#define PCU_EMUL 1\n#line 1 "bj_factor.cu"\n' + text)
    gen2 = os.path.join(BUILD, "bj_solve_emul.cpp")
    open(gen2, "w").write('#define PCU_EMUL 1\n#include "%s"\n' % os.path.join(CSRC, "bj_solve.cu"))
    extra = ["-g", "-fsanitize=address", "-fno-omit-frame-pointer"] if asan else []
    cmd = ["g++", "-O1"] + extra + ["-std=c++17", "-fPIC", "-shared", "-pthread", "-Wno-unknown-pragmas", "-I" + EMUL, "-I" + CSRC,
           gen, gen2, os.path.join(CSRC, "bj_symbolic.cpp"), os.path.join(EMUL, "cuda_emul.cpp"),
           os.path.join(EMUL, "bj_emul_glue.cpp"), METIS_A, "-o", so, "-lm", "-Wl,-Bsymbolic", "-Wl,--exclude-libs=ALL"]  # own symbols first: the product library may be loaded RTLD_GLOBAL
    subprocess.check_call(cmd)
    return so


if __name__ == "__main__":
    print(build())

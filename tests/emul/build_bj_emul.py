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
    """every `name<targs><<<grid, block, smem, stream>>>(args)` becomes `emul_launch(grid, block, smem, [&] { name<targs>(args); })`
    (an expression: it also works inside macro bodies and macro arguments)"""
    out, pos, n = "", 0, 0
    while True:
        i = text.find("<<<", pos)
        if i < 0:
            return out + text[pos:], n
        # kernel name: identifier, optionally followed by balanced template arguments
        j = i
        if text[j - 1] == ">":
            depth = 0
            while True:
                j -= 1
                if text[j] == ">":
                    depth += 1
                elif text[j] == "<":
                    depth -= 1
                    if depth == 0:
                        break
        k = j
        while k > 0 and (text[k - 1].isalnum() or text[k - 1] in "_:"):
            k -= 1
        name = text[k:i]
        e = text.find(">>>", i)
        cfg = split_top(text[i + 3:e])
        assert len(cfg) == 4, (name, cfg)
        a0 = e + 3
        assert text[a0] == "(", text[a0:a0 + 20]
        depth, a1 = 0, a0
        while True:
            if text[a1] == "(":
                depth += 1
            elif text[a1] == ")":
                depth -= 1
                if depth == 0:
                    break
            a1 += 1
        args = text[a0 + 1:a1]
        out += text[pos:k] + "emul_launch(%s, %s, %s, [&] { %s(%s); })" % (cfg[0], cfg[1], cfg[2], name, args)
        pos = a1 + 1
        n += 1


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
    # the sanitizer build keeps one OS thread per CUDA thread (EMUL_PTHREADS); the default engine schedules fibers
    extra = ["-g", "-fsanitize=address", "-fno-omit-frame-pointer", "-DEMUL_PTHREADS"] if asan else []
    cmd = ["g++", "-O1"] + extra + ["-std=c++17", "-fPIC", "-shared", "-pthread", "-Wno-unknown-pragmas", "-I" + EMUL, "-I" + CSRC,
           gen, gen2, os.path.join(CSRC, "bj_symbolic.cpp"), os.path.join(EMUL, "cuda_emul.cpp"),
           os.path.join(EMUL, "bj_emul_glue.cpp"), METIS_A, "-o", so, "-lm", "-Wl,-Bsymbolic", "-Wl,--exclude-libs=ALL"]  # own symbols first: the product library may be loaded RTLD_GLOBAL
    subprocess.check_call(cmd)
    return so


def build_full():
    """the whole product stack for the emulation: every .cu of libprealps_cuda and the plain-C host layer, as
    tests/_build/emul_lib/{libprealps_cuda,libprealps_b200,libmpishim}.so (same names: the host library finds the emulated
    CUDA library next to itself).  Loaded only by tests (prealps_b200.capi honours PREALPS_B200_LIBDIR)."""
    out = os.path.join(BUILD, "emul_lib")
    os.makedirs(out, exist_ok=True)
    cu = ["ctx.cu", "spmm.cu", "ecg_kernels.cu", "bj_factor.cu", "bj_solve.cu"]
    host = [os.path.join(ROOT, "prealps_b200", "host", f) for f in ("pa_csr.c", "pa_operator.c", "pa_block_jacobi.c", "pa_ecg.c", "pa_driver.c")]
    deps = [os.path.join(CSRC, f) for f in cu + ["bj_symbolic.cpp", "bj.h", "bj_symbolic.h", "common.cuh", "spmm_kernels.cuh"]] + host
    deps += [os.path.join(EMUL, f) for f in ("cuda_emul.h", "cuda_emul.cpp", "build_bj_emul.py")]
    deps += [os.path.join(ROOT, "mpishim", "mpishim.c")]
    so_cuda, so_host, so_mpi = (os.path.join(out, n) for n in ("libprealps_cuda.so", "libprealps_b200.so", "libmpishim.so"))
    if all(os.path.exists(x) for x in (so_cuda, so_host, so_mpi)) and all(os.path.getmtime(d) <= os.path.getmtime(so_host) for d in deps):
        return out
    gens = []
    for f in cu:
        text, _ = rewrite_launches(open(os.path.join(CSRC, f)).read())
        assert "<<<" not in text
        g = os.path.join(out, f.replace(".cu", "_emul.cpp"))
        open(g, "w").write('#define PCU_EMUL 1\n#line 1 "%s"\n' % f + text)
        gens.append(g)
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-pthread", "-Wno-unknown-pragmas", "-I" + EMUL, "-I" + CSRC,
                           "-I" + os.path.join(ROOT, "include")] + gens +
                          [os.path.join(CSRC, "bj_symbolic.cpp"), os.path.join(EMUL, "cuda_emul.cpp"), METIS_A, "-o", so_cuda, "-lm", "-ldl",
                           "-Wl,-Bsymbolic", "-Wl,--exclude-libs=ALL"])
    subprocess.check_call(["gcc", "-O2", "-fPIC", "-std=gnu99", "-shared", "-Wno-unused-result", "-I" + os.path.join(ROOT, "mpishim"),
                           os.path.join(ROOT, "mpishim", "mpishim.c"), "-o", so_mpi, "-lpthread"])
    subprocess.check_call(["gcc", "-O2", "-fPIC", "-std=gnu99", "-shared", "-Wno-unused-result", "-I" + os.path.join(ROOT, "include"),
                           "-I" + os.path.join(ROOT, "mpishim")] + host +
                          [METIS_A, "-Wl,--exclude-libs=ALL", "-L" + out, "-lprealps_cuda", "-lmpishim", "-Wl,-rpath,$ORIGIN", "-o", so_host,
                           "-lm", "-lpthread"])
    return out


if __name__ == "__main__":
    import sys
    print(build_full() if "--full" in sys.argv else build(asan="--asan" in sys.argv))

#!/bin/bash
# tools/gpu_tests_on_emulation.sh -- the `-m gpu` parity tests, UNCHANGED, against the CPU emulation of the whole stack
# (tests/emul: every .cu of libprealps_cuda compiled against a miniature CUDA runtime, plus the C host layer).  For when no
# GPU minute is left: logic and data flow of kernels and host code, not PTX semantics, asynchrony or time.
#     bash tools/gpu_tests_on_emulation.sh [extra pytest arguments, e.g. -k spmm]
# The tests that are sized for a GPU are left out (long panels cut across CTAs: 17 minutes here; the 48^3 properties; the
# 32^3 - 128^3 runs against the pinned reference, the elasticity ADAPT_BS cases), and so are the tests that start the
# driver BINARIES: those link the product library, which refuses to run without a GPU (no CPU fallback).
# PREALPS_TEST_CANDIDATES=1 adds the opt-in kernels.
set -e
cd "$(dirname "$0")/.."
libdir=$(python tests/emul/build_bj_emul.py --full | tail -n 1)
PREALPS_B200_LIBDIR=$libdir python -m pytest tests/test_gpu_kernels.py tests/test_gpu_ecg.py -q -m gpu -p no:cacheprovider \
    -k "not long_panels and not 48cubed and not larger_sizes and not full_size_configs and not elasticity and not unchanged and not driver_error and not kernel_bench_harness" "$@"

#!/bin/bash
set -u
out=gpurun_out; mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "block_jacobi" > $out/r02_t1.log 2>&1; echo "bj tests rc=$?"; tail -n 2 $out/r02_t1.log
timeout 300 python tools/variants.py 64 1 8 2>&1 | grep " levels "
timeout 300 python tools/variants.py 128 8 8 2>&1 | grep " levels "
python - <<'PY' > /tmp/gen.log 2>&1
import sys; sys.path.insert(0, "oracle")
import gen_matrices
gen_matrices.write_mtx("/tmp/stencil27_48.mtx", gen_matrices.stencil27(48))
PY
MPISHIM_NP=1 timeout 600 prealps_b200/bin/bench_kernels -m /tmp/stencil27_48.mtx -k both > $out/r02_bench_kernels.log 2>&1; echo "bench_kernels rc=$?"; tail -n 25 $out/r02_bench_kernels.log

"""tools/spmm_sweep.py -- BASELINE config 3: SpMM microbench on the 27-point stencil (and the 7-point operator),
t = 1..32, 10 timed repetitions after 2 warm-ups, L2 flushed between repetitions (shape of the reference's
examples/test_bench_spmm.c:194-215).  Prints one JSON line per (operator, t)."""
import ctypes as C
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from prealps_b200 import capi  # noqa: E402
n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
ts = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1, 2, 4, 8, 16, 32]
peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"] \
    if os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) else 6650.0
for kind, name in ((1, "stencil27"), (0, "poisson7")):
    assert capi.lib.preAlps_b200_OperatorBuildStencil(kind, n, 8, 0, 8) == 0
    for t in ts:
        ms = C.c_float()
        capi.lib.preAlps_b200_BenchKernel(0, t, 10, 1, C.byref(ms))
        b = capi.stat("spmm_bytes_t%d" % t)
        gbs = b / ms.value / 1e6
        print(json.dumps({"operator": "%s %d^3" % (name, n), "bulk": os.environ.get("PREALPS_SPMM_BULK", "1"), "shape": os.environ.get("PREALPS_SPMM_SHAPE", "0"), "t": t, "us": round(ms.value * 1e3, 2), "algorithmic_MB": round(b / 1e6, 1),
                          "GB/s": round(gbs, 1), "frac_of_measured_peak": round(gbs / peak, 3), "frac_of_8TBs": round(gbs / 8000, 3)}), flush=True)

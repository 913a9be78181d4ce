#!/usr/bin/env python
"""bench.py -- ECG(t=8) + block-Jacobi on the synthetic 3-D Poisson operator of BASELINE.json.

    python bench.py [--gpus N --steps K --warmup W]          # this library, N processes x 1 GPU (torchrun for N > 1)
    python bench.py --impl reference [...]                    # the reference's CPU path on the host cores

A "step" is one ECG iteration (SpMM + block-Jacobi apply + the fused dense passes + stopping test) of the
configs[1] workload: 7-point Poisson 128^3 (2.1 M rows), t = 8, 8 METIS subdomains = 8 block-Jacobi blocks,
strong-scaled over N GPUs (each GPU owns 8/N consecutive subdomains).  `value` is iterations/s measured with
CUDA events over exactly K iterations, operands resident in HBM (working set ~17 GB >> the 126 MB L2).  `e2e`
is the same metric through the reference-facing RCI API with HOST buffers: rhs in, solution out, one 8-byte
residual read-back per iteration, i.e. iterations / time-to-solution of a whole solve.

`--impl reference` runs the UNMODIFIED reference (oracle/_ref/ecg_dump_ref: its sources + shims, 8 ranks x 1 thread) on
the SAME operator, partition, right-hand side and tolerance, for warmup + steps iterations (its setup -- MatrixMarket
load, METIS, the shim's scalar Cholesky: ~6 minutes at 128^3 -- is not part of iterations/s on either arm).  `--ref-n`
selects a smaller grid for a quick look; the line then names the grid that ran in `config`, not the headline one.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

METRIC = "ecg_t8_bjacobi_iterations_per_s"
UNIT = "iterations/s"
# --operator: (kind of preAlps_b200_OperatorBuildStencil, generator of oracle/gen_matrices.py for the CPU arm, description);
# the default (BASELINE configs[1]) is the 7-point Poisson operator, the others serve configs[2..4]
OPERATORS = {"poisson7": (0, lambda g, n: g.poisson7(n), "3D Poisson 7-point %d^3"),
             "stencil27": (1, lambda g, n: g.stencil27(n), "3D 27-point stencil %d^3"),
             "elasticity": (2, lambda g, n: g.elasticity3d(n, n, n), "3D Q1 linear elasticity %d^3 nodes x 3 dof")}


def n_rows(operator, n):
    return n ** 3 if operator != "elasticity" else 3 * n * n * (n - 1)  # the face x = 0 is clamped


def metric_name(args):
    return METRIC if args.t == 8 else "ecg_t%d_bjacobi_iterations_per_s" % args.t


def measured_peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)"""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        self.gpu, self.rows, self.proc = gpu, [], None

    def start(self):
        """nvidia-smi needs ~0.3 s before its first line: started well before the region, rows are kept by time stamp"""
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def wait_first_sample(self, timeout=3.0):
        t0 = time.time()
        while self.proc and not self.rows and time.time() - t0 < timeout:
            time.sleep(0.02)

    def __enter__(self):
        self.t_in = time.time()
        return self

    def __exit__(self, *a):
        self.t_out = time.time()

    def stop(self):
        if self.proc:
            time.sleep(0.05)
            self.proc.terminate()
            self.th.join(timeout=2)

    def summary(self):
        """samples taken inside the `with` region (the K timed iterations run back to back inside it; the region is
        padded by one sampling period on both sides so that a 20 ms region still sees the clocks under load)"""
        sm, mx, reasons = [], [], set()
        rows = [r for ts, r in self.rows if self.t_in - 0.03 <= ts <= self.t_out + 0.03]
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def reference_run(args, n_s, max_iter):
    """one run of the unmodified reference (oracle/_ref: its sources + shims) on grid n_s, stopped after max_iter iterations"""
    import gen_matrices
    exe = os.path.join(ROOT, "oracle", "_ref", "ecg_dump_ref")
    if not os.path.exists(exe):
        return None
    with tempfile.TemporaryDirectory() as d:
        mtx = os.path.join(d, "A.mtx")
        gen_matrices.write_mtx(mtx, OPERATORS[args.operator][1](gen_matrices, n_s))
        env = dict(os.environ, MPISHIM_NP=str(args.nsub))
        t0 = time.time()
        subprocess.run([exe, "-m", mtx, "-e", str(args.t), "-o", "0", "-r", str(args.bs_red), "-t", repr(args.tol), "-i", str(max_iter),
                        "-d", d, "-q"], check=True, env=env, stdout=subprocess.DEVNULL)
        wall = time.time() - t0
        s = json.load(open(os.path.join(d, "summary.json")))
    s["wall_s"] = wall
    return s


def cpu_baseline_sample(args):
    """cpu_baseline of the GPU line: the unmodified reference on a BOUNDED sample (a 64^3 grid of the same operator, the
    same 8 subdomains, t and tolerance: ~10 s).  Its value is the sample's own iterations/s -- nothing is scaled; the
    same-size measurement is `bench.py --impl reference` and tests/golden/big/ (the reference's run in the build container)."""
    n_s = min(args.n, args.sample_n)
    s = reference_run(args, n_s, args.max_iter)
    if s is None:
        return None
    cores = min(args.nsub, os.cpu_count() or 1)
    out = {"value": s["iter"] / s["t_solve"], "unit": UNIT, "cores": cores, "kind": "reference",
           "sample": (args.operator + " %d^3 (%d rows; the GPU line above is %d^3), S=%d ranks x 1 thread over mpishim, t=%d, tol %g: %d iterations "
                      "in %.2f s (SpMM %.2f s, block-Jacobi %.2f s; factorisation %.1f s not counted).  NOT scaled to the headline size"
                      % (n_s, n_rows(args.operator, n_s), args.n, args.nsub, args.t, args.tol, s["iter"], s["t_solve"], s["t_op"], s["t_prec"],
                         s["t_factor"])),
           "sample_grid": n_s, "sample_iterations": s["iter"], "wall_s": s["wall_s"]}
    big = os.path.join(ROOT, "tests", "golden", "big", "%s_n%d_s%d_t%d.json" % (args.operator, args.n, args.nsub, args.t))
    if os.path.exists(big):  # the reference's own run of THIS configuration, recorded in the build container (8 cores)
        g = json.load(open(big))
        out["same_config_reference_run"] = {"iterations": g["iter"], "t_solve_s": g["t_solve"], "iterations_per_s": g["iter"] / g["t_solve"],
                                            "cores": g["cores"], "cpu": g["cpu"], "where": "build container, tests/golden/big/make_big.py",
                                            "final_res": g["res"], "true_relres": g["true_relres"]}
    return out


def reference_arm(args):
    """`--impl reference`: the unmodified reference on the SAME configuration as the GPU arm (grid args.ref_n = args.n unless
    overridden), stopped after warmup + steps iterations; iterations/s = iterations / the reference's own solve time"""
    n_s = args.ref_n if args.ref_n else args.n
    iters = args.ref_iters if args.ref_iters else args.warmup + args.steps
    s = reference_run(args, n_s, iters)
    if s is None:
        return None
    cores = min(args.nsub, os.cpu_count() or 1)
    value = s["iter"] / s["t_solve"]
    ran = argparse.Namespace(**vars(args))
    ran.n = n_s  # config names the grid that ran
    sample = (args.operator + " %d^3, S=%d ranks x 1 thread over mpishim on %d host cores, t=%d, tol %g: %d iterations in %.2f s (SpMM %.2f s, "
              "block-Jacobi %.2f s); setup not counted: load + partition %.1f s, factorisation (plain-C shim for PARDISO) %.1f s"
              % (n_s, args.nsub, cores, args.t, args.tol, s["iter"], s["t_solve"], s["t_op"], s["t_prec"], s["t_build"], s["t_factor"]))
    base = {"value": value, "unit": UNIT, "cores": cores, "kind": "reference", "sample": sample, "sample_grid": n_s,
            "sample_iterations": s["iter"], "wall_s": s["wall_s"]}
    return {"metric": metric_name(args), "value": value, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1000.0 / value, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(ran), "cpu_baseline": base,
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "residual_after_%d_iterations" % s["iter"]: s["res"]}


def workload_config(args):
    rows = n_rows(args.operator, args.n)
    return {"workload": "synthetic " + OPERATORS[args.operator][2] % args.n + " (%d rows), ECG t=%d + block Jacobi, %d METIS "
            "subdomains, tol %g" % (rows, args.t, args.nsub, args.tol),
            "n": args.n, "t": args.t, "subdomains": args.nsub, "tol": args.tol, "ortho_alg": "ORTHODIR",
            "bs_red": "ADAPT_BS (whole solves only; the K timed iterations run NO_BS_RED)" if args.bs_red else "NO_BS_RED",
            "l2": "inputs larger than L2 (factor + blocks ~17 GB per pass); per-kernel timings flush L2 between repetitions"}


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum over the launches of one block-Jacobi apply, from the newest committed
    ncu capture (profiles/r*_ncu_traffic.json; ncu cannot run inside a timed bench).  Returns (bytes, file, commit it was
    taken at) or (None, None, None)."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_ncu_traffic.json")))
    for f in reversed(files):
        try:
            d = json.load(open(f))
            return float(d["traffic_bytes_per_apply"]), os.path.relpath(f, ROOT), d.get("commit")
        except (OSError, KeyError, ValueError):
            continue
    return None, None, None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--grid", dest="n", type=int, default=128, help="grid points per dimension")
    ap.add_argument("--enl", dest="t", type=int, default=8, help="enlarging factor t")
    ap.add_argument("--nsub", type=int, default=8, help="METIS subdomains = block-Jacobi blocks")
    ap.add_argument("--tol", type=float, default=1e-8)
    ap.add_argument("--operator", default="poisson7", choices=sorted(OPERATORS))
    ap.add_argument("--bs-red", type=int, default=0, choices=[0, 1], help="1 = ADAPT_BS in the whole solves (-r 1)")
    ap.add_argument("--max-iter", type=int, default=1000)
    ap.add_argument("--ref-n", type=int, default=0, help="--impl reference: grid of the reference run (default: --grid, the same configuration)")
    ap.add_argument("--ref-iters", type=int, default=0, help="--impl reference: iterations to run (default: warmup + steps)")
    ap.add_argument("--sample-n", type=int, default=64, help="grid of the bounded cpu_baseline sample inside the GPU line")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if rank != 0:
            return 0
        line = reference_arm(args)
        if line is None:
            line = {"impl": "reference", "unavailable": "oracle/_ref/ecg_dump_ref is not built (make -C oracle needs /root/reference)"}
        print(json.dumps(line))
        return 0

    # exactly ONE JSON line may reach stdout: send everything libraries print (NCCL's version banner, ...)
    # to stderr and keep the real stdout for the result
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    os.environ.setdefault("NCCL_DEBUG", "WARN")

    import ctypes as C
    import numpy as np
    import torch
    from prealps_b200 import capi

    if capi.device_count() <= 0:
        raise SystemExit("bench.py: no CUDA device; this library has no CPU path")
    if args.nsub % world != 0:
        raise SystemExit("bench.py: the number of subdomains (%d) must be a multiple of --gpus (%d)" % (args.nsub, world))
    torch.cuda.set_device(local)
    capi.lib.preAlps_b200_SetDevice(local)
    dist = None
    t_nccl = 0.0
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        uid = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            buf = (C.c_ubyte * 128)()
            assert capi.lib.preAlps_b200_NcclUniqueId(buf) == 0
            uid = torch.tensor(list(buf), dtype=torch.uint8)
        uid = uid.cuda()
        dist.broadcast(uid, 0)
        raw = bytes(uid.cpu().tolist())
        t_nccl = time.time()
        assert capi.lib.preAlps_b200_InitNccl(world, rank, raw) == 0   # includes NCCL's first collective (channel set-up)
        t_nccl = time.time() - t_nccl

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        tt = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    clk = ClockSampler(local).start()
    per = args.nsub // world
    t_setup = time.time()
    assert capi.lib.preAlps_b200_OperatorBuildStencil(OPERATORS[args.operator][0], args.n, args.nsub, rank * per, (rank + 1) * per) == 0
    t_part = time.time() - t_setup
    assert capi.lib.preAlps_b200_BlockJacobiCreate() == 0
    t_setup = time.time() - t_setup
    t_setup_max = t_setup
    arr_m = C.c_int(); arr_M = C.c_int()
    capi.lib.preAlps_OperatorGetSizes(C.byref(arr_M), C.byref(arr_m))
    m = arr_m.value
    rhs = capi.driver_rhs(m)

    # ---- e2e: whole solves through the RCI API with host buffers (the first one also warms everything up)
    barrier()
    sol, hist, info0 = capi.solve(rhs, args.t, args.tol, max_iter=args.max_iter, bs_red=args.bs_red)
    tts_all = []
    for _ in range(3):  # median of three timed solves (each one: host rhs in, ~60 iterations, host solution out)
        barrier()
        sol, hist, info = capi.solve(rhs, args.t, args.tol, max_iter=args.max_iter, bs_red=args.bs_red)
        barrier()
        tts_all.append(max_over_ranks(info.t_solve))
    tts = sorted(tts_all)[1]
    e2e_value = info.iter / tts
    t_setup_max = max_over_ranks(t_setup)

    # ---- timed region: exactly K iterations, device resident, CUDA events, max over ranks
    ms = C.c_float()
    launches = C.c_longlong()
    barrier()
    clk.wait_first_sample()
    with clk:
        assert capi.lib.preAlps_b200_BenchIterations(args.t, C.c_double(args.tol), 0, capi.dp(rhs), args.warmup, args.steps,
                                                     C.byref(ms), C.byref(launches)) == 0
        barrier()
    clk.stop()
    ms_total = max_over_ranks(float(ms.value))
    value = args.steps / (ms_total * 1e-3)

    # ---- per-kernel roofline, each timed alone: median of 11 repetitions, L2 flushed before each, CUDA events on the library stream
    peak, peak_kind = measured_peaks()
    kern = {}
    for what, name, bytes_name in ((1, "block_jacobi_apply", "bj_bytes_t%d" % args.t), (0, "spmm", "spmm_bytes_t%d" % args.t),
                                   (2, "ecg_dense_passes", None)):
        kms = C.c_float()
        assert capi.lib.preAlps_b200_BenchKernel(what, args.t, 11, 1, C.byref(kms)) == 0
        kt = max_over_ranks(float(kms.value))
        if bytes_name:
            b = capi.stat(bytes_name)
        else:
            b = 18.0 * m * args.t * 8.0  # fused dense traffic, 18 block passes per iteration
        kern[name] = {"ms": kt, "algorithmic_bytes": b, "achieved_gbs": b / (kt * 1e-3) / 1e9,
                      "frac_of_%s_peak" % peak_kind: b / (kt * 1e-3) / 1e9 / peak}
    bj = kern["block_jacobi_apply"]
    stored = capi.stat("bj_stored_bytes_t%d" % args.t)
    traffic, traffic_file, traffic_commit = ncu_traffic()
    roofline = {"bound": "hbm",
                "kernel": "block-Jacobi apply = assemble_kernel + sweep_kernel<%d> over all levels of the forest, "
                          "forward + backward (the launch group of one pcu_bj_apply: ~86 %% of an iteration)" % args.t,
                "achieved": bj["achieved_gbs"], "peak": peak, "peak_source": peak_kind, "unit": "GB/s",
                "frac": bj["achieved_gbs"] / peak, "traffic": traffic,
                "traffic_source": ("%s (ncu --set full, taken at commit %s, not during this run)" % (traffic_file, traffic_commit)) if traffic else None,
                "frac_of_nominal_8TBs": bj["achieved_gbs"] / 8000.0,
                "algorithmic_bytes_per_apply": bj["algorithmic_bytes"],
                "stored_bytes_per_apply": stored, "achieved_counting_stored_bytes": stored / (bj["ms"] * 1e-3) / 1e9,
                "note": "algorithmic bytes (SURVEY.md 8d, dense-supernode form): 2 x (8 B x EXACT nnz(L) + 8 B x rows) + 4 x m x t x 8 B; "
                        "stored bytes = what the sweeps move: the panels once per sweep with the explicit zeros of the relaxed supernodes "
                        "and the 32-row panel padding, the work vectors and the update rows; median of 11 applies timed alone, "
                        "L2 flushed before each, CUDA events on the library stream"}
    # SURVEY.md 8(d), whole solve: the algorithmic bytes of one iteration (block-Jacobi apply + SpMM + fused dense passes, per GPU)
    # over the measured time of an iteration inside the timed region
    it_bytes = sum(k["algorithmic_bytes"] for k in kern.values())
    it_gbs = it_bytes / (ms_total / args.steps * 1e-3) / 1e9
    roofline["whole_iteration"] = {"algorithmic_bytes": it_bytes, "achieved": it_gbs, "unit": "GB/s", "frac": it_gbs / peak}

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            try:
                cpu = cpu_baseline_sample(args)
            except Exception as e:  # the baseline must never take the GPU line down
                cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": "failed: %r" % (e,)}
        line = {
            "metric": metric_name(args), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(args),
            "e2e": {"value": e2e_value, "unit": UNIT,
                    "h2d_bytes_per_step": int((m * 8 + m * 4) / max(info.iter, 1)),
                    "d2h_bytes_per_step": int(m * 8 / max(info.iter, 1) + 8),
                    "time_to_solution_s": tts, "iterations": info.iter, "final_res": info.res,
                    "true_relres": info.true_relres, "true_relres_bound": args.tol * (args.t ** 0.5), "device_ms": info.t_dev_ms,
                    "time_to_solution_with_setup_s": tts + t_setup_max,
                    "note": "the reference's stopping test is ||R||_F of the ENLARGED residual <= tol ||b|| (ecg.c:264), which bounds the "
                            "true residual by tol sqrt(t) (SURVEY.md H5); the reference's own run ends at the same true residual"},
            "gpu_launches": int(launches.value),
            "clocks": clk.summary(),
            "roofline": roofline,
            "kernels": kern,
            "cpu_baseline": cpu,
            "setup": {"nccl_init_s": t_nccl, "operator_build_s": t_part, "total_s": t_setup, "bj_analysis_s": capi.stat("bj_analysis_s"),
                      "bj_factor_s": capi.stat("bj_factor_s"), "bj_nnz_exact": capi.stat("bj_nnz_exact"),
                      "bj_nnz_stored": capi.stat("bj_nnz_stored"), "bj_supernodes": capi.stat("bj_supernodes"),
                      "bj_levels": capi.stat("bj_levels"), "rows_per_gpu": m},
        }
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    barrier()
    capi.lib.preAlps_OperatorFree()
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())

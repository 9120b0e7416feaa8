#!/usr/bin/env python
"""bench.py - headline benchmark: OIS trades/sec for PV + 32-pillar delta + 32x32 gamma (FP64).

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched under torchrun)
    python bench.py --impl reference ...                      (CPU arm: the oracle port of the reference)

Workload (BASELINE.json configs[2]): 1M synthetic SONIA OIS (1Y-50Y, annual) per GPU on the
README 32-pillar LINEAR_ZERO_RATES curve; every step values the whole book: per-trade PV,
delta[32] and gamma[32x32] written to HBM plus the portfolio totals (the Portfolio.compute
result), all-reduced over ranks when N > 1 (weak scaling: each rank owns its own 1M trades).

Printed JSON keys follow the driver contract; additionally
  roofline      dominant kernel (per-trade expansion, HBM-write bound) against MEASURED_PEAKS.json
  cpu_baseline  oracle/liboracle.so (C port of the reference algorithm) on this box's host cores
  e2e           same metric through the reference-facing call with HOST buffers: H2D of the flattened
                book from pinned memory + valuation + D2H of the portfolio totals, every step
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

BYTES_PER_TRADE = 9712          # 1256 B explicit-cashflow input + 8 + 256 + 8192 B output
FLOPS_NOTE = "see DESIGN.md section 5"
METRIC = "OIS trades/sec PV+delta+gamma FP64"
WORKLOAD = ("BASELINE configs[2]: 1M synthetic SONIA OIS (1Y-50Y annual, 50% forward-starting) per GPU, "
            "PV + 32-pillar delta + full 32x32 gamma incl. par-rate Jacobian chain, per-trade outputs written")


def load_curve():
    from adrates_b200.curves import OISCurve
    from adrates_b200.global_types import InterpTypes
    from tests.util_trades import make_calibration_swaps
    with open(os.path.join(ROOT, "tests", "golden", "ref_curves.json")) as f:
        cv = json.load(f)["gbp_readme_lzr"]
    vd, swaps = make_calibration_swaps(cv)
    return cv, OISCurve(vd, swaps, InterpTypes[cv["interp"]])


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu, self.rows, self.stop_flag, self.proc = gpu_index, [], False, None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            for line in self.proc.stdout:
                self.rows.append(line.strip())
                if self.stop_flag:
                    break
        except Exception:  # noqa: BLE001
            pass

    def finish(self):
        self.stop_flag = True
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            p = [x.strip() for x in r.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1]))
                mx.append(float(p[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        busy = sorted(sm)[len(sm) // 2:]
        return {"sm_mhz": float(np.median(busy)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:  # noqa: BLE001
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def cpu_oracle_rate(cv, curve, book, sample, dense, threads=0):
    """trades/s of the C oracle on `sample` trades of the book (all host threads by default)."""
    from oracle import cavour_oracle as orc, c_oracle
    from adrates_b200.synthetic import reference_leg_tables
    plan = orc.plan_path_b(cv["swap_times"], cv["year_fracs"])
    d, J, C = orc.bootstrap_tables(cv["swap_rates"], plan)
    lt = reference_leg_tables(book)
    tr = dict(sched=book.sched[:sample], coupon=book.coupon[:sample], notional=book.notional[:sample],
              spread=book.spread[:sample], fixed_sign=book.fixed_sign[:sample])
    if not threads:
        # all the host cores this process may use: torchrun exports OMP_NUM_THREADS=1 to every rank, which would
        # silently turn the CPU baseline into a single-thread number
        threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    t0 = time.perf_counter()
    out = c_oracle.ois_batch((plan["times"], d, J, C), curve._interp_type.value, lt, tr, want=7, dense=dense,
                             n_threads=threads)
    dt = time.perf_counter() - t0
    return sample / dt, threads, out


def run_reference(args):
    """CPU arm: the reference's algorithm (oracle port; the reference itself needs JAX, which is not
    installed) on a bounded sample of the same workload, all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cv, curve = load_curve()
    from adrates_b200.synthetic import make_book
    sample = args.ref_sample
    book = make_book(curve, max(sample, 1000), seed=20240430)
    rates = []
    cores = 0
    for i in range(args.warmup + args.steps):
        r, cores, _ = cpu_oracle_rate(cv, curve, book, sample, dense=True)
        if i >= args.warmup:
            rates.append(r)
    value = float(np.mean(rates))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "trades/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sample / value, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample_trades_per_step": sample},
        "cpu_baseline": {"value": value, "unit": "trades/s", "cores": cores, "kind": "port",
                         "sample": f"{sample} trades/step of the same book; oracle/liboracle.so dense chain rule "
                                   "(reference algorithm, curve tables built once), OpenMP over all host threads"},
        "e2e": {"value": value, "unit": "trades/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--trades", type=int, default=1_000_000, help="trades per GPU")
    ap.add_argument("--layout", default="dedup", choices=["dedup", "private"],
                    help="dedup: trades share schedule units (product default); private: one unit per trade")
    ap.add_argument("--ref-sample", type=int, default=2000)
    ap.add_argument("--cpu-sample", type=int, default=4000)
    ap.add_argument("--no-extra", action="store_true", help="skip the secondary measurements")
    ap.add_argument("--scenarios", type=int, default=2000, help="shocked curves for the config-4 extra")
    args = ap.parse_args()
    # A benchmark must never hold a GPU box hostage: if anything (a lost rank, a collective only some ranks reach)
    # stalls the run, leave with an error instead of waiting for the launcher's limit.
    deadline = float(os.environ.get("BENCH_DEADLINE_S", "900"))
    if "BENCH_DEADLINE_S" not in os.environ and args.impl == "reference":
        deadline = max(deadline, 300.0 + 0.6 * (args.steps + args.warmup))    # ~0.2 s of CPU work per step
    if deadline > 0:
        import threading

        def _expired():
            sys.stderr.write(f"bench.py: no result after {deadline:.0f} s (BENCH_DEADLINE_S) - aborting this rank\n")
            sys.stderr.flush()
            os._exit(124)
        t = threading.Timer(deadline, _expired)
        t.daemon = True
        t.start()
    if args.impl == "reference":
        return run_reference(args)

    # libraries (NCCL banner, warnings) must not pollute the one-line JSON contract: route fd 1 to
    # stderr for the whole run and print the result on the saved descriptor
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        os.write(json_fd, (json.dumps(obj) + "\n").encode())

    import torch
    import torch.distributed as dist
    from adrates_b200 import _native
    from adrates_b200.synthetic import make_book, flatten_book

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the valuation path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = None
    if world > 1:
        # one rank per GPU: keep the rank (its pinned upload buffers and the library's host threads) on the CPUs next to
        # its GPU, so that eight end-to-end uploads do not all cross the socket interconnect
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(torch.cuda.current_device() if "CUDA_VISIBLE_DEVICES" not in os.environ
                                                  else int(os.environ["CUDA_VISIBLE_DEVICES"].split(",")[local]))
            pynvml.nvmlDeviceSetCpuAffinity(h)
            numa = len(os.sched_getaffinity(0))
        except Exception as ex:          # noqa: BLE001  (placement is an optimisation, never a requirement)
            numa = f"unbound ({type(ex).__name__})"
        dist.init_process_group("nccl", device_id=dev)

    cv, curve = load_curve()
    n = args.trades
    t0 = time.perf_counter()
    book = make_book(curve, n, seed=20240430 + rank)       # weak scaling: every rank has its own book
    flat = flatten_book(book, dedup=(args.layout == "dedup"))
    flatten_s = time.perf_counter() - t0

    # pinned host copies of the flattened book (the e2e arm uploads from these every step)
    def pin(a):
        if a is None:
            return None
        t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        return t
    pinned = {k: pin(getattr(flat, k)) for k in ("unit_offsets", "amt", "weight", "node", "comp_weight",
                                                 "group_offsets", "group_units", "out_index", "unit_weight")}
    import copy
    flat_pinned = copy.copy(flat)
    for k, t in pinned.items():
        setattr(flat_pinned, k, None if t is None else t.numpy())

    ctx = _native.Context(local)
    stream = torch.cuda.current_stream(dev)
    ctx.set_stream(stream.cuda_stream)                     # kernels, events and NCCL order on one stream
    ctx.curve_build(curve._interp_type.value, curve.swap_rates, curve.path_b_plan(), order=2)
    ctx.portfolio_upload(flat_pinned)
    pv = torch.empty(n, dtype=torch.float64, device=dev)
    dl = torch.empty(n, 32, dtype=torch.float64, device=dev)
    gm = torch.empty(n, 32, 32, dtype=torch.float64, device=dev)
    agg = torch.zeros(_native.NOUT, dtype=torch.float64, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2
    MASK = _native.REQ_VALUE | _native.REQ_DELTA | _native.REQ_GAMMA

    def step():
        ctx.portfolio_value(MASK, pv.data_ptr(), dl.data_ptr(), gm.data_ptr(), agg.data_ptr())
        if world > 1:
            dist.all_reduce(agg)          # portfolio PV / ladder / gamma: 1057 doubles over NVLink

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()

    # ---- parity gate on a sample of this rank's book before any number is reported ----
    gate = None
    if rank == 0:
        _, _, (pv_c, dl_c, gm_c) = cpu_oracle_rate(cv, curve, book, 256, dense=False)
        N = book.notional[:256]
        e_pv = np.max(np.abs(pv[:256].cpu().numpy() - pv_c) / np.maximum(np.abs(pv_c), N))
        e_dl = np.max(np.abs(dl[:256, :].cpu().numpy() - dl_c) / np.maximum(np.abs(dl_c), (N * 1e-4)[:, None]))
        e_gm = np.max(np.abs(gm[:256].cpu().numpy() - gm_c) / np.maximum(np.abs(gm_c), (N * 1e-8)[:, None, None]))
        gate = float(max(e_pv, e_dl, e_gm))
        if not gate < 1e-10:
            raise SystemExit(f"parity gate failed: scaled error {gate:.3e} >= 1e-10")

    # ---- timed region: K steps, per-step CUDA events, L2 flushed between steps ----
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.3)
    ctx.profile(True)
    launches0 = ctx.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    kern_ms = []
    barrier()
    wall0 = time.perf_counter()
    for a, b in ev:
        flush.zero_()
        a.record(stream)
        step()
        b.record(stream)
        b.synchronize()
        kern_ms.append(ctx.last_kernel_ms())
    barrier()
    wall = time.perf_counter() - wall0
    launches = ctx.launch_count() - launches0
    ctx.profile(False)
    clocks = sampler.finish()
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms.item())
    ms_per_step = total_ms / args.steps
    value = world * n / (ms_per_step * 1e-3)

    # ---- e2e: host buffers -> H2D -> valuation -> D2H of portfolio totals, every step ----
    agg_host = np.empty(_native.NOUT)
    agg_resident = agg.cpu().numpy().copy() if world == 1 else None
    gm_check = float(gm[:: max(1, n // 4096)].sum().item())
    ctx.set_async_upload(True)       # per-trade arrays stream in behind the units kernel (pinned buffers stay alive)
    for _ in range(2):
        ctx.portfolio_upload(flat_pinned)
        ctx.portfolio_value_host(MASK, pv.data_ptr(), dl.data_ptr(), gm.data_ptr(), agg_host)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ctx.portfolio_upload(flat_pinned)
        if world > 1:
            ctx.portfolio_value(MASK, pv.data_ptr(), dl.data_ptr(), gm.data_ptr(), agg.data_ptr())
            dist.all_reduce(agg)
            agg_host[:] = agg.cpu().numpy()
        else:
            ctx.portfolio_value_host(MASK, pv.data_ptr(), dl.data_ptr(), gm.data_ptr(), agg_host)
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = world * n * args.steps / float(e2e_s.item())
    ctx.set_async_upload(False)
    # the pipelined path must reproduce the resident-input results exactly (same kernels, same order)
    if agg_resident is not None and not np.array_equal(agg_host, agg_resident):
        raise SystemExit("e2e (pipelined upload) totals differ from the device-resident totals")
    if float(gm[:: max(1, n // 4096)].sum().item()) != gm_check:
        raise SystemExit("e2e (pipelined upload) gamma rows differ from the device-resident run")

    # ---- secondary measurements (same book, device-resident inputs; reported under "extras") ----
    extras = {}
    if world > 1 and not args.no_extra:
        # per-GPU side measurements: the N=1 run reports them; a multi-rank run keeps to the contract line (and to code
        # every rank executes - a rank-0-only measurement must never sit behind a collective)
        extras = {"skipped": "secondary per-GPU measurements are reported by the single-GPU run"}
    if not args.no_extra and world == 1:
        def timed(fn, reps=10, all_ranks=True):
            # all_ranks=False: measurements only rank 0 takes - no collective there (the other ranks have left)
            for _ in range(3):
                fn()
            if all_ranks:
                barrier()
            else:
                torch.cuda.synchronize(dev)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            for _ in range(reps):
                fn()
            b.record(stream)
            b.synchronize()
            return a.elapsed_time(b) / reps
        M_PD = _native.REQ_VALUE | _native.REQ_DELTA
        ms = timed(lambda: ctx.portfolio_value(M_PD, pv.data_ptr(), dl.data_ptr(), None, agg.data_ptr()))
        extras["pv_delta_config2"] = {"ms_per_step": ms, "trades_per_s_per_gpu": n / ms * 1e3,
                                      "hbm_frac_algorithmic_1520B": n * 1520 / (ms * 1e-3) / 1e9 / measured_peak()[0]}
        ms = timed(lambda: ctx.portfolio_value(MASK, None, None, None, agg.data_ptr()))
        extras["portfolio_totals_only"] = {"ms_per_step": ms, "trades_per_s_per_gpu": n / ms * 1e3,
                                           "note": "PV+delta+gamma of the portfolio, no per-trade rows written"}
        if rank == 0:
            # chain rule as a DMMA GEMM on the private layout (one node-gradient row per trade)
            try:
                ng = min(n, 1_000_000)
                sub_b = type(book)(curve, book.schedules, book.sched[:ng], book.coupon[:ng], book.notional[:ng],
                                   book.fixed_sign[:ng], book.spread[:ng])
                priv = flatten_book(sub_b, dedup=False, sort_units=False)
                ctx3 = _native.Context(local)
                ctx3.set_stream(stream.cuda_stream)
                ctx3.curve_build(curve._interp_type.value, curve.swap_rates, curve.path_b_plan(), order=1)
                ctx3.portfolio_upload(priv)
                best = None
                for _ in range(5):
                    g_ms, g_fl = ctx3.portfolio_delta_gemm(pv.data_ptr(), dl.data_ptr())
                    best = g_ms if best is None else min(best, g_ms)
                tot_ms = timed(lambda: ctx3.portfolio_delta_gemm(pv.data_ptr(), dl.data_ptr()), reps=5, all_ranks=False)
                extras["chain_gemm_dmma"] = {
                    "units": ng, "gemm_ms": best, "gemm_tflops": g_fl / (best * 1e-3) / 1e12,
                    "dmma_peak_tflops_measured": 37.13, "tensor_pipe_frac": g_fl / (best * 1e-3) / 1e12 / 37.13,
                    "pv_delta_total_ms": tot_ms,
                    "note": "delta[U][32] = Q[U][264] x (1e-4 J/d)[264][32], mma.sync.m8n8k4.f64; dense formulation "
                            "(2.6x the flops of the fused sparse chain), reported as an alternative, not the default path"}
                ctx3.close()
                del priv
            except Exception as ex:  # noqa: BLE001
                extras["chain_gemm_dmma"] = {"error": str(ex)}
            from adrates_b200.synthetic import shocked_rate_scenarios
            S = args.scenarios
            nt = min(n, 100_000)
            shocked = shocked_rate_scenarios(curve, S)
            sub = flatten_book(type(book)(curve, book.schedules, book.sched[:nt], book.coupon[:nt], book.notional[:nt],
                                         book.fixed_sign[:nt], book.spread[:nt]), dedup=True)
            ctx2 = _native.Context(local)
            ctx2.set_stream(stream.cuda_stream)
            ctx2.curve_build(curve._interp_type.value, curve.swap_rates, curve.path_b_plan(), order=0)
            ctx2.portfolio_upload(sub)
            pnl = torch.empty(S, nt, dtype=torch.float64, device=dev)
            ms = timed(lambda: ctx2.scenarios(shocked, pnl.data_ptr()), reps=3, all_ranks=False)
            extras["scenarios_config4"] = {"scenarios": S, "trades": nt, "ms": ms, "revaluations_per_s": S * nt / ms * 1e3,
                                           "pnl_bytes": S * nt * 8,
                                           "note": "shocked curves re-bootstrapped on device (DFs only) + full revaluation (one exp per "
                                                   "distinct DF query and scenario, units gather them); includes H2D of the shocked rates"}
            del pnl
            ctx2.close()
            # XCCY basis swaps (BASELINE config 5, cross-currency half): 500k GBP/USD swaps on SONIA + SOFR + the
            # GBP/USD basis curve, PV + domestic / foreign / basis ladders per trade (VALUE + DELTA, like the reference)
            try:
                from adrates_b200.synthetic_xccy import make_xccy_book, XccyBookValuer
                from tests.conftest import load_golden
                from tests.util_xccy import build_xccy_model
                gx = load_golden("ref_xccy.json")
                t1 = time.perf_counter()
                xbook = make_xccy_book(build_xccy_model(gx), 500_000, seed=11, spot=gx["spot_fx"])
                xval = XccyBookValuer(xbook, device=local, stream=stream.cuda_stream)
                x_prep = time.perf_counter() - t1
                ms = timed(lambda: xval.value(), reps=10, all_ranks=False)
                extras["xccy_config5"] = {"trades": xbook.n_trades, "ms_per_step": ms, "trades_per_s": xbook.n_trades / ms * 1e3,
                                          "units": int(xval.flat_for.n_units), "flatten_seconds_untimed": x_prep,
                                          "note": "per-trade PV + three 32-wide ladder rows written (776 B/trade); parity of "
                                                  "this path against the per-trade engine path: tests/test_gpu_xccy_book.py"}
            except Exception as ex:  # noqa: BLE001
                extras["xccy_config5"] = {"error": str(ex)}
            # ZCIS leg PVs (BASELINE config 5, inflation half): 500k swaps = 1M cashflows on the path-A nodes of the
            # same curve, host buffers in, per-trade PVs + total out (cav_cashflow_pv)
            try:
                nz = 500_000
                rz = np.random.Generator(np.random.PCG64(11))
                tz = np.repeat(rz.integers(1, 31, nz).astype(np.float64) + rz.uniform(0.0, 0.02, nz), 2)
                az = rz.uniform(-1e6, 1e6, 2 * nz)
                oz = np.arange(0, 2 * nz + 1, 2, dtype=np.int64)
                ctxz = _native.Context(local)
                zargs = (curve._interp_type.value, curve._times, curve._dfs, 0.0, oz, tz, az)
                for _ in range(2):
                    pvz, totz = ctxz.cashflow_pv(*zargs)
                t1 = time.perf_counter()
                for _ in range(5):
                    pvz, totz = ctxz.cashflow_pv(*zargs)
                dtz = (time.perf_counter() - t1) / 5
                ref = np.array([curve._node_df(float(x)) for x in tz[:64]])
                errz = float(np.max(np.abs(pvz[:32] - (az[:64] * ref).reshape(-1, 2).sum(1)) / 1e6))
                extras["zcis_cashflow_pv_config5"] = {"trades": nz, "cashflows": 2 * nz, "ms_e2e_host_buffers": dtz * 1e3,
                                                      "trades_per_s": nz / dtz, "check_scaled_err_vs_host_df": errz,
                                                      "note": "discounting of the ZCIS legs on the path-A curve; the CPI index "
                                                              "arithmetic that produces the amounts is host logic"}
                ctxz.close()
            except Exception as ex:  # noqa: BLE001
                extras["zcis_cashflow_pv_config5"] = {"error": str(ex)}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peak()
    k = np.array(kern_ms)                      # [steps, 3] units / expand / totals
    dom = int(np.argmax(k.mean(0)))
    dom_name = ["k_units_mma (fused interpolation + PV + delta + gamma per schedule unit, FP64 DMMA tiles)",
                "k_expand (per-trade PV/delta/gamma rows from unit results, streaming stores)",
                "k_reduce_partials"][dom]
    dom_ms = float(k[:, dom].mean())
    achieved = n * BYTES_PER_TRADE / (dom_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": dom_name, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                "kernel_ms": dom_ms, "kernel_share_of_step": dom_ms / ms_per_step,
                "algorithmic_bytes_per_launch": n * BYTES_PER_TRADE,
                "step_frac": (n * BYTES_PER_TRADE / (ms_per_step * 1e-3) / 1e9) / peak,
                "all_kernels_ms": {"k_units": float(k[:, 0].mean()), "k_expand": float(k[:, 1].mean()),
                                   "k_reduce_partials": float(k[:, 2].mean())}}
    prof = os.path.join(ROOT, "profiles", "r01_traffic.json")
    if os.path.exists(prof):
        try:
            with open(prof) as f:
                roofline["traffic"] = json.load(f).get(args.layout)
        except Exception:  # noqa: BLE001
            pass
    if roofline["traffic"]:
        # DRAM bytes actually moved per launch (ncu) over the live kernel time.  The algorithmic figure above
        # charges every trade its own 1 256 B of cashflow input; the dedup layout reads shared schedule units
        # once, so `frac` can exceed the physical rate (and 1.0) there - both are reported.
        roofline["physical_gbs"] = roofline["traffic"] * (n / 1_000_000) / (dom_ms * 1e-3) / 1e9
        roofline["physical_frac"] = roofline["physical_gbs"] / peak

    cpu_dense, cores, _ = cpu_oracle_rate(cv, curve, book, min(args.cpu_sample, n), dense=True)
    cpu_sparse, _, _ = cpu_oracle_rate(cv, curve, book, min(20 * args.cpu_sample, n), dense=False)
    cpu_baseline = {"value": cpu_dense, "unit": "trades/s", "cores": cores, "kind": "port",
                    "sample": f"first {min(args.cpu_sample, n)} trades of the same book, oracle/liboracle.so with the "
                              "reference's dense chain rule (J^T H J + sum g_k C_k per leg), curve tables built once, "
                              "OpenMP over all host threads",
                    "value_sparse_port": cpu_sparse,
                    "sparse_note": "same arithmetic skipping structurally-zero rows (fastest CPU port we have)"}

    line = {
        "metric": METRIC, "value": value, "unit": "trades/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "trades_per_gpu": n, "layout": args.layout,
                   "units": flat.n_units, "terms": flat.n_terms, "groups": flat.n_groups,
                   "l2": "flushed (256 MiB memset) before every timed step; each step also streams "
                         f"{n * 8456 / 1e9:.2f} GB of outputs (>> 126 MB L2)",
                   "timing": "per-step CUDA events on the launch stream, summed over K steps, max over ranks",
                   "parity_gate_scaled_err": gate, "flatten_seconds_untimed": flatten_s,
                   "wall_seconds_timed_region": wall},
        "clocks": clocks,
        "host_cpus_per_rank": numa,
        "e2e": {"value": e2e_value, "unit": "trades/s", "h2d_bytes_per_step": flat.h2d_bytes(),
                "d2h_bytes_per_step": _native.NOUT * 8,
                "note": "cav_portfolio_upload from pinned host arrays + cav_portfolio_value(_host) + totals D2H; "
                        "per-trade rows stay in HBM"},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "cpu_baseline": cpu_baseline,
        "extras": extras,
    }
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

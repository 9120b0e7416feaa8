#!/usr/bin/env python
"""bench.py - headline benchmark: OIS trades/sec for PV + 32-pillar delta + 32x32 gamma (FP64).

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched under torchrun)
    python bench.py --impl reference ...                      (CPU arm: the oracle port of the reference)

Workload (BASELINE.json configs[2]): 1M synthetic SONIA OIS (1Y-50Y, annual) per GPU on the README 32-pillar
LINEAR_ZERO_RATES curve; every step values the whole book: per-trade PV, delta[32] and gamma[32x32] written to HBM plus the
portfolio totals (the Portfolio.compute result), summed over the ranks inside the totals kernel over NVLink when N > 1 (weak
scaling: each rank owns its own 1M trades).

Printed JSON keys follow the driver contract; additionally
  roofline      dominant kernel (per-trade expansion, HBM-write bound) against MEASURED_PEAKS.json; `frac` is on the bytes the
                kernel must physically move, `frac_algorithmic` on SURVEY 8(d)'s 9 712 B/trade
  cpu_baseline  oracle/liboracle.so (C port of the reference algorithm) on this box's host cores
  e2e           the same metric through the public call `OISBook.from_arrays(host arrays).compute([VALUE, DELTA, GAMMA])`:
                H2D of the per-trade arrays from pinned memory, device-side flattening (schedules, day counts, brackets, tile
                plan), valuation and D2H of the portfolio totals, every step
  extras        secondary measurements (config 2 / 4 / 5, strong scaling, the pre-flattened upload path, ...)
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
METRIC = "OIS trades/sec PV+delta+gamma FP64"
WORKLOAD = ("BASELINE configs[2]: 1M synthetic SONIA OIS (1Y-50Y annual, 50% forward-starting) per GPU, "
            "PV + 32-pillar delta + full 32x32 gamma incl. par-rate Jacobian chain, per-trade outputs written")


def load_curve():
    """(cv, curve): the README SONIA curve and the dict of engine inputs the oracle takes."""
    from adrates_b200.market_data import readme_gbp_curve
    curve = readme_gbp_curve()
    cv = {"swap_rates": list(curve.swap_rates), "swap_times": list(curve.swap_times), "year_fracs": [list(f) for f in curve.year_fracs],
          "interp": curve._interp_type.name}
    return cv, curve


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index, enabled=True, period_ms=100):
        # ONE sampler per job (rank 0 polls every GPU of the job): a polling nvidia-smi per rank makes eight processes take the
        # driver's locks ten times a second each, which stretched every launch-bound phase of the other ranks' end-to-end steps
        # (device flattening 1.37 -> 2.1 ms per book at 8 ranks)
        super().__init__(daemon=True)
        self.gpu, self.rows, self.stop_flag, self.proc = gpu_index, [], False, None
        self.enabled, self.period_ms = enabled, period_ms

    def run(self):
        if not self.enabled:
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", str(self.period_ms), "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
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


def host_threads():
    # all the host cores this process may use: torchrun exports OMP_NUM_THREADS=1 to every rank, which would silently turn
    # the CPU baseline into a single-thread number
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


def oracle_tables(cv):
    from oracle import cavour_oracle as orc
    plan = orc.plan_path_b(cv["swap_times"], cv["year_fracs"])
    d, J, C = orc.bootstrap_tables(cv["swap_rates"], plan)
    return (plan["times"], d, J, C)


def book_trades(book, lo, hi):
    return dict(sched=book.sched[lo:hi], coupon=book.coupon[lo:hi], notional=book.notional[lo:hi], spread=book.spread[lo:hi],
                fixed_sign=book.fixed_sign[lo:hi])


def cpu_oracle_rate(cv, curve, book, sample, dense, threads=0):
    """trades/s of the C oracle on `sample` trades of the book (all host threads by default)."""
    from oracle import c_oracle
    from adrates_b200.synthetic import reference_leg_tables
    lt = reference_leg_tables(book)
    threads = threads or host_threads()
    t0 = time.perf_counter()
    out = c_oracle.ois_batch(oracle_tables(cv), curve._interp_type.value, lt, book_trades(book, 0, sample), want=7, dense=dense,
                             n_threads=threads)
    dt = time.perf_counter() - t0
    return sample / dt, threads, out


def cpu_units_rate(cv, curve, book, sample, threads=0):
    """trades/s of the CPU port WITH the unit factorisation the GPU path uses: both legs of every distinct schedule of the
    sample valued once (sparse chain rule), every trade expanded as a weighted sum of two unit rows."""
    from oracle import c_oracle
    from adrates_b200.synthetic import reference_leg_tables
    lt = reference_leg_tables(book)
    threads = threads or host_threads()
    R = len(cv["swap_rates"])
    out = (np.empty(sample), np.empty((sample, R)), np.empty((sample, R, R)))     # allocated (and touched by numpy) untimed
    for a in out:
        a.fill(0.0)
    t0 = time.perf_counter()
    pv, dl, gm, t_units, t_expand = c_oracle.ois_batch_units(oracle_tables(cv), curve._interp_type.value, lt,
                                                             book_trades(book, 0, sample), want=7, n_threads=threads, out=out)
    dt = time.perf_counter() - t0
    return sample / dt, t_units, t_expand, (pv, dl, gm)


def run_reference(args):
    """CPU arm: the reference's algorithm (oracle port; the reference itself needs JAX, which is not installed) on a
    bounded sample of the same workload, all host threads."""
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


def git_commit():
    try:
        return subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True, timeout=5).stdout.strip() or None
    except Exception:  # noqa: BLE001
        return None


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
    ap.add_argument("--scenarios", type=int, default=10_000, help="shocked curves of the config-4 extra (whole job)")
    args = ap.parse_args()
    # A benchmark must never hold a GPU box hostage: if anything (a lost rank, a collective only some ranks reach)
    # stalls the run, leave with an error instead of waiting for the launcher's limit.
    deadline = float(os.environ.get("BENCH_DEADLINE_S", "900"))
    if "BENCH_DEADLINE_S" not in os.environ and args.impl == "reference":
        deadline = max(deadline, 300.0 + 0.6 * (args.steps + args.warmup))    # ~0.2 s of CPU work per step
    if deadline > 0:
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
    from adrates_b200 import RequestTypes, _native
    from adrates_b200.batch import OISBook
    from adrates_b200.parallel import init_device_allreduce
    from adrates_b200.position import CurveSession
    from adrates_b200.synthetic import make_array_book, make_book, flatten_book

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

    ALL = [RequestTypes.VALUE, RequestTypes.DELTA, RequestTypes.GAMMA]
    MASK = _native.REQ_VALUE | _native.REQ_DELTA | _native.REQ_GAMMA
    cv, curve = load_curve()
    n = args.trades
    seed = 20240430 + rank                                  # weak scaling: every rank has its own book
    t0 = time.perf_counter()
    abook = make_array_book(curve, n, seed=seed)            # per-trade arrays (what a user holds)
    arrays_s = time.perf_counter() - t0
    book = make_book(curve, n, seed=seed)                   # the same trades over schedule objects: input of the CPU oracle

    stream = torch.cuda.current_stream(dev)
    ctx = _native.Context(local)
    ctx.set_stream(stream.cuda_stream)                      # kernels and events order on one stream
    ctx.curve_build(curve._interp_type.value, curve.swap_rates, curve.path_b_plan(), order=2)
    all_ranks = world > 1 and init_device_allreduce(ctx, rank, world)
    VMASK = MASK | (_native.REQ_ALLREDUCE if all_ranks else 0)
    t0 = time.perf_counter()
    flat = None
    if args.layout == "dedup":
        where = abook.upload(ctx)                           # device-side flattening (cav_book_from_arrays)
        ctx.sync()
    else:
        flat = flatten_book(book, dedup=False)
        ctx.portfolio_upload(flat)
        where = "host"
    flatten_s = time.perf_counter() - t0
    info = ctx.book_info()
    pv = torch.empty(n, dtype=torch.float64, device=dev)
    dl = torch.empty(n, 32, dtype=torch.float64, device=dev)
    gm = torch.empty(n, 32, 32, dtype=torch.float64, device=dev)
    agg = torch.zeros(_native.NOUT, dtype=torch.float64, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def step():
        # portfolio PV / ladder / gamma of ALL ranks: summed inside the totals kernel over NVLink peer memory
        ctx.portfolio_value(VMASK, pv.data_ptr(), dl.data_ptr(), gm.data_ptr(), agg.data_ptr())

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
    job_gpus = ",".join(str(g) for g in range(world)) if "CUDA_VISIBLE_DEVICES" not in os.environ else \
        ",".join(os.environ["CUDA_VISIBLE_DEVICES"].split(",")[:world])
    sampler = ClockSampler(job_gpus if world > 1 else local, enabled=(rank == 0), period_ms=100 if world == 1 else 200)
    sampler.start()
    t_wait = time.perf_counter()
    while rank == 0 and not sampler.rows and time.perf_counter() - t_wait < 5.0:      # nvidia-smi needs a moment before its first sample
        time.sleep(0.05)
    sampler.rows.clear()                                                # keep only samples taken under load
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
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms.item())
    ms_per_step = total_ms / args.steps
    value = world * n / (ms_per_step * 1e-3)
    agg_resident = agg.cpu().numpy().copy()
    gm_check = float(gm[:: max(1, n // 4096)].sum().item())

    # ---- e2e: the public call from HOST arrays, every step:  OISBook.from_arrays(...).compute([VALUE, DELTA, GAMMA]) ----
    # H2D of the per-trade arrays (pinned), device-side flattening, valuation, D2H of the totals (summed over the ranks in the
    # totals kernel).  Result rows stay in HBM (caller-owned tensors, reused).
    def pin(a):
        return torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
    # (a book as a risk system would hold it: int32 day serials, int32 tenors, int8 sides, float64 coupons and notionals - the
    # device flattener reads the narrow arrays as they are: 25 bytes per trade over the host link)
    h_eff, h_tenor, h_sign, h_cpn, h_notl = (pin(a) for a in (abook.effective.astype(np.int32), abook._tenor,
                                                              abook.fixed_sign.astype(np.int8), abook.coupon, abook.notional))
    conv = dict(fixed_freq_type=abook.fixed_freq_type, fixed_dc_type=abook.fixed_dc_type, float_freq_type=abook.float_freq_type,
                float_dc_type=abook.float_dc_type, bd_type=abook.bd_type)
    out_rows = {"pv": pv, "delta": dl, "gamma": gm}
    e2e_dedup = args.layout == "dedup"

    def e2e_step():
        b = OISBook.from_arrays(curve, h_eff, tenor_years=h_tenor, fixed_sign=h_sign, fixed_coupon=h_cpn, notional=h_notl, **conv)
        return b.compute(ALL, device=local, dedup=e2e_dedup, all_ranks=world > 1, out=out_rows)[0]

    e2e_steps = args.steps if e2e_dedup else min(args.steps, 3)      # private units are flattened on the host: seconds per step
    for _ in range(2 if e2e_dedup else 1):
        res = e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        res = e2e_step()
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = world * n * e2e_steps / float(e2e_s.item())
    if rank == 0 and len(sampler.rows) < 3:              # a short timed region: keep the GPU under the same load until there are samples
        t_wait = time.perf_counter()
        while len(sampler.rows) < 3 and time.perf_counter() - t_wait < 3.0:
            # (rank-local filler: no exchange - the ranks do not run the same number of these)
            ctx.portfolio_value(MASK, pv.data_ptr(), dl.data_ptr(), gm.data_ptr(), agg.data_ptr())
            torch.cuda.synchronize(dev)
    clocks = sampler.finish()              # sampled over the device-timed steps and the end-to-end steps (both under load)
    e2e_h2d = int(sum(a.nbytes for a in (h_eff, h_tenor, h_sign, h_cpn, h_notl)))
    # the public path must reproduce the resident-input results (same kernels; totals differ only in summation order)
    tot_scale = float(np.sum(np.abs(book.notional))) * world
    if abs(res.value.amount - agg_resident[0]) > 1e-12 * tot_scale:
        raise SystemExit("e2e (from_arrays) portfolio PV differs from the device-resident run")
    if e2e_dedup and float(gm[:: max(1, n // 4096)].sum().item()) != gm_check:
        raise SystemExit("e2e (from_arrays) gamma rows differ from the device-resident run")

    # ---- secondary measurements ----
    extras = {}

    def timed(fn, reps=10, sync_all=False):
        for _ in range(3):
            fn()
        if sync_all:
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

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    peak, peak_src = measured_peak()
    if not args.no_extra:
        # ---- measurements EVERY rank takes (multi-rank runs execute exactly this code on every rank) ----
        # (1) strong scaling of the north-star target: ONE 1M-trade book over all ranks (compute_distributed: shards balanced
        # by coupon count, totals summed in the totals kernel), per-trade rows of the shard written
        try:
            sbook = make_array_book(curve, n, seed=20240430)        # the same book on every rank
            shard, _ = sbook.shard_by_schedule(rank, world)       # whole schedules per rank: the units stage is not repeated
            sctx = CurveSession.get(curve, local).ctx
            shard.upload(sctx)
            if world > 1:
                init_device_allreduce(sctx, rank, world)
            ns = shard.n_trades
            smask = MASK | (_native.REQ_ALLREDUCE if world > 1 else 0)
            sagg = torch.zeros(_native.NOUT, dtype=torch.float64, device=dev)
            sctx.set_stream(stream.cuda_stream)
            ms = timed(lambda: sctx.portfolio_value(smask, pv.data_ptr(), dl.data_ptr(), gm.data_ptr(), sagg.data_ptr()),
                       reps=10, sync_all=True)
            ms = max_over_ranks(ms)
            extras["strong_scaling_1m_book"] = {
                "ranks": world, "trades_total": n, "trades_this_rank": ns, "ms_per_step": ms, "trades_per_s": n / ms * 1e3,
                "note": "one 1M-trade book sharded over the ranks by schedule (OISBook.shard_by_schedule: whole schedules per rank, balanced by coupon count), per-trade rows of "
                        "the shard written, totals of the WHOLE book on every rank via the in-kernel NVLink all-reduce; "
                        "back-to-back steps without an L2 flush, max over ranks"}
        except Exception as ex:  # noqa: BLE001
            extras["strong_scaling_1m_book"] = {"error": repr(ex)}
        if world > 1:
            barrier()
        # (2) BASELINE config 4 in full: 10 000 shocked curves x 100 000 trades, scenarios sharded over the ranks (no collective)
        try:
            from adrates_b200.scenarios import scenario_bounds
            from adrates_b200.synthetic import shocked_rate_scenarios
            S = args.scenarios
            nt = min(n, 100_000)
            shocked = shocked_rate_scenarios(curve, S)
            lo, hi = scenario_bounds(S, world)[rank]
            sub = make_array_book(curve, nt, seed=20240430)
            ctx2 = _native.Context(local)
            ctx2.set_stream(stream.cuda_stream)
            ctx2.curve_build(curve._interp_type.value, curve.swap_rates, curve.path_b_plan(), order=0)
            sub.upload(ctx2, tiles=False)
            mine = np.ascontiguousarray(shocked[lo:hi])
            pnl = torch.empty(hi - lo, nt, dtype=torch.float64, device=dev)
            ms = max_over_ranks(timed(lambda: ctx2.scenarios(mine, pnl.data_ptr()), reps=3, sync_all=True))
            extras["scenarios_config4"] = {
                "scenarios_total": S, "scenarios_this_rank": hi - lo, "trades": nt, "ranks": world, "ms": ms,
                "revaluations_per_s": S * nt / ms * 1e3, "pnl_bytes_total": S * nt * 8,
                "kernels": ctx2.scenarios_info(),
                "note": "BASELINE config 4: every shocked curve re-bootstrapped on the device (DFs only) + full revaluation (one "
                        "exp per distinct DF query and scenario; unit values by prefix chains: units whose term lists extend "
                        "their predecessor's are summed in one walk); scenarios sharded over the ranks, no collective; includes "
                        "the H2D of the shocked rates; max over ranks"}
            del pnl
            ctx2.close()
        except Exception as ex:  # noqa: BLE001
            extras["scenarios_config4"] = {"error": repr(ex)}
        if world > 1:
            barrier()
        # (3) BASELINE config 5, XCCY half: 500k GBP/USD basis swaps over the ranks, PV + the three ladders per trade
        try:
            from adrates_b200.market_data import SPOT_FX, readme_model
            from adrates_b200.synthetic_xccy import make_xccy_book, XccyBookValuer
            nx = 500_000 // world
            t1 = time.perf_counter()
            xbook = make_xccy_book(readme_model(with_basis=True), nx, seed=11 + rank, spot=SPOT_FX)
            xval = XccyBookValuer(xbook, device=local, stream=stream.cuda_stream)
            x_prep = time.perf_counter() - t1
            ms = max_over_ranks(timed(lambda: xval.value(), reps=10, sync_all=True))
            del xval
            xvg = XccyBookValuer(xbook, device=local, stream=stream.cuda_stream, gamma=True)
            ms_g = max_over_ranks(timed(lambda: xvg.value(), reps=5, sync_all=True))
            extras["xccy_config5"] = {"trades_total": nx * world, "ranks": world, "ms_per_step": ms,
                                      "trades_per_s": nx * world / ms * 1e3, "units": int(xvg.flat_for.n_units),
                                      "ms_per_step_with_gamma": ms_g, "trades_per_s_with_gamma": nx * world / ms_g * 1e3,
                                      "gamma_bytes_per_trade": 4 * 8192,
                                      "flatten_seconds_untimed": x_prep,
                                      "note": "BASELINE config 5, XCCY half: per-trade PV + three 32-wide ladder rows (776 B/trade); "
                                              "with gamma also the three per-curve 32x32 gamma matrices and the foreign x basis "
                                              "cross-gamma matrix of every trade (32 KB/trade); trades sharded over the ranks, no "
                                              "collective in the step; parity: tests/test_gpu_xccy_book.py, tests/test_gpu_xccy_gamma.py"}
            del xvg
        except Exception as ex:  # noqa: BLE001
            extras["xccy_config5"] = {"error": repr(ex)}
        if world > 1:
            barrier()

        # (4) BASELINE config 5, ZCIS half: 500k zero-coupon inflation swaps (two cashflows each on the path-A nodes) sharded over
        # the ranks, device-resident cashflows, per-trade PVs + the shard total; the book total is one all-reduced double
        try:
            nzr = 500_000 // world
            rzr = np.random.Generator(np.random.PCG64(101 + rank))
            tzr = np.repeat(rzr.integers(1, 31, nzr).astype(np.float64) + rzr.uniform(0.0, 0.02, nzr), 2)
            azr = rzr.uniform(-1e6, 1e6, 2 * nzr)
            ozr = np.arange(0, 2 * nzr + 1, 2, dtype=np.int64)
            ctxzr = _native.Context(local)
            ctxzr.set_stream(stream.cuda_stream)
            o_r, t_r, a_r = (torch.from_numpy(x).to(dev) for x in (ozr, tzr, azr))
            pv_r = torch.empty(nzr, dtype=torch.float64, device=dev)
            tot_r = torch.zeros(1, dtype=torch.float64, device=dev)
            ms_z = max_over_ranks(timed(lambda: ctxzr.cashflow_pv_dev(curve._interp_type.value, curve._times, curve._dfs, 0.0, nzr,
                                                                      o_r.data_ptr(), t_r.data_ptr(), a_r.data_ptr(), pv_r.data_ptr(),
                                                                      tot_r.data_ptr()), reps=10, sync_all=True))
            book_total = tot_r.clone()
            if world > 1:
                dist.all_reduce(book_total)
            ref_r = np.array([curve._node_df(float(x)) for x in tzr[:64]])
            err_r = float(np.max(np.abs(pv_r[:32].cpu().numpy() - (azr[:64] * ref_r).reshape(-1, 2).sum(1)) / 1e6))
            extras["zcis_config5"] = {"trades_total": nzr * world, "cashflows_total": 2 * nzr * world, "ranks": world, "ms_per_step": ms_z,
                                      "cashflows_per_s": 2 * nzr * world / ms_z * 1e3, "book_total_pv": float(book_total.item()),
                                      "check_scaled_err_vs_host_df": err_r,
                                      "note": "BASELINE config 5, ZCIS half: leg discounting on the path-A curve (cav_cashflow_pv_dev), "
                                              "cashflows resident in HBM, trades sharded over the ranks, no collective in the step (the "
                                              "book total is one all-reduced double afterwards); max over ranks"}
            ctxzr.close()
        except Exception as ex:  # noqa: BLE001
            extras["zcis_config5"] = {"error": repr(ex)}
        if world > 1:
            barrier()

    if not args.no_extra and world == 1:
        # XccyCurve bootstrap on the device: the GBP/USD basis curve with its first- and second-order spread tangents, and
        # a batch of shocked basis curves re-bootstrapped in one launch (DFs only)
        try:
            from adrates_b200.market_data import readme_model as _rm
            xc = _rm(with_basis=True).curves.GBP_USD_BASIS
            ctxx = _native.Context(local)
            nb_x = len(xc.basis_spreads)
            t1 = time.perf_counter(); xc.device_tables(ctx=ctxx, order=2); xc.device_tables(ctx=ctxx, order=2)
            t2 = time.perf_counter()
            for _ in range(5):
                xc.device_tables(ctx=ctxx, order=2)
            ms2 = (time.perf_counter() - t2) / 5 * 1e3
            S_x = 4096
            shocks = np.array(xc.basis_spreads)[None, :] + np.random.Generator(np.random.PCG64(3)).normal(0, 5e-4, (S_x, nb_x))
            xc.device_tables(ctx=ctxx, spreads=shocks, order=0)
            t3 = time.perf_counter()
            for _ in range(3):
                xc.device_tables(ctx=ctxx, spreads=shocks, order=0)
            ms0 = (time.perf_counter() - t3) / 3 * 1e3
            extras["xccy_curve_device"] = {
                "payment_points": len(xc._pts), "pillars": nb_x, "ms_values_jacobian_hessian": ms2,
                "shocked_curves": S_x, "ms_shocked_curves_dfs": ms0, "curves_per_s": S_x / ms0 * 1e3,
                "note": "cav_xccy_curve_scan (k_xccy_scan): wall time of the host-pointer call incl. plan arrays, copies and the "
                        "read-back; parity: tests/test_gpu_xccy_curve_device.py"}
            ctxx.close()
        except Exception as ex:  # noqa: BLE001
            extras["xccy_curve_device"] = {"error": repr(ex)}
        # ---- single-GPU side measurements (device-resident inputs) ----
        M_PD = _native.REQ_VALUE | _native.REQ_DELTA
        ms = timed(lambda: ctx.portfolio_value(M_PD, pv.data_ptr(), dl.data_ptr(), None, agg.data_ptr()))
        U, T = info["n_units"], info["n_terms"]
        pd_bytes = n * 264 + n * (2 * 8 + 8 + 2 * 4) + U * 264 + T * 28        # rows out, per-trade weights / ids in, unit rows, terms
        extras["pv_delta_config2"] = {
            "ms_per_step": ms, "trades_per_s_per_gpu": n / ms * 1e3, "physical_bytes_model": pd_bytes,
            "hbm_frac_physical": pd_bytes / (ms * 1e-3) / 1e9 / peak,
            "hbm_frac_algorithmic_1520B": n * 1520 / (ms * 1e-3) / 1e9 / peak,
            "note": "BASELINE config 2 (PV + 32-pillar ladder per trade).  The dedup layout reads shared schedule units, so the "
                    "algorithmic figure (every trade charged its own 1 256 B of cashflows) overstates what moves; the physical "
                    "fraction is the honest one: this step is latency-bound, not bandwidth-bound"}
        ms = timed(lambda: ctx.portfolio_value(MASK, None, None, None, agg.data_ptr()))
        extras["portfolio_totals_only"] = {"ms_per_step": ms, "trades_per_s_per_gpu": n / ms * 1e3,
                                           "note": "PV+delta+gamma of the portfolio, no per-trade rows written"}
        # the pre-flattened upload path (round 1's e2e): host flat arrays + tile plan -> pipelined upload -> valuation
        try:
            hflat = abook.flatten(dedup=True, tiles=True) if args.layout == "dedup" else flat
            import copy
            fp = copy.copy(hflat)
            for k in ("unit_offsets", "amt", "weight", "node", "comp_weight", "group_offsets", "group_units", "out_index", "unit_weight"):
                a = getattr(hflat, k)
                setattr(fp, k, None if a is None else pin(a))
            agg_host = np.empty(_native.NOUT)
            ctx.set_async_upload(True)
            for _ in range(2):
                ctx.portfolio_upload(fp)
                ctx.portfolio_value_host(MASK, pv.data_ptr(), dl.data_ptr(), gm.data_ptr(), agg_host)
            torch.cuda.synchronize(dev)
            t1 = time.perf_counter()
            reps = max(5, min(args.steps, 30))
            for _ in range(reps):
                ctx.portfolio_upload(fp)
                ctx.portfolio_value_host(MASK, pv.data_ptr(), dl.data_ptr(), gm.data_ptr(), agg_host)
            dt = (time.perf_counter() - t1) / reps
            ctx.set_async_upload(False)
            extras["e2e_flat_upload"] = {"ms_per_step": dt * 1e3, "trades_per_s": n / dt, "h2d_bytes_per_step": fp.h2d_bytes(),
                                         "note": "cav_portfolio_upload of host-flattened arrays (+ tile plan) from pinned memory + "
                                                 "valuation + totals D2H; the flattening itself (seconds on the host) is NOT in it"}
            if args.layout == "dedup":
                abook.upload(ctx)        # back to the device-built book
                ctx.sync()
        except Exception as ex:  # noqa: BLE001
            extras["e2e_flat_upload"] = {"error": repr(ex)}
        # chain rule as a DMMA GEMM on the private layout (one node-gradient row per trade)
        try:
            ng = min(n, 1_000_000)
            sub_b = type(book)(curve, book.schedules, book.sched[:ng], book.coupon[:ng], book.notional[:ng],
                               book.fixed_sign[:ng], book.spread[:ng])
            priv = flatten_book(sub_b, dedup=False, sort_units=False)
            ctx3 = _native.Context(local)
            ctx3.set_stream(stream.cuda_stream)
            ctx3.curve_build(curve._interp_type.value, curve.swap_rates, curve.path_b_plan(), order=2)
            ctx3.portfolio_upload(priv)
            best = None
            for _ in range(5):
                g_ms, g_fl = ctx3.portfolio_delta_gemm(pv.data_ptr(), dl.data_ptr())
                best = g_ms if best is None else min(best, g_ms)
            tot_ms = timed(lambda: ctx3.portfolio_delta_gemm(pv.data_ptr(), dl.data_ptr()), reps=5)
            extras["chain_gemm_dmma"] = {
                "units": ng, "gemm_ms": best, "gemm_tflops": g_fl / (best * 1e-3) / 1e12,
                "dmma_peak_tflops_measured": 37.13, "tensor_pipe_frac": g_fl / (best * 1e-3) / 1e12 / 37.13,
                "pv_delta_total_ms": tot_ms,
                "note": "delta[U][32] = Q[U][264] x (1e-4 J/d)[264][32], mma.sync.m8n8k4.f64; dense formulation "
                        "(2.6x the flops of the fused sparse chain), reported as an alternative, not the default path"}
            # the irregular-book layout (one private unit per trade) through the full PV + delta + gamma valuation
            if args.layout == "dedup":
                privs = flatten_book(sub_b, dedup=False)
                ctx3.portfolio_upload(privs)
                pagg = torch.zeros(_native.NOUT, dtype=torch.float64, device=dev)

                def pstep():
                    flush.zero_()
                    ctx3.portfolio_value(MASK, pv.data_ptr(), dl.data_ptr(), gm.data_ptr(), pagg.data_ptr())
                fl_ms = timed(lambda: flush.zero_(), reps=10)
                p_ms = timed(pstep, reps=10) - fl_ms
                extras["private_layout"] = {
                    "units": ng, "ms_per_step": p_ms, "trades_per_s": ng / p_ms * 1e3,
                    "roofline": {"bound": "hbm", "frac": ng * BYTES_PER_TRADE / (p_ms * 1e-3) / 1e9 / peak,
                                 "achieved": ng * BYTES_PER_TRADE / (p_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s"},
                    "note": "one private unit per trade (the form irregular trades take): no sharing, every trade's terms are "
                            "read and its Greeks computed by k_units_mma; L2 flushed between steps (flush time subtracted)"}
                del privs
            ctx3.close()
            del priv
        except Exception as ex:  # noqa: BLE001
            extras["chain_gemm_dmma"] = {"error": repr(ex)}
        # ZCIS leg PVs (BASELINE config 5, inflation half): 500k swaps = 1M cashflows on the path-A nodes of the same curve
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
            # the same valuation on device-resident cashflows (cav_cashflow_pv_dev): nothing crosses PCIe
            ctxz.set_stream(stream.cuda_stream)
            o_d, t_d, a_d = (torch.from_numpy(x).to(dev) for x in (oz, tz, az))
            pv_d = torch.empty(nz, dtype=torch.float64, device=dev)
            tot_d = torch.empty(1, dtype=torch.float64, device=dev)
            ms_dev = timed(lambda: ctxz.cashflow_pv_dev(curve._interp_type.value, curve._times, curve._dfs, 0.0, nz, o_d.data_ptr(),
                                                        t_d.data_ptr(), a_d.data_ptr(), pv_d.data_ptr(), tot_d.data_ptr()), reps=10)
            same = bool(np.array_equal(pv_d.cpu().numpy(), pvz))
            extras["zcis_cashflow_pv_config5"] = {"trades": nz, "cashflows": 2 * nz, "ms_e2e_host_buffers": dtz * 1e3,
                                                  "trades_per_s": nz / dtz, "check_scaled_err_vs_host_df": errz,
                                                  "ms_device_resident": ms_dev, "cashflows_per_s_device_resident": 2 * nz / ms_dev * 1e3,
                                                  "device_resident_equals_host_call": same,
                                                  "note": "discounting of the ZCIS legs on the path-A curve; the CPI index "
                                                          "arithmetic that produces the amounts is host logic"}
            ctxz.close()
        except Exception as ex:  # noqa: BLE001
            extras["zcis_cashflow_pv_config5"] = {"error": repr(ex)}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    k = np.array(kern_ms)                      # [steps, 3] units / expand / totals
    dom = int(np.argmax(k.mean(0)))
    dom_name = ["k_units_mma (fused interpolation + PV + delta + gamma per schedule unit, FP64 DMMA tiles)",
                "k_expand_c + k_expand_rows2 (per-trade gamma / PV / delta rows from the compact unit results, streaming 32-byte stores)",
                "k_reduce_partials"][dom]
    dom_ms = float(k[:, dom].mean())
    U, T, K = info["n_units"], info["n_terms"], info["n_comp"]
    if dom == 1:      # what the expansion stage must move: rows out; per-trade weights / row ids (group order) and unit ids /
        # weights (row order) in; unit PV / delta rows in (the compact unit gammas are read from L2, not from DRAM)
        phys = n * 8456 + n * (K * 8 + 8) + n * K * 12 + U * 264
    else:             # the units stage: terms in, unit rows (or, private layout, the trade rows) out
        phys = T * 28 + (n if args.layout == "private" else U) * 8456
    achieved_alg = n * BYTES_PER_TRADE / (dom_ms * 1e-3) / 1e9
    achieved_phys = phys / (dom_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": dom_name, "achieved": achieved_phys, "peak": peak, "unit": "GB/s",
                "frac": achieved_phys / peak, "traffic": None, "peak_source": peak_src,
                "bytes_basis": "physical: bytes the dominant stage must move through HBM per launch (rows written; per-trade weights, row ids "
                               "and unit PV / delta rows read; model from the array sizes, see `traffic` for the ncu capture)",
                "physical_bytes_per_launch": phys,
                "achieved_algorithmic": achieved_alg, "frac_algorithmic": achieved_alg / peak,
                "algorithmic_bytes_per_launch": n * BYTES_PER_TRADE,
                "kernel_ms": dom_ms, "kernel_share_of_step": dom_ms / ms_per_step,
                "step_frac_physical": ((n * 8456 + n * (K * 8 + 8) + n * K * 12 + U * 264 * 2 + T * 28) / (ms_per_step * 1e-3) / 1e9) / peak,
                "step_frac_algorithmic": (n * BYTES_PER_TRADE / (ms_per_step * 1e-3) / 1e9) / peak,
                "all_kernels_ms": {"k_units": float(k[:, 0].mean()), "k_expand": float(k[:, 1].mean()),
                                   "k_reduce_partials": float(k[:, 2].mean())}}
    prof = os.path.join(ROOT, "profiles", "r02h_traffic.json")
    if os.path.exists(prof):
        try:
            with open(prof) as f:
                tr = json.load(f)
            roofline["traffic"] = tr.get(args.layout)
            roofline["traffic_source"] = {"file": "profiles/r02h_traffic.json", "captured_at_commit": tr.get("commit"),
                                          "note": tr.get("note"), "bench_commit": git_commit()}
        except Exception:  # noqa: BLE001
            pass

    cpu_dense, cores, _ = cpu_oracle_rate(cv, curve, book, min(args.cpu_sample, n), dense=True)
    cpu_sparse, _, _ = cpu_oracle_rate(cv, curve, book, min(20 * args.cpu_sample, n), dense=False)
    us = min(200_000, n)
    cpu_units, t_units, t_expand, _ = cpu_units_rate(cv, curve, book, us)
    cpu_baseline = {"value": cpu_dense, "unit": "trades/s", "cores": cores, "kind": "port",
                    "sample": f"first {min(args.cpu_sample, n)} trades of the same book, oracle/liboracle.so with the "
                              "reference's dense chain rule (J^T H J + sum g_k C_k per leg), curve tables built once, "
                              "OpenMP over all host threads",
                    "value_sparse_port": cpu_sparse,
                    "sparse_note": "same arithmetic skipping structurally-zero rows",
                    "value_units_port": cpu_units,
                    "units_note": f"first {us} trades with the SAME unit factorisation as the GPU path (legs of every distinct "
                                  f"schedule valued once: {t_units:.2f} s; per-trade rows expanded as weighted sums of two unit rows: "
                                  f"{t_expand:.2f} s, memory-bound); the fastest CPU formulation we have - the GPU/CPU ratio against "
                                  "this one is the hardware ratio, against `value` it also contains the algorithm"}

    line = {
        "metric": METRIC, "value": value, "unit": "trades/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "trades_per_gpu": n, "layout": args.layout, "flattened_on": where,
                   "units": info["n_units"], "terms": info["n_terms"], "groups": info["n_groups"], "tiles": info["n_tiles"],
                   "l2": "flushed (256 MiB memset) before every timed step; each step also streams "
                         f"{n * 8456 / 1e9:.2f} GB of outputs (>> 126 MB L2)",
                   "timing": "per-step CUDA events on the launch stream, summed over K steps, max over ranks",
                   "collective": ("portfolio totals summed over ranks inside k_reduce_partials_ar (NVLink peer stores + flags), "
                                  "no NCCL call in the step") if all_ranks else "none (single rank)",
                   "parity_gate_scaled_err": gate, "flatten_seconds_untimed": flatten_s, "array_book_seconds_untimed": arrays_s,
                   "wall_seconds_timed_region": wall},
        "clocks": clocks,
        "host_cpus_per_rank": numa,
        "e2e": {"value": e2e_value, "unit": "trades/s", "h2d_bytes_per_step": e2e_h2d,
                "d2h_bytes_per_step": _native.NOUT * 8, "ms_per_step": 1e3 * float(e2e_s.item()) / e2e_steps, "steps": e2e_steps,
                "note": "OISBook.from_arrays(pinned per-trade arrays).compute([VALUE, DELTA, GAMMA]): H2D of effective date (int32 "
                        "serial) / tenor (int32) / side (int8) / coupon / notional (float64), device-side flattening (cav_book_from_arrays: schedules, day counts, "
                        "brackets, units, tile plan), valuation, D2H of the totals; per-trade rows stay in HBM"},
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

#!/usr/bin/env python
"""Headline benchmark: HOTRG_3D on the 3D Ising model (Trivial sector) at chi=24,
seconds per RG step (BASELINE.json `metric`, configs[4]).

    python bench.py --gpus N --steps K --warmup W [--chi 24] [--impl reference]

A "step" is one `step!(::HOTRG_3D, truncrank(chi))` (three z-compressions + leg rotations,
/root/reference/src/schemes/hotrg3d.jl:131-139) followed by `finalize!`, exactly what one
iteration of `run!(scheme, truncrank(chi), maxiter(n))` executes.  The run starts from the
real `classical_ising_3D(Trivial, beta_c)` tensor, so the W warm-up steps are RG iterations
1..W (bond dimensions grow 2 -> 16 -> chi, all legs are chi from iteration 3 on) and the K
timed steps are steady-state chi^6 tensors.  Synthetic data: the model tensor is analytic.

N > 1 (torchrun, one process per GPU, NCCL): the chi^11 contraction of every z-compression is
sharded along the new open x-bond (strong scaling).  T' is replicated either by NVLink peer stores
fused into the slab-producing kernel (torch symmetric memory) or by one NCCL all-gather per
z-compression; the JSON line says which.

Timing: CUDA events on the engine stream, barrier + synchronize on both sides, max over
ranks.  Inputs (1.5 GB at chi=24) exceed the 126 MB L2, so no explicit flush is needed.

Time box.  One steady-state chi=24 step is 9.15e15 FP64 flop, i.e. >= 229 s at the nominal
40 TFLOP/s of the chip, so `--steps 20 --warmup 5` cannot fit any driver limit at N=1.  The run
therefore has a wall budget (`--time-budget`, default 780 s counted from interpreter start;
0 disables it): warm-up stops as soon as >= 2 iterations ran AND every leg equals chi (RG
iteration 2) when one more warm-up plus two timed steps would overrun; timed steps run until the next one would
overrun, never fewer than min(2, K).  The JSON line reports the steps and warm-ups actually
done (`steps`, `warmup`) and what was asked for (`steps_requested`, `warmup_requested`).
(The legs are saturated after RG iteration 2; iteration 3 is already a full-cost step, so on one
GPU the default budget leaves room for 2 warm-up iterations + 2 timed steps.)
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

T_START = time.time()
ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NOMINAL_FP64_TFLOPS = 40.0      # B200 FP64 tensor peak (vendor figure; no measured FP64 entry in
#                                 MEASURED_PEAKS.json and none in the profiling guide)
DMMA_FMA_PER_CLK_SM = 64        # DMMA.8x8x4 pipe: 64 FP64 FMA / clk / SM (148 SMs)

METRIC = "HOTRG_3D Ising chi=%d s/RG-step"
METRIC_ATRG = "ATRG_3D Ising chi=%d s/RG-step"


def _atrg_stats():
    try:
        from tnrkit.jl_b200 import atrg3d_factored as af
        return json.loads(json.dumps(af.LAST_STATS, default=str))
    except Exception:
        return None


def log(*a):
    print("[bench]", *a, file=sys.stderr, flush=True)


# The contract is ONE JSON line on stdout.  Native libraries (NCCL prints "NCCL version ..." from
# C code) also write to file descriptor 1, so the descriptor itself is pointed at stderr for the
# duration of the run and the JSON line is written to the saved original descriptor.
_REAL_STDOUT = None


def capture_stdout():
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(obj):
    line = (json.dumps(obj) + "\n").encode()
    sys.stdout.flush()
    if _REAL_STDOUT is not None:
        os.write(_REAL_STDOUT, line)
    else:
        sys.stdout.write(line.decode())
        sys.stdout.flush()


def step_flops(chi: int) -> float:
    """Algorithmic FP64 flops of one steady-state RG step (all six legs = chi): per
    z-compression 2*chi^11 (the (f,d)-chunked A1*A2 contraction) + 2*2*chi^8 (Q and P)
    + 8*2*chi^8 (projector Gram matrices) + 2*chi^9 + 2*chi^8 (Uy applications)."""
    c = float(chi)
    per = 2 * c ** 11 + 4 * c ** 8 + 16 * c ** 8 + 2 * c ** 9 + 2 * c ** 8
    return 3.0 * per


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-i", str(self.idx), "-lms", "500"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._pump, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None,
                "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------------
# CPU arms (oracle port; the reference itself is Julia and cannot run in this image)
# ----------------------------------------------------------------------------------
_ORACLE_STEP_CACHE = {}


def _host_threads():
    try:
        import threadpoolctl
        return max([p.get("num_threads", 1) for p in threadpoolctl.threadpool_info()] or [1])
    except Exception:
        return os.cpu_count() or 1


def oracle_full_step_seconds(chi_s: int):
    """Times ONE REAL steady-state `step!` + `finalize!` of the oracle's HOTRG_3D (projector
    Gram contractions, eigh, the chi^11 contraction, permutes) at a bond dimension the host can
    hold (chi_s^8 doubles), on the 3D Ising tensor after the iterations that saturate the legs."""
    import numpy as np

    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import tnr_oracle as o

    if chi_s not in _ORACLE_STEP_CACHE:
        sch = o.HOTRG_3D(o.classical_ising_3D())
        sch.finalize()
        while not all(d == chi_s for d in sch.T.shape):
            sch.step(chi_s)
            sch.finalize()
        _ORACLE_STEP_CACHE[chi_s] = sch.T.copy()
    if ("t", chi_s) not in _ORACLE_STEP_CACHE:      # timed once per process
        sch = o.HOTRG_3D(_ORACLE_STEP_CACHE[chi_s])
        t0 = time.perf_counter()
        sch.step(chi_s)
        n = sch.finalize()
        _ORACLE_STEP_CACHE[("t", chi_s)] = time.perf_counter() - t0
        assert np.isfinite(n)
    return _ORACLE_STEP_CACHE[("t", chi_s)]


def cpu_sample(chi: int, target_s: float = 10.0, chi_small: int = 10):
    """Bounded CPU sample of the workload on the host cores (oracle port, numpy + LAPACK/BLAS).

    (1) the oracle's dominant contraction at the FULL size: R[(a y1' y1), cols] = Qk^T Pk[:, cols]
        with K = M = chi^3 -- a column block of ONE of the 3*chi^2 (f,d) chunk contractions of
        an RG step, sized by a calibration run to about `target_s` seconds.  s/RG-step =
        step_flops(chi) / (dgemm flop rate of that sample): a LOWER bound of the CPU time (the
        oracle's SVD / permute / memory-bound phases only add to it), i.e. the comparison
        favours the CPU.
    (2) validation of that flop model: one REAL full oracle step at chi_small (everything
        included), compared with step_flops(chi_small) / the same flop rate."""
    import numpy as np

    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import tnr_oracle as o

    rng = np.random.default_rng(0)
    m = chi ** 3
    cal = 1536
    a = rng.standard_normal((cal, cal)); b = rng.standard_normal((cal, cal))
    o.hotrg3d_chunk_contract(a, b)
    t0 = time.perf_counter(); o.hotrg3d_chunk_contract(a, b); t1 = time.perf_counter()
    gfs = 2.0 * cal ** 3 / (t1 - t0)
    ncols = int(max(8, min(m, target_s * gfs / (2.0 * m * m))))
    Qk = rng.standard_normal((m, m))
    Pk = rng.standard_normal((m, ncols))
    t0 = time.perf_counter()
    R = o.hotrg3d_chunk_contract(Qk, Pk)
    dt = time.perf_counter() - t0
    assert R.shape == (m, ncols)
    del Qk, Pk, R
    flops = 2.0 * m * m * ncols
    rate = flops / dt
    sec_per_step = step_flops(chi) / rate
    cores = _host_threads()
    sample = (f"{ncols} of {m} columns of one of the {3 * chi * chi} (f,d) chunk contractions "
              f"(K=M={m}) per RG step, numpy/BLAS dgemm, {dt:.1f}s, {rate / 1e9:.0f} GF/s; "
              f"scaled by flops to the full step (lower bound: SVD/permute phases not added)")
    validation = None
    if chi_small and chi_small < chi:
        real = oracle_full_step_seconds(chi_small)
        model = step_flops(chi_small) / rate
        validation = {"chi": chi_small, "real_full_oracle_step_s": real,
                      "flop_model_s": model, "real_over_model": real / model}
        sample += (f"; flop model checked against one REAL full oracle step! + finalize! at "
                   f"chi={chi_small}: {real:.2f}s measured vs {model:.2f}s modelled "
                   f"(x{real / model:.2f})")
    return sec_per_step, cores, sample, validation


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    vals = []
    done = 0
    cores = sample = validation = None
    for i in range(args.warmup + args.steps):
        # every step is a bounded sample (about 8 s of dgemm at the full size + one real oracle
        # step at chi=10, timed once); the budget keeps the whole arm within a few minutes
        if args.time_budget > 0 and i >= min(args.warmup + 2, args.warmup + args.steps) and \
                time.time() - T_START > 0.4 * args.time_budget:
            break
        v, cores, sample, validation = cpu_sample(args.chi, target_s=6.0)
        if i >= args.warmup:
            vals.append(v)
            done += 1
        log(f"reference sample {i}: {v:.0f} s/RG-step (extrapolated by flops)")
    value = sum(vals) / len(vals)
    out = {
        "impl": "reference", "metric": METRIC % args.chi, "value": value, "unit": "s/RG-step",
        "n_gpus": args.gpus, "steps": done, "warmup": args.warmup,
        "steps_requested": args.steps,
        "ms_per_step": value * 1e3, "higher_is_better": False, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"HOTRG_3D classical_ising_3D(Trivial) chi={args.chi}, "
                               "steady-state RG step (oracle port of the reference on host cores; "
                               "the Julia reference cannot run in this image)",
                   "flop_model_validation": validation},
        "cpu_baseline": {"value": value, "unit": "s/RG-step", "cores": cores, "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": "s/RG-step", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
    }
    emit(out)


# ----------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------
def measure_fp64_dgemm(torch, dev):
    """cuBLAS DGEMM 8192^3 burst (best of 5): context for the roofline, NOT its denominator."""
    n = 8192
    a = torch.randn(n, n, dtype=torch.float64, device=dev)
    b = torch.randn(n, n, dtype=torch.float64, device=dev)
    torch.matmul(a, b)
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b); e1.record(); e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    del a, b
    torch.cuda.empty_cache()
    return 2.0 * n ** 3 / (best * 1e-3) / 1e12


def run_gpu(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("bench.py --gpus N>1 must be launched with torch.distributed.run")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    import tnrkit.jl_b200 as tk

    ctx = tk.default_context()
    if args.engine == "ozaki":
        ctx.set_option("ozaki", args.ozaki_planes)
    elif args.engine == "ozaki_crt":
        ctx.set_option("ozaki_crt", args.crt_moduli)
    chi = args.chi
    cublas = measure_fp64_dgemm(torch, dev) if rank == 0 else None
    budget = float(args.time_budget)

    def agree(*vals):
        """max over ranks of a few host floats, so every rank takes the same decision."""
        if world == 1:
            return list(vals)
        t = torch.tensor(list(vals), dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.tolist()

    atrg = args.workload == "atrg3d"
    if atrg:
        # BASELINE.json configs[3]: ATRG_3D, factored step (no chi^6 object), chunks of the open
        # bond sharded over the ranks; a dense step is used below chi = 40 unless --factored
        scheme = tk.ATRG_3D(tk.classical_ising_3D(tk.Trivial), shard=world > 1,
                            factored=True if (world > 1 or args.factored or
                                              tk.ATRG_3D.wants_factored(chi)) else False,
                            rfactor=args.rfactor)
    else:
        scheme = tk.HOTRG_3D(tk.classical_ising_3D(tk.Trivial), shard=world > 1)
    trunc = tk.truncrank(chi)
    norms = [scheme.finalize()]

    def state_bufs():
        """Device buffers holding the scheme's tensor (what a host would upload per step)."""
        F = getattr(scheme, "factors", None)
        if F is not None:
            return [F.P.buf[:F.P.size], F.Q.buf[:F.Q.size]]
        return [scheme.T.buf[:scheme.T.size]]

    def state_dims():
        F = getattr(scheme, "factors", None)
        return tuple(F.dims) if F is not None else tuple(scheme.T.dims)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def saturated():
        return all(d == chi for d in state_dims())

    # ---- warm-up: RG iterations 1..W (untimed) -------------------------------------------
    # Bond dimensions grow 2 -> 16 -> chi: every leg equals chi after RG iteration 2, and
    # iteration 3 already costs a full steady-state step (260 s at chi = 24 on one GPU).  When the
    # requested schedule cannot fit the budget, warm-up ends as soon as the legs are saturated
    # (>= 2 iterations: all kernels of the step have run at nearly full size) and one more
    # warm-up plus the minimum of two timed steps would overrun.
    w_done = 0
    last = 0.0
    est_full = 0.0
    min_steps = min(2, args.steps)
    full_flop_rank = (step_flops(chi) / world) if not atrg else None
    while w_done < args.warmup:
        if budget > 0 and w_done >= 2 and saturated():
            elapsed, est_ = agree(time.time() - T_START, est_full)
            if elapsed + (1 + min_steps) * est_ > budget:
                break
        f0 = ctx.counters()["gemm_flops"]
        t0 = time.perf_counter()
        scheme.step(trunc)
        norms.append(scheme.finalize())
        torch.cuda.synchronize()
        last = time.perf_counter() - t0
        flop_it = ctx.counters()["gemm_flops"] - f0
        # a full steady-state step, extrapolated by flops from this (smaller) iteration
        est_full = last * max(1.0, (full_flop_rank / flop_it) if (full_flop_rank and flop_it > 0)
                              else 1.0)
        w_done += 1
        if rank == 0:
            log(f"warm-up step {w_done}/{args.warmup}: dims {state_dims()} norm "
                f"{norms[-1]:.6e} {last:.1f}s, full step ~{est_full:.0f}s (t+{time.time() - T_START:.0f}s)")
    if rank == 0 and not saturated():
        log(f"WARNING: bond dimensions not yet saturated after warm-up: {state_dims()}")

    # ---- timed steps: RG iterations W+1.. ------------------------------------------------
    nelem = sum(b.numel() for b in state_bufs())
    pinned = torch.empty(max(nelem, 2 * chi ** 5 if atrg else chi ** 6), dtype=torch.float64,
                         pin_memory=True)
    sampler = ClockSampler(local)
    ctx.reset_counters()
    ctx.gemm_timing(True)
    ev = []
    barrier()
    if rank == 0:
        sampler.start()
    wall0 = time.perf_counter()
    h2d = d2h = 0
    est = est_full
    k_done = 0
    first_iter = w_done + 1
    for s in range(args.steps):
        if budget > 0 and s >= min_steps:
            elapsed, est_ = agree(time.time() - T_START, est)
            if elapsed + 1.03 * est_ + 25.0 > budget:
                if rank == 0:
                    log(f"time budget: stopping after {s} timed steps (t+{elapsed:.0f}s, next "
                        f"step ~{est_:.0f}s, budget {budget:.0f}s)")
                break
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        # stage the step's input on the host (untimed), then time H2D + step + finalize
        bufs = state_bufs()
        nelem = sum(b.numel() for b in bufs)
        if nelem > pinned.numel():
            pinned = torch.empty(nelem, dtype=torch.float64, pin_memory=True)
        off = 0
        for b in bufs:
            pinned[off:off + b.numel()].copy_(b)
            off += b.numel()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        e[0].record()
        off = 0
        for b in bufs:
            b.copy_(pinned[off:off + b.numel()], non_blocking=True)
            off += b.numel()
        e[1].record()
        scheme.step(trunc)
        norms.append(scheme.finalize())  # device->host read of the step's result (the norm)
        e[2].record()
        torch.cuda.synchronize()
        est = max(est, time.perf_counter() - t0)
        ev.append(e)
        h2d += nelem * 8
        d2h += 8
        k_done += 1
        if rank == 0:
            log(f"timed step {k_done}/{args.steps}: norm {norms[-1]:.12e} "
                f"{e[1].elapsed_time(e[2]) / 1e3:.2f}s (t+{time.time() - T_START:.0f}s)")
    barrier()
    wall = time.perf_counter() - wall0
    clocks = sampler.stop() if rank == 0 else None
    dev_ms = sum(e[1].elapsed_time(e[2]) for e in ev)
    e2e_ms = sum(e[0].elapsed_time(e[2]) for e in ev)
    gemm_ms, gemm_fl, gemm_n = ctx.gemm_timing_read()
    ctr = ctx.counters()
    import ctypes as _C
    _v = _C.c_double()
    ctx.call("tnr_get_counter", b"peer_scatter_launches", _C.byref(_v))
    peer_launches = int(_v.value)
    phases = {}
    for nm in ("hotrg3d.projectors", "hotrg3d.pk_build", "hotrg3d.q_build", "hotrg3d.uy_scatter"):
        ctx.call("tnr_get_counter", ("phase_ms." + nm).encode(), _C.byref(_v))
        phases[nm] = _v.value
    ctx.gemm_timing(False)
    dev_ms, e2e_ms = agree(dev_ms, e2e_ms)
    if rank == 0:
        K = k_done
        sec = dev_ms / 1e3 / K
        e2e_sec = e2e_ms / 1e3 / K
        fl = step_flops(chi) if (saturated() and not atrg) else None
        achieved = gemm_fl / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else None
        if atrg and achieved is None:
            # no single GEMM of the factored step reaches the 1e11-flop timing threshold: report
            # the GEMM flop of the step over the whole step time (see gpu_launches for the rest)
            achieved = ctr["gemm_flops"] / K / sec / 1e12
        sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
        clock_peak = 148 * DMMA_FMA_PER_CLK_SM * 2 * sm_mhz * 1e6 / 1e12
        cpu_v = cpu_cores = cpu_sample_desc = cpu_val = None
        if world == 1 and not args.no_cpu_baseline and not atrg:
            cpu_v, cpu_cores, cpu_sample_desc, cpu_val = cpu_sample(chi)
        out = {
            "metric": (METRIC_ATRG if atrg else METRIC) % chi, "value": sec, "unit": "s/RG-step", "n_gpus": world,
            "steps": K, "warmup": w_done, "steps_requested": args.steps,
            "warmup_requested": args.warmup, "ms_per_step": dev_ms / K,
            "higher_is_better": False, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64" if args.engine == "dmma" else
            (f"f64 emulated by {args.ozaki_planes} int8 digit planes (Ozaki scheme, tcgen05 kind::i8)"
             if args.engine == "ozaki" else
             f"f64 emulated by int8 residue products modulo {args.crt_moduli} coprime moduli "
             f"(CRT / Ozaki scheme II, tcgen05 kind::i8)"),
            "data": "synthetic",
            "config": {
                "engine": args.engine,
                "workload": f"{'ATRG_3D' if atrg else 'HOTRG_3D'} on classical_ising_3D(Trivial, "
                            f"beta_c) truncrank({chi}); timed steps are RG iterations "
                            f"{first_iter}..{first_iter + K - 1} of run! (legs {state_dims()})" +
                            (f"; {'factored' if scheme.factors is not None else 'dense'} step, "
                             f"R factors: {(_atrg_stats() or {}).get('rfactor', args.rfactor)}"
                             if atrg else ""),
                "time_budget_s": budget,
                "time_box": (f"{K} of {args.steps} requested timed steps and {w_done} of "
                             f"{args.warmup} requested warm-up iterations fit the wall budget; "
                             "all legs equal chi from RG iteration 2 on, so every timed step is "
                             "a steady-state step") if (K < args.steps or w_done < args.warmup)
                else "all requested steps ran",
                "parallelism": "1 GPU" if world == 1 else
                (f"{world} GPUs: block columns of the subspace-iteration products and chunks of the "
                 "open bond of AX / YD (H, G) dealt to the ranks and all-gathered (NCCL), the two "
                 "projector pairs built on ranks 0 / 1 and broadcast; orthonormalisations and "
                 "Rayleigh-Ritz steps replicated") if atrg else
                f"open x-bond sharded over {world} GPUs; projector halves dealt to the ranks "
                "and broadcast; exchange: " +
                ("every T' slab stored to all ranks over NVLink by the kernel that produces it "
                 f"({peer_launches} fused scatter launches, symmetric memory), no collective"
                 if peer_launches else "1 NCCL all-gather per z-compression"),
                "l2": "working set (%.2f GB of chi^6 / chunk tensors) exceeds L2; no flush needed" %
                      (chi ** 6 * 8 / 1e9),
                "gemm_flop_per_step": ctr["gemm_flops"] / K,
                "steps_per_s": 1.0 / sec,
                "step_flop": fl,
                "step_tflops": (fl / sec / 1e12) if fl else None,
                "step_frac_of_nominal_fp64_peak": (fl / sec / 1e12 / NOMINAL_FP64_TFLOPS / world)
                if fl else None,
                "norms_tail": norms[-min(3, len(norms)):],
                # size-independent sanity at the full workload: free energy of the run so far
                # (series sum_i log(z_i) 8^(1-i), remainder < 8^-n) against the value the
                # reference tests against (test/schemes.jl:11, f = -3.507, rtol 1e-3)
                "free_energy": tk.free_energy(norms, tk.ising_βc_3D, scalefactor=8.0),
                "free_energy_benchmark": -3.507,
                "atrg3d_stats": _atrg_stats() if atrg else None,
                "wall_s_timed_region": wall,
                "wall_s_since_start": time.time() - T_START,
                "cpu_flop_model_validation": cpu_val,
                # CUDA-event time of the phases of the z-compressions on rank 0, per RG step
                # (the chunk GEMM is `roofline`; projectors are dealt to the ranks when N > 1)
                "phases_s_per_step": {k: v / 1e3 / K for k, v in phases.items()},
                "chunk_gemm_s_per_step": gemm_ms / 1e3 / K,
            },
            "roofline": {
                "bound": "tensor", "achieved": achieved, "peak": NOMINAL_FP64_TFLOPS,
                "unit": "TFLOP/s",
                "frac": (achieved / NOMINAL_FP64_TFLOPS) if achieved else None,
                "peak_source": "of NOMINAL: B200 FP64 tensor peak 40 TFLOP/s (vendor figure); "
                               "MEASURED_PEAKS.json and B200_PROFILING.md hold no FP64 number",
                "frac_of_clock_derived_peak": (achieved / clock_peak) if achieved else None,
                "clock_derived_peak": clock_peak,
                "clock_derived_how": f"148 SM x 64 FP64 FMA/clk (DMMA.8x8x4 pipe) x 2 x "
                                     f"{sm_mhz:.0f} MHz (median SM clock sampled in this run)",
                "frac_of_cublas_dgemm": (achieved / cublas) if achieved and cublas else None,
                "cublas_dgemm_tflops": cublas,
                "cublas_how": "torch.matmul FP64 8192^3, best of 5, measured in this process "
                              "before the run (context only)",
                # dram bytes of one launch are an ncu quantity and are not measured by this
                # run: see profiles/ (ncu --set full raw export of this kernel)
                "traffic": None,
                "kernel": ("all DMMA GEMM launches of the factored ATRG_3D step (chunk "
                           "contractions, Gram matrices, subspace iterations); when no launch reaches "
                           "1e11 flop `achieved` is the GEMM flop of the step over the STEP time "
                           "(GEMMs + orthonormalisations + Rayleigh-Ritz steps + projector SVDs)")
                if atrg else
                          ("gemm_dmma_tma_kernel (TMA + mbarrier producer warp, 8 DMMA consumer "
                           "warps; the (f,d)-chunked chi^3 x chi^3 x chi^3 contraction and the "
                           "projector Gram GEMMs above 1e11 flop)") if args.engine == "dmma" else
                          ("ozaki_tile_kernel (tcgen05.mma kind::i8, TMEM accumulators; achieved = "
                           "FP64-EQUIVALENT flop rate of the emulated chunk contraction, so frac "
                           "against the FP64 peak may exceed 1) + DMMA kernels for the Gram GEMMs"),
                "launches_timed": gemm_n,
                "tma_gemm_launches": ctr.get("tma_gemm_launches"),
            },
            "cpu_baseline": {"value": cpu_v, "unit": "s/RG-step", "cores": cpu_cores,
                             "kind": "port", "sample": cpu_sample_desc},
            "e2e": {"value": e2e_sec, "unit": "s/RG-step", "h2d_bytes_per_step": h2d // K,
                    "d2h_bytes_per_step": d2h // K},
            "gpu_launches": ctr["launches"],
            "clocks": clocks,
        }
        emit(out)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--chi", type=int, default=24)
    ap.add_argument("--time-budget", type=float, default=780.0,
                    help="wall seconds from interpreter start within which the run must end "
                         "(0 = run exactly --steps / --warmup)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="hotrg3d", choices=["hotrg3d", "atrg3d"],
                    help="hotrg3d: the headline (BASELINE.json metric, configs[4], chi=24); atrg3d: "
                         "configs[3] (use --chi 48), same time box and JSON contract")
    ap.add_argument("--rfactor", default=None, choices=["tsqr", "gram", "gram_eigh"],
                    help="atrg3d: R factors of the factored step (default: the scheme's own choice, "
                         "'gram' from chi = 40)")
    ap.add_argument("--factored", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--crt-moduli", type=int, default=16)
    ap.add_argument("--engine", default="dmma", choices=["dmma", "ozaki", "ozaki_crt"],
                    help="dmma: FP64 tensor cores (default, the measured configuration); ozaki: "
                         "EXPERIMENTAL FP64 emulation of the chunk GEMM on the INT8 tensor cores")
    ap.add_argument("--ozaki-planes", type=int, default=8)
    args = ap.parse_args()
    capture_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()

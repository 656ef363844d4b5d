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

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "HOTRG_3D Ising chi=%d s/RG-step"


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
def cpu_sample(chi: int, target_s: float = 12.0):
    """Times the oracle's dominant contraction of the same workload on the host cores.

    Sample: R[(a y1' y1), cols] = Qk^T Pk[:, cols] with K = M = chi^3, i.e. a column block of
    ONE of the 3*chi^2 (f,d) chunk contractions of an RG step, sized by a calibration run to
    about `target_s` seconds; scaled to the full step by flops."""
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
    flops = 2.0 * m * m * ncols
    sec_per_step = dt * step_flops(chi) / flops
    try:
        import threadpoolctl
        cores = max([p.get("num_threads", 1) for p in threadpoolctl.threadpool_info()] or [1])
    except Exception:
        cores = os.cpu_count() or 1
    sample = (f"{ncols} of {m} columns of one of the {3 * chi * chi} (f,d) chunk contractions "
              f"(K=M={m}) per RG step, numpy/BLAS dgemm, {dt:.1f}s, {flops / dt / 1e9:.0f} GF/s; "
              f"scaled by flops to the full step")
    return sec_per_step, cores, sample


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    vals = []
    for i in range(args.warmup + args.steps):
        v, cores, sample = cpu_sample(args.chi, target_s=8.0)
        if i >= args.warmup:
            vals.append(v)
        log(f"reference sample {i}: {v:.0f} s/RG-step (extrapolated)")
    value = sum(vals) / len(vals)
    out = {
        "impl": "reference", "metric": METRIC % args.chi, "value": value, "unit": "s/RG-step",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": value * 1e3, "higher_is_better": False, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"HOTRG_3D classical_ising_3D(Trivial) chi={args.chi}, "
                               "steady-state RG step (oracle port of the reference on host cores; "
                               "the Julia reference cannot run in this image)"},
        "cpu_baseline": {"value": value, "unit": "s/RG-step", "cores": cores, "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": "s/RG-step", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
    }
    emit(out)


# ----------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------
def measure_fp64_peak(torch, dev):
    """cuBLAS DGEMM 8192^3 burst (best of 5) -- the FP64 tensor-core denominator measured in
    this run (MEASURED_PEAKS.json only holds HBM and bf16 numbers)."""
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
    peak = measure_fp64_peak(torch, dev) if rank == 0 else None

    scheme = tk.HOTRG_3D(tk.classical_ising_3D(tk.Trivial), shard=world > 1)
    trunc = tk.truncrank(chi)
    norms = [scheme.finalize()]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for w in range(args.warmup):
        t0 = time.perf_counter()
        scheme.step(trunc)
        norms.append(scheme.finalize())
        torch.cuda.synchronize()
        if rank == 0:
            log(f"warm-up step {w + 1}/{args.warmup}: dims {scheme.T.dims} norm {norms[-1]:.6e} "
                f"{time.perf_counter() - t0:.1f}s")
    if rank == 0 and not all(d == chi for d in scheme.T.dims):
        log(f"WARNING: bond dimensions not yet saturated after warm-up: {scheme.T.dims}")

    nelem = scheme.T.size
    pinned = torch.empty(nelem, dtype=torch.float64, pin_memory=True)
    sampler = ClockSampler(local)
    ctx.reset_counters()
    ctx.gemm_timing(True)
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    barrier()
    if rank == 0:
        sampler.start()
    wall0 = time.perf_counter()
    h2d = d2h = 0
    for s in range(args.steps):
        # stage the step's input on the host (untimed), then time H2D + step + finalize
        pinned[:nelem].copy_(scheme.T.buf[:nelem])
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ev[s][0].record()
        scheme.T.buf[:nelem].copy_(pinned[:nelem], non_blocking=True)
        ev[s][1].record()
        scheme.step(trunc)
        norms.append(scheme.finalize())  # device->host read of the step's result (the norm)
        ev[s][2].record()
        h2d += nelem * 8
        d2h += 8
        nelem = scheme.T.size
        if rank == 0:
            log(f"timed step {s + 1}/{args.steps}: norm {norms[-1]:.12e}")
    barrier()
    wall = time.perf_counter() - wall0
    clocks = sampler.stop() if rank == 0 else None
    dev_ms = sum(e[1].elapsed_time(e[2]) for e in ev)
    e2e_ms = sum(e[0].elapsed_time(e[2]) for e in ev)
    gemm_ms, gemm_fl, gemm_n = ctx.gemm_timing_read()
    ctx.gemm_timing(False)
    ctr = ctx.counters()
    import ctypes as _C
    _v = _C.c_double()
    ctx.call("tnr_get_counter", b"peer_scatter_launches", _C.byref(_v))
    peer_launches = int(_v.value)
    if world > 1:
        t = torch.tensor([dev_ms, e2e_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, e2e_ms = t.tolist()
    if rank == 0:
        K = args.steps
        sec = dev_ms / 1e3 / K
        e2e_sec = e2e_ms / 1e3 / K
        fl = step_flops(chi) if all(d == chi for d in scheme.T.dims) else None
        achieved = gemm_fl / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else None
        cpu_v, cpu_cores, cpu_sample_desc = (None, None, None)
        if world == 1 and not args.no_cpu_baseline:
            cpu_v, cpu_cores, cpu_sample_desc = cpu_sample(chi)
        out = {
            "metric": METRIC % chi, "value": sec, "unit": "s/RG-step", "n_gpus": world,
            "steps": K, "warmup": args.warmup, "ms_per_step": dev_ms / K,
            "higher_is_better": False, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64" if args.engine == "dmma" else
            (f"f64 emulated by {args.ozaki_planes} int8 digit planes (Ozaki scheme, tcgen05 kind::i8)"
             if args.engine == "ozaki" else
             f"f64 emulated by int8 residue products modulo {args.crt_moduli} coprime moduli "
             f"(CRT / Ozaki scheme II, tcgen05 kind::i8)"),
            "data": "synthetic",
            "config": {
                "engine": args.engine,
                "workload": f"HOTRG_3D on classical_ising_3D(Trivial, beta_c) truncrank({chi}); "
                            f"timed steps are RG iterations {args.warmup + 1}.."
                            f"{args.warmup + K} of run! (legs {scheme.T.dims})",
                "parallelism": "1 GPU" if world == 1 else
                f"open x-bond sharded over {world} GPUs; exchange: " +
                ("every T' slab stored to all ranks over NVLink by the kernel that produces it "
                 f"({peer_launches} fused scatter launches, symmetric memory), no collective"
                 if peer_launches else "1 NCCL all-gather per z-compression"),
                "l2": "inputs (chi^6 doubles = %.2f GB) exceed L2; no flush needed" %
                      (scheme.T.size * 8 / 1e9),
                "steps_per_s": 1.0 / sec,
                "step_tflops": (fl / sec / 1e12) if fl else None,
                "step_frac_of_fp64_peak": (fl / sec / 1e12 / peak / world) if fl else None,
                "norms_tail": norms[-min(3, len(norms)):],
                # size-independent sanity at the full workload: free energy of the run so far
                # (series sum_i log(z_i) 8^(1-i), remainder < 8^-n) against the value the
                # reference tests against (test/schemes.jl:11, f = -3.507, rtol 1e-3)
                "free_energy": tk.free_energy(norms, tk.ising_βc_3D, scalefactor=8.0),
                "free_energy_benchmark": -3.507,
                "wall_s_timed_region": wall,
                # not measured in this run: the opt-in INT8 emulation engine, for context
                "experimental_ozaki_engine": None if args.engine == "ozaki" else {
                    "s_per_rg_step_n1_chi24": 146.0,
                    "how": "python bench.py --engine ozaki (separate run, same B200 pool)",
                    "source": "profiles/r01_bench_n1_chi24_ozaki_experimental.json"},
            },
            "roofline": {
                "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                "frac": (achieved / peak) if achieved else None,
                # dram__bytes_read.sum + dram__bytes_write.sum of ONE 13824^3 launch of this
                # kernel (ncu --set full, gpurun_out/prof_gemm_tma_13824_r01.ncu-rep, summarised
                # in profiles/r01_summary.md): 32.13 GB + 1.53 GB for 4.59 GB of operand+result
                # bytes (operands re-read from L2, hit rate 81 %); 227 GB/s, 3 % of HBM
                "traffic": 33.66e9 if (chi == 24 and args.engine == "dmma") else None,
                "traffic_unit": "bytes per launch",
                "kernel": ("gemm_dmma_tma_kernel (TMA + mbarrier producer warp, 8 DMMA consumer "
                           "warps; the (f,d)-chunked chi^3 x chi^3 x chi^3 contraction and the "
                           "projector Gram GEMMs above 1e11 flop)") if args.engine == "dmma" else
                          ("ozaki_tile_kernel (tcgen05.mma kind::i8, TMEM accumulators; achieved = "
                           "FP64-EQUIVALENT flop rate of the emulated chunk contraction, so frac "
                           "against the FP64 peak may exceed 1) + DMMA kernels for the Gram GEMMs"),
                "launches_timed": gemm_n,
                "tma_gemm_launches": ctr.get("tma_gemm_launches"),
                "peak_source": "cuBLAS DGEMM 8192^3 burst measured in this run (of measured; "
                               "MEASURED_PEAKS.json holds no FP64 figure); nominal FP64 tensor "
                               "peak 40 TFLOP/s",
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
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--chi", type=int, default=24)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
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

#!/usr/bin/env python
"""bench.py -- RD forward+adjoint time-steps/sec (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

One bench "step" = one forward (solveState) + adjoint (solveAdjoint) solve of `nt` Strang time
steps over one synthetic atlas-shaped input (config[1] of BASELINE.json: 256^3, single
precision, kappa = 0.01, rho = 8, nt = 25, dt = 0.04, time histories stored).  The reported
`value` is nt * K / T time-steps per second, T timed with CUDA events on the library's stream
between barriers, max over ranks.

  value     inputs resident in HBM when the timed region starts (glia_rd_forward_adjoint)
  e2e       the same through the host-buffer C-ABI call (glia_rd_forward_adjoint_host):
            H2D of c0 and d1 and D2H of c(T) and p(0) inside the timed region
  roofline  dominant kernel: algorithmic bytes per launch / average launch duration (CUDA
            events around every launch of one extra, untimed-for-`value` step) vs the measured
            HBM peak of MEASURED_PEAKS.json; `kernels` lists every kernel family the same way
  cpu_baseline / --impl reference
            the CPU restatement of the reference path (oracle/rd_oracle_torch.py: reference
            operation order, all host threads) on a bounded sample of the same workload

  parity    (inside `config`, so that the driver's record keeps it) the GPU path against the CPU
            restatement on the cpu_baseline sample itself: same k, rho, c0, d1, nt -- relative L2 of
            c(T) and p(0) and the PCG iteration counts of both
  cufft_baseline
            the reference's own CUDA call sequence (cuFFT 3-D R2C/C2R through torch.fft + unfused
            elementwise kernels + host dot read-backs, scripts/cufft_compare.py) on the same GPU in the
            same run -- comparison only, nothing in glia_b200 calls cuFFT

N > 1 (torchrun, one rank per GPU): ONE 512^3 grid cut into N x-slabs (BASELINE config[3], "scaling":
"strong"); the line carries its own 1-GPU point of the same grid (`config.strong_scaling_base`) and the
slab result's relative L2 against that 1-GPU solve (`config.parity`).  `--workload replicas` runs N
independent 256^3 solves instead ("weak").
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

import numpy as np  # noqa: E402

METRIC = "rd_forward_adjoint_time_steps_per_sec"
UNIT = "time-steps/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="auto", choices=["auto", "rd256", "rd512", "slab512", "replicas", "ensemble"],
                    help="auto: 1 GPU -> rd256 (BASELINE config[1]); N GPUs -> slab512 (config[3]: one 512^3 grid "
                         "cut into N x-slabs, strong scaling). rd512 = the same 512^3 solve on one GPU (the strong-"
                         "scaling base). replicas = N independent rd256 solves (weak).")
    ap.add_argument("--n", type=int, default=None)
    ap.add_argument("--nt", type=int, default=None)
    ap.add_argument("--dt", type=float, default=None)
    ap.add_argument("--rho", type=float, default=8.0)
    ap.add_argument("--kappa", type=float, default=0.01)
    ap.add_argument("--precision", default="f32", choices=["f32", "f64"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the config-3 gradient / Hessian / Phi timings")
    ap.add_argument("--members", type=int, default=64, help="ensemble workload: number of members (<= 64)")
    ap.add_argument("--concurrency", type=int, default=1,
                    help="ensemble workload: ensemble handles (host threads) per GPU; a rank's members are split over them")
    ap.add_argument("--no-base", action="store_true", help="slab workload: skip the 1-GPU run of the same grid")
    ap.add_argument("--ref-nt", type=int, default=None,
                    help="time steps per CPU sample (default: 3 for --impl reference, 1 for the in-line cpu_baseline)")
    ap.add_argument("--ref-budget-s", type=float, default=240.0)
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if a.workload == "auto":
        a.workload = "rd256" if max(world, a.gpus) == 1 else "slab512"
    big = a.workload in ("rd512", "slab512")
    if a.n is None:
        a.n = 512 if big else (128 if a.workload == "ensemble" else 256)
    if a.nt is None:
        a.nt = 10 if big else 25      # config 4 uses nt = 10, config 2 nt = 25
    if a.dt is None:
        a.dt = 1.0 / a.nt
    if a.ref_nt is None:
        a.ref_nt = (3 if a.n <= 256 else 1) if a.impl == "reference" else 1
    return a


def workload_config(a, world):
    return {
        "workload": f"{a.n}^3 RD forward+adjoint (solveState + solveAdjoint), synthetic WM/GM/CSF atlas, "
                    f"{'single' if a.precision == 'f32' else 'double'} precision",
        "n": a.n, "nt": a.nt, "dt": a.dt, "rho": a.rho, "kappa": a.kappa, "r_gm": 0.0, "k_gm": 0.0,
        "time_steps_per_bench_step": a.nt,
        "histories": "c_, c_half_, p_ stored (adjoint_store=1)",
        "parallelism": ("single GPU" if world == 1 else
                        (f"one grid cut into {world} x-slabs (one per GPU); x sweeps on peer memory over NVLink"
                         if a.workload == "slab512" else f"{world} independent replicas (one per GPU)")),
        "l2": "working set (time histories, GBs per GPU) exceeds the 126 MB L2; no flush needed",
    }


def make_inputs(a):
    from glia_b200 import synthetic as S
    dtype = np.float32 if a.precision == "f32" else np.float64
    # the 512^3 generator costs about a minute of host FFTs: keep its output for the next run on this box
    cache = os.path.join("/tmp", f"glia_b200_synth_{a.n}_{a.precision}_seed0.npz") if a.n >= 512 else None
    if cache and os.path.exists(cache):
        try:
            z = np.load(cache)
            return {k: z[k] for k in ("wm", "gm", "csf", "vt", "filter")}, z["c0"], dtype
        except Exception:
            pass
    atlas = S.make_atlas(a.n, seed=0, dtype=dtype)
    c0 = S.make_initial_condition(atlas, seed=0, dtype=dtype)
    if cache:
        try:
            np.savez(cache + ".tmp.npz", c0=c0, **atlas)
            os.replace(cache + ".tmp.npz", cache)
        except Exception:
            pass
    return atlas, c0, dtype


# ------------------------------------------------------------------ clocks ----
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._pump, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
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
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)),
                "samples": len(sm), "reasons": sorted(reasons)}


# -------------------------------------------------------- algorithmic bytes ----
# Sweep model of SURVEY.md 8(d) / DESIGN.md: bytes per launch in units of F = one real field.
ALG_F = {
    "kz_deriv2": 3, "ks_deriv2.y": 4, "ks_deriv2.x.matvec": 4, "ks_deriv2.x.rhs": 5, "ks_deriv2.x": 4,
    "kz_r2c": 2, "kz_r2c.axpy": 4, "ks_c2c.y": 2, "ks_pc": 2, "kz_c2r.rz": 3, "kz_c2r": 2, "kz_c2r.norm": 1,
    "k_cg_update": 5, "k_reaction": 4, "k_reaction_lin": 4, "k_axpby": 3,
    "k_pcg_alpha": 0, "k_pcg_beta": 0, "k_pcg_init": 0,
    # slab-decomposed order: x sweep first (peer rows), z adds, y carries the epilogue
    "kx_deriv2_dist": 3, "kz_deriv2.add": 4, "ks_deriv2.y.matvec": 4, "ks_deriv2.y.rhs": 5, "ks_deriv2.y.epi": 4,
    "kx_pc_dist": 2, "k_peer_barrier": 0,
}
# fraction of a kernel's algorithmic bytes that crosses NVLink in slab mode is (G-1)/G of these (in F units)
NVLINK_F = {"kx_deriv2_dist": 2, "kx_pc_dist": 2}
NVLINK_PEAK_GBS = 900.0  # per direction per GPU (NVLink 5), B200_PROFILING.md


def step_bytes_model(F, its_per_solve_total, nsolves, nt):
    """A_min of SURVEY.md 8(d): F * [ sum_solves (33 + 30 m_i) + 10 per time step ]."""
    return F * (33.0 * nsolves + 30.0 * its_per_solve_total + 10.0 * nt)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(tag):
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get(tag)
        except Exception:
            return None
    return None


# ---------------------------------------------------------------- CPU arm ----
def cpu_reference_sample(a, atlas, c0, nt_sample):
    """Build the callable that runs ONE bounded sample (nt_sample forward+adjoint time steps at the
    bench grid) of the restated reference CPU path; returns (fn, cores, description)."""
    import torch
    from oracle import rd_oracle_torch as OT
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    tdt = torch.float32 if a.precision == "f32" else torch.float64
    wm = torch.from_numpy(atlas["wm"]).to(tdt)
    k = (a.kappa * wm).contiguous()
    rho = (a.rho * wm).contiguous()
    kavg = float(k.double().sum() / float(atlas["filter"].astype(np.float64).sum()))
    c0t = torch.from_numpy(c0).to(tdt)
    d1 = (0.9 * c0t).contiguous()

    keep = {}

    def run():
        # dt_ctx = dt/2: the GPU arm's precFactor() runs after an earlier solve has left the context at dt/2
        # (trap T2; in an inversion every gradient evaluation is in that state, DerivativeOperators.cpp:340)
        cT, p0, pde = OT.forward_adjoint(k, kavg, a.kappa, rho, c0t, d1, nt_sample, a.dt, dt_ctx=a.dt / 2)
        keep.update(cT=cT, p0=p0, ks=pde.ksp_state, ka=pde.ksp_adj, d1=d1)
        return pde.ksp_state + pde.ksp_adj

    run.keep = keep
    desc = (f"the first {nt_sample} forward+adjoint time step(s) at {a.n}^3 {a.precision} (d1 = 0.9 c0), restated reference "
            f"CPU path (3-D FFT grad/div, 12 FFTs per PCG iteration, torch CPU FFT + elementwise, {cores} threads); "
            f"no MPI/PETSc/AccFFT on this box")
    return run, cores, desc


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    atlas, c0, _ = make_inputs(a)
    run, cores, desc = cpu_reference_sample(a, atlas, c0, a.ref_nt)
    t_start = time.perf_counter()
    done_w = 0
    # one sample at 512^3 is minutes of host FFTs: no untimed repeats there (the line reports the warm-up it did)
    for _ in range(0 if a.n >= 512 else a.warmup):
        run()
        done_w += 1
        if time.perf_counter() - t_start > a.ref_budget_s / 3:
            break
    times = []
    its = 0
    t0 = time.perf_counter()
    for _ in range(a.steps):
        t1 = time.perf_counter()
        its = run()
        times.append(time.perf_counter() - t1)
        if time.perf_counter() - t_start > a.ref_budget_s:
            break
    T = time.perf_counter() - t0
    k_done = len(times)
    value = a.ref_nt * k_done / T
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": k_done,
        "warmup": done_w, "ms_per_step": 1e3 * T / k_done, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": a.precision, "data": "synthetic",
        "config": dict(workload_config(a, 1), time_steps_per_bench_step=a.ref_nt,
                       note="reference arm: each step is a bounded sample of the workload"),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc,
                         "pcg_iterations_per_sample": its, "pcg_iterations_per_solve": its / (4.0 * a.ref_nt),
                         "note": "the first time steps are the stiffest (more PCG iterations per solve than the mean of "
                                 "the full nt-step run the GPU arm times); compare per-iteration rates with "
                                 "pcg_iterations_per_solve"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64).ravel()
    b = np.asarray(b, dtype=np.float64).ravel()
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def gpu_parity_vs_cpu_sample(a, atlas, c0, keep, device, fsum):
    """The GPU path on exactly the cpu_baseline sample (same k, rho, c0, d1, nt, same solver-context state):
    relative L2 of c(T) and p(0) against the CPU restatement's fields and the PCG iteration counts of both."""
    import torch
    from glia_b200.rd import RDHandle
    dev = torch.device("cuda", device)
    put = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)
    h = RDHandle(a.n, a.precision, device=device, dt_ctx=a.dt / 2)
    wm, gm, csf = put(atlas["wm"]), put(atlas["gm"]), put(atlas["csf"])
    h.set_diffusion_tissue(wm, gm, csf, a.kappa, 0.0, 0.0, fsum)
    h.set_reaction_tissue(wm, gm, csf, a.rho, 0.0, 0.0)
    h.prec_factor()
    h.resize_history(a.ref_nt, a.dt)
    c0d = put(c0)
    d1 = keep["d1"].numpy()
    cT, p0 = torch.empty_like(c0d), torch.empty_like(c0d)
    torch.cuda.synchronize()
    ks, ka = h.forward_adjoint(c0d, put(d1), cT, p0)
    out = {"what": f"GPU path vs the CPU restatement on the cpu_baseline sample ({a.ref_nt} time step(s) at {a.n}^3)",
           "rel_l2_cT": rel_l2(cT.cpu().numpy(), keep["cT"].numpy()),
           "rel_l2_p0": rel_l2(p0.cpu().numpy(), keep["p0"].numpy()),
           "its_gpu": [int(ks), int(ka)], "its_oracle": [int(keep["ks"]), int(keep["ka"])],
           "tolerance": 1e-5 if a.precision == "f32" else 1e-10}
    out["ok"] = bool(out["rel_l2_cT"] < out["tolerance"] and out["rel_l2_p0"] < out["tolerance"]
                     and out["its_gpu"] == out["its_oracle"])
    h.close()
    return out


def cufft_baseline(a, atlas, c0, device, fsum, nt_sample=2):
    """The reference's CUDA call sequence (3-D cuFFT R2C/C2R per gradient / divergence / preconditioner apply,
    unfused elementwise kernels, host dot read-backs: src/grad/SpectralOperators.cpp:100-261,
    src/mat/DiffCoef.cpp:206-271) on the same GPU -- comparison only."""
    import importlib.util
    import torch
    spec = importlib.util.spec_from_file_location("cufft_compare", os.path.join(ROOT, "scripts", "cufft_compare.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    dev = torch.device("cuda", device)
    tdt = torch.float32 if a.precision == "f32" else torch.float64
    wm = torch.from_numpy(atlas["wm"]).to(dev).to(tdt)
    k, rho = a.kappa * wm, a.rho * wm
    kavg = float(k.sum(dtype=torch.float64)) / fsum
    c0d = torch.from_numpy(c0).to(dev).to(tdt)
    d1 = 0.9 * c0d
    path = mod.CufftPath(k, kavg, rho, a.dt, dev)
    path.forward_adjoint(c0d, d1, 1)  # warm-up: cuFFT plans, allocator
    path.nfft = path.its = 0
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    path.forward_adjoint(c0d, d1, nt_sample)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    its, nfft = path.its, path.nfft
    x = torch.randn(a.n, a.n, a.n, device=dev, dtype=tdt)
    xh = torch.fft.rfftn(x)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(10):
        xh = torch.fft.rfftn(x)
        x = torch.fft.irfftn(xh, s=x.shape, norm="forward")
    e1.record()
    torch.cuda.synchronize()
    fft_ms = e0.elapsed_time(e1) / 20
    del path, x, xh
    torch.cuda.empty_cache()
    return {"value": nt_sample / (ms * 1e-3), "unit": UNIT, "kind": "cuFFT call sequence of the reference's CUDA backend "
            "(torch.fft on this GPU), comparison only", "sample": f"the first {nt_sample} forward+adjoint time steps at {a.n}^3",
            "its_per_solve": its / (4.0 * nt_sample), "ffts": nfft, "cufft_ms_per_3d_fft": fft_ms,
            "share_in_cufft": nfft * fft_ms / ms}


# ---------------------------------------------------------------- GPU arm ----
def run_b200(a):
    import torch
    from glia_b200.rd import RDHandle

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != a.gpus and world > 1:
        raise SystemExit(f"--gpus {a.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    atlas, c0, dtype = make_inputs(a)
    F = float(np.dtype(dtype).itemsize) * a.n ** 3
    h = RDHandle(a.n, a.precision, device=local, dt_ctx=a.dt)
    put = lambda x: torch.from_numpy(x).to(dev)
    wm, gm, csf = put(atlas["wm"]), put(atlas["gm"]), put(atlas["csf"])
    fsum = float(atlas["filter"].astype(np.float64).sum())
    c0d = put(c0)
    cT, p0, d1 = torch.empty_like(c0d), torch.empty_like(c0d), torch.empty_like(c0d)
    torch.cuda.synchronize()
    h.resize_history(a.nt, a.dt)
    # data d1: forward solve with (rho, kappa) = (10, 0.025)   -- set-up, untimed
    h.set_diffusion_tissue(wm, gm, csf, 0.025, 0.0, 0.0, fsum)
    h.set_reaction_tissue(wm, gm, csf, 10.0, 0.0, 0.0)
    h.prec_factor()
    h.solve_state(c0d, d1, 0)
    # the benchmarked coefficients
    h.set_diffusion_tissue(wm, gm, csf, a.kappa, 0.0, 0.0, fsum)
    h.set_reaction_tissue(wm, gm, csf, a.rho, 0.0, 0.0)
    h.prec_factor()

    # ---- device-resident timed region -------------------------------------------------------
    for _ in range(a.warmup):
        ks, ka = h.forward_adjoint(c0d, d1, cT, p0)
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    l0 = h.launch_count
    h.timer_start()
    for _ in range(a.steps):
        ks, ka = h.forward_adjoint(c0d, d1, cT, p0)
    ms = h.timer_stop_ms()
    barrier()
    launches = h.launch_count - l0
    clocks = sampler.stop()
    ms = max_over_ranks(ms)
    value = world * a.nt * a.steps / (ms * 1e-3)

    # ---- end-to-end through the host-buffer call ----------------------------------------------
    tdt = torch.float32 if a.precision == "f32" else torch.float64
    hp = [torch.empty((a.n, a.n, a.n), dtype=tdt).pin_memory() for _ in range(4)]
    hp[0].copy_(torch.from_numpy(c0))
    hp[1].copy_(d1.cpu())
    hn = [t.numpy() for t in hp]
    h.forward_adjoint_host(hn[0], hn[1], hn[2], hn[3])
    barrier()
    h.timer_start()
    for _ in range(a.steps):
        h.forward_adjoint_host(hn[0], hn[1], hn[2], hn[3])
    ms_e2e = h.timer_stop_ms()
    barrier()
    ms_e2e = max_over_ranks(ms_e2e)
    e2e_value = world * a.nt * a.steps / (ms_e2e * 1e-3)
    e2e_ok = bool(np.array_equal(hn[2], cT.cpu().numpy()))

    # ---- per-kernel profile of one more step (CUDA events around every launch) ---------------
    h.profile_begin()
    h.forward_adjoint(c0d, d1, cT, p0)
    prof = h.profile_end()
    peak, peak_src = peaks()
    kern = {}
    tot_ms = sum(v[1] for v in prof.values())
    for tag, (cnt, tms) in sorted(prof.items(), key=lambda kv: -kv[1][1]):
        avg = tms / cnt
        fb = ALG_F.get(tag)
        ach = (fb * F / (avg * 1e-3) / 1e9) if fb else None
        kern[tag] = {"launches": cnt, "avg_us": 1e3 * avg, "share": tms / tot_ms,
                     "alg_bytes_per_launch": fb * F if fb is not None else None,
                     "achieved_GBs": ach, "frac": (ach / peak) if ach else None}
    dom = next(iter(kern))
    nsolves = 4 * a.nt
    model_bytes = step_bytes_model(F, ks + ka, nsolves, a.nt)
    step_ach = model_bytes / (ms * 1e-3 / a.steps) / 1e9
    roofline = {
        "bound": "hbm", "kernel": dom, "achieved": kern[dom]["achieved_GBs"], "peak": peak, "unit": "GB/s",
        "frac": kern[dom]["frac"], "traffic": ncu_traffic(dom), "peak_source": peak_src,
        "alg_bytes_per_launch": kern[dom]["alg_bytes_per_launch"], "avg_launch_us": kern[dom]["avg_us"],
        "share_of_step": kern[dom]["share"],
        "whole_step": {"alg_bytes": model_bytes, "achieved": step_ach, "frac": step_ach / peak,
                       "model": "F*[sum_solves(33+30*m_i)+10*nt] (SURVEY 8d A_min)"},
    }

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": a.precision, "data": "synthetic", "config": workload_config(a, world),
        "pcg_iterations": {"state": ks, "adjoint": ka, "solves": nsolves, "mean_per_solve": (ks + ka) / nsolves},
        "notes": {
                  "kernels": "per-kernel times come from a separate event-bracketed pass, which runs the serial order "
                             "(no x||z overlap, no programmatic dependent launch); the timed region uses both"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(2 * F), "d2h_bytes_per_step": int(2 * F),
                "ms_per_step": ms_e2e / a.steps, "matches_device_path": e2e_ok},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
        "kernels": kern,
    }
    if world == 1 and not a.no_extras:
        # one evaluateObjectiveAndGradient, one Gauss-Newton Hessian product with
        # diffusivity inversion, and Phi (np = 8 Gaussians) either side of them -- reported, not the headline
        gfield, yfield = torch.empty_like(c0d), torch.empty_like(c0d)
        h.timer_start()
        og = h.objective_gradient(c0d, d1, wm, gm, csf, beta=1e-4, g_c0=gfield)
        t_grad = h.timer_stop_ms()
        h.set_secondary_tissue(wm, gm, csf, 1.0, 0.0, 0.0)
        h.timer_start()
        _, hits = h.hessian_matvec(c0d, yfield, wm, gm, csf, beta=1e-4, diffusivity_inversion=True)
        t_hess = h.timer_stop_ms()
        ctr = [(math.pi + dx, math.pi + dy, math.pi + dz) for dx in (-0.25, 0.25) for dy in (-0.25, 0.25) for dz in (-0.25, 0.25)]
        h.phi_set(ctr, 2 * math.pi / 64, put(atlas["filter"]), 2 * math.pi / a.n)
        h.timer_start()
        h.phi_apply(yfield, [1.0] * 8)
        t_phi = h.timer_stop_ms()
        h.timer_start()
        h.phi_apply_transpose(gfield)
        t_phit = h.timer_stop_ms()
        line["config3"] = {"objective_gradient_ms": t_grad, "objective_gradient_pcg_its": list(og["its"]),
                           "hessian_matvec_kappa_ms": t_hess, "hessian_matvec_pcg_its": list(hits),
                           "phi_apply_np8_ms": t_phi, "phi_apply_transpose_np8_ms": t_phit}
    h.close()
    mean_its = (ks + ka) / nsolves
    line["config"]["pcg_iterations"] = line["pcg_iterations"]
    line["config"]["whole_step_roofline_frac"] = roofline["whole_step"]["frac"]
    if world == 1 and not a.no_extras:
        try:
            cb = cufft_baseline(a, atlas, c0, local, fsum)
            # per PCG iteration the two do the same numerical work: normalise the sample's stiffer first steps
            cb["value_at_gpu_arm_its_per_solve"] = cb["value"] * cb["its_per_solve"] / mean_its
            cb["speedup_of_this_library"] = value / cb["value_at_gpu_arm_its_per_solve"]
            line["cufft_baseline"] = cb
            line["config"]["cufft_baseline"] = {k: cb[k] for k in ("value", "its_per_solve", "share_in_cufft",
                                                                   "value_at_gpu_arm_its_per_solve", "speedup_of_this_library")}
        except Exception as e:  # comparison only: never fail the bench line over it
            line["cufft_baseline"] = {"unavailable": repr(e)[:200]}
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        run, cores, desc = cpu_reference_sample(a, atlas, c0, a.ref_nt)
        t0 = time.perf_counter()
        its = run()
        tc = time.perf_counter() - t0
        cpu_ips = its / (4.0 * a.ref_nt)
        line["cpu_baseline"] = {"value": a.ref_nt / tc, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": desc, "seconds": tc, "pcg_iterations_per_sample": its,
                                "pcg_iterations_per_solve": cpu_ips,
                                "value_at_gpu_arm_its_per_solve": a.ref_nt / tc * cpu_ips / mean_its}
        par = gpu_parity_vs_cpu_sample(a, atlas, c0, run.keep, local, fsum)
        line["parity"] = par
        line["config"]["parity"] = par
        if not par["ok"]:
            print(json.dumps(line), flush=True)
            raise SystemExit("PARITY FAILURE against the CPU restatement: " + json.dumps(par))
    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


# ------------------------------------------------------- slab-decomposed arm ----
def run_slab(a):
    """One 512^3 (or --n) grid cut into WORLD_SIZE x-slabs, one rank per GPU (strong scaling)."""
    import torch
    import torch.distributed as dist
    from glia_b200.rd import RDHandle

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world < 2:
        raise SystemExit("--workload slab512 needs torchrun with >= 2 ranks (use --workload rd512 on one GPU)")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    gloo = dist.new_group(backend="gloo")  # carries the 64-byte IPC handles (host objects)

    def all_gather(b):
        out = [None] * world
        dist.all_gather_object(out, b, group=gloo)
        return out

    def barrier():
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    dtype = np.float32 if a.precision == "f32" else np.float64
    tdt = torch.float32 if a.precision == "f32" else torch.float64
    lsh = (a.n // world, a.n, a.n)
    F = float(np.dtype(dtype).itemsize) * a.n ** 3          # one GLOBAL field
    Fl = F / world                                          # this rank's share
    # rank 0 generates the global synthetic atlas and scatters the slabs
    meta = [None]
    fields = {}
    if rank == 0:
        atlas, c0, _ = make_inputs(a)
        meta[0] = float(atlas["filter"].astype(np.float64).sum())
        src = {"wm": atlas["wm"], "gm": atlas["gm"], "csf": atlas["csf"], "c0": c0}
    dist.broadcast_object_list(meta, src=0, group=gloo)
    fsum = meta[0]
    for key in ("wm", "gm", "csf", "c0"):
        out = torch.empty(lsh, dtype=tdt, device=dev)
        if rank == 0:
            g = torch.from_numpy(src[key]).to(dev)
            dist.scatter(out, [g[r * lsh[0]:(r + 1) * lsh[0]].contiguous() for r in range(world)], src=0)
            del g
        else:
            dist.scatter(out, None, src=0)
        fields[key] = out
    wm, gm, csf, c0d = fields["wm"], fields["gm"], fields["csf"], fields["c0"]
    cT, p0, d1 = torch.empty_like(c0d), torch.empty_like(c0d), torch.empty_like(c0d)
    torch.cuda.synchronize()

    h = RDHandle(a.n, a.precision, device=local, dt_ctx=a.dt, rank=rank, nranks=world, all_gather=all_gather)
    h.resize_history(a.nt, a.dt)
    h.set_diffusion_tissue(wm, gm, csf, 0.025, 0.0, 0.0, fsum)
    h.set_reaction_tissue(wm, gm, csf, 10.0, 0.0, 0.0)
    h.prec_factor()
    h.solve_state(c0d, d1, 0)
    h.set_diffusion_tissue(wm, gm, csf, a.kappa, 0.0, 0.0, fsum)
    h.set_reaction_tissue(wm, gm, csf, a.rho, 0.0, 0.0)
    h.prec_factor()

    for _ in range(a.warmup):
        ks, ka = h.forward_adjoint(c0d, d1, cT, p0)
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    l0 = h.launch_count
    h.timer_start()
    for _ in range(a.steps):
        ks, ka = h.forward_adjoint(c0d, d1, cT, p0)
    ms = h.timer_stop_ms()
    barrier()
    launches = h.launch_count - l0
    clocks = sampler.stop()
    ms = max_over_ranks(ms)
    value = a.nt * a.steps / (ms * 1e-3)   # ONE job over all ranks: strong scaling

    # ---- end to end: this rank's slabs from / to pinned host memory ---------------------------
    hp = [torch.empty(lsh, dtype=tdt).pin_memory() for _ in range(4)]
    hp[0].copy_(c0d.cpu())
    hp[1].copy_(d1.cpu())
    hn = [t.numpy() for t in hp]
    h.forward_adjoint_host(hn[0], hn[1], hn[2], hn[3])
    barrier()
    h.timer_start()
    for _ in range(a.steps):
        h.forward_adjoint_host(hn[0], hn[1], hn[2], hn[3])
    ms_e2e = h.timer_stop_ms()
    barrier()
    ms_e2e = max_over_ranks(ms_e2e)
    e2e_value = a.nt * a.steps / (ms_e2e * 1e-3)
    e2e_ok = bool(np.array_equal(hn[2], cT.cpu().numpy()))

    # ---- per-kernel profile of one more step (this rank's stream) -----------------------------
    barrier()
    h.profile_begin()
    h.forward_adjoint(c0d, d1, cT, p0)
    prof = h.profile_end()
    barrier()
    peak, peak_src = peaks()
    kern = {}
    tot_ms = sum(v[1] for v in prof.values())
    for tag, (cnt, tms) in sorted(prof.items(), key=lambda kv: -kv[1][1]):
        avg = tms / cnt
        fb = ALG_F.get(tag)
        ach = (fb * Fl / (avg * 1e-3) / 1e9) if fb else None
        k = {"launches": cnt, "avg_us": 1e3 * avg, "share": tms / tot_ms,
             "alg_bytes_per_launch": fb * Fl if fb is not None else None,
             "achieved_GBs": ach, "frac": (ach / peak) if ach else None}
        if tag in NVLINK_F:
            # an x sweep fetches (G-1)/G of its rows from peers and returns as many; every rank does
            # both at once, so EACH direction of this GPU's links carries both amounts
            nvb = NVLINK_F[tag] * Fl * (world - 1) / world
            k["nvlink_bytes_per_direction"] = nvb
            k["nvlink_GBs_per_direction"] = nvb / (avg * 1e-3) / 1e9
            k["nvlink_frac"] = k["nvlink_GBs_per_direction"] / NVLINK_PEAK_GBS
        kern[tag] = k
    dom = next(t for t in kern if kern[t]["achieved_GBs"] is not None)
    nsolves = 4 * a.nt
    model_bytes = step_bytes_model(Fl, ks + ka, nsolves, a.nt)          # per GPU
    step_s = ms * 1e-3 / a.steps
    step_ach = model_bytes / step_s / 1e9
    nv_bytes = (world - 1) / world ** 2 * F * (6.0 * nsolves + 4.0 * (ks + ka))   # per direction per GPU
    roofline = {
        "bound": "hbm", "kernel": dom, "achieved": kern[dom]["achieved_GBs"], "peak": peak, "unit": "GB/s",
        "frac": kern[dom]["frac"], "traffic": ncu_traffic(dom), "peak_source": peak_src,
        "alg_bytes_per_launch": kern[dom]["alg_bytes_per_launch"], "avg_launch_us": kern[dom]["avg_us"],
        "share_of_step": kern[dom]["share"],
        "whole_step": {"alg_bytes_per_gpu": model_bytes, "achieved": step_ach, "frac": step_ach / peak,
                       "model": "per GPU: (F/G)*[sum_solves(33+30*m_i)+10*nt] (SURVEY 8d A_min)"},
        "nvlink": {"bytes_per_direction_per_gpu": nv_bytes, "achieved": nv_bytes / step_s / 1e9,
                   "peak": NVLINK_PEAK_GBS, "unit": "GB/s", "frac": nv_bytes / step_s / 1e9 / NVLINK_PEAK_GBS,
                   "model": "per direction per GPU: (G-1)/G^2 * F * [sum_solves(6 + 4 m_i)] (SURVEY 8e: one exchange "
                            "in and one out per x sweep); averaged over the whole step -- the x sweeps alone are "
                            "kernels[kx_*].nvlink_frac; measured SM-issued peer copy rate on this box: ~740 GB/s "
                            "pull, ~690 GB/s push (scripts/probes/p2p_probe.cu)"},
    }
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": a.precision, "data": "synthetic", "config": workload_config(a, world),
        "pcg_iterations": {"state": ks, "adjoint": ka, "solves": nsolves, "mean_per_solve": (ks + ka) / nsolves},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(2 * F), "d2h_bytes_per_step": int(2 * F),
                "ms_per_step": ms_e2e / a.steps, "matches_device_path": e2e_ok},
        "gpu_launches": int(launches) * world,
        "clocks": clocks,
        "roofline": roofline,
        "kernels": kern,
    }
    h.close()
    barrier()
    line["config"]["pcg_iterations"] = line["pcg_iterations"]
    line["config"]["whole_step_roofline_frac_per_gpu"] = roofline["whole_step"]["frac"]
    line["config"]["nvlink_frac_per_direction"] = roofline["nvlink"]["frac"]
    # the slab result, gathered on rank 0, against the 1-GPU solve of the same grid below
    slabs_cT = [torch.empty_like(cT) for _ in range(world)] if rank == 0 else None
    slabs_p0 = [torch.empty_like(p0) for _ in range(world)] if rank == 0 else None
    if not a.no_base:
        dist.gather(cT, slabs_cT, dst=0)
        dist.gather(p0, slabs_p0, dst=0)
        torch.cuda.synchronize()
    if rank == 0 and not a.no_base:
        # the SAME global grid on ONE GPU (rank 0's, the slab handle released), in the same run: the
        # denominator of the strong-scaling speed-up.  The driver's N = 1 line is the 256^3 headline
        # workload, so this series carries its own 1-GPU point.
        put = lambda x: torch.from_numpy(x).to(dev)
        wm1, gm1, csf1, c01 = put(src["wm"]), put(src["gm"]), put(src["csf"]), put(src["c0"])
        d11, cT1, p01 = torch.empty_like(c01), torch.empty_like(c01), torch.empty_like(c01)
        h1 = RDHandle(a.n, a.precision, device=local, dt_ctx=a.dt)
        h1.resize_history(a.nt, a.dt)
        h1.set_diffusion_tissue(wm1, gm1, csf1, 0.025, 0.0, 0.0, fsum)
        h1.set_reaction_tissue(wm1, gm1, csf1, 10.0, 0.0, 0.0)
        h1.prec_factor()
        h1.solve_state(c01, d11, 0)
        h1.set_diffusion_tissue(wm1, gm1, csf1, a.kappa, 0.0, 0.0, fsum)
        h1.set_reaction_tissue(wm1, gm1, csf1, a.rho, 0.0, 0.0)
        h1.prec_factor()
        h1.forward_adjoint(c01, d11, cT1, p01)
        nb = max(1, min(2, a.steps))
        torch.cuda.synchronize()
        h1.timer_start()
        for _ in range(nb):
            ks1, ka1 = h1.forward_adjoint(c01, d11, cT1, p01)
        ms1 = h1.timer_stop_ms()
        v1 = a.nt * nb / (ms1 * 1e-3)
        line["strong_scaling_base"] = {
            "n_gpus": 1, "value": v1, "unit": UNIT, "steps": nb, "ms_per_step": ms1 / nb,
            "pcg_iterations": {"state": ks1, "adjoint": ka1},
            "same_iterations_as_slab_run": bool((ks1, ka1) == (ks, ka)),
            "speedup": value / v1, "efficiency": value / v1 / world,
            "note": "same global grid and inputs on rank 0's GPU alone, timed after the slab run in this process"}
        h1.close()
        # parity of the slab-decomposed solve: global relative L2 against the 1-GPU solve of the same inputs
        # (the 1-GPU path is the one the oracle parity tests and the N = 1 bench line's `parity` pin)
        def gl2(slabs, full):
            num = sum(float(((torch.cat([sl]).double() - full[r * lsh[0]:(r + 1) * lsh[0]].double()) ** 2).sum())
                      for r, sl in enumerate(slabs))
            return math.sqrt(num) / max(float(full.double().norm()), 1e-300)
        tol = 1e-5 if a.precision == "f32" else 1e-10
        # the adjoint starts from p_T = d1 - c(T): a relative difference of c(T) reaches p_T, and with it p(0),
        # amplified by ||c(T)|| / ||c(T) - d1|| (cancellation) -- the adjoint's own bar is scaled by that factor
        amp = max(1.0, float(cT1.double().norm()) / max(float((cT1.double() - d11.double()).norm()), 1e-300))
        par = {"what": f"{world}-slab solve vs the 1-GPU solve of the same {a.n}^3 inputs, global relative L2",
               "rel_l2_cT": gl2(slabs_cT, cT1), "rel_l2_p0": gl2(slabs_p0, p01),
               "its_slab": [int(ks), int(ka)], "its_1gpu": [int(ks1), int(ka1)], "tolerance": tol,
               "terminal_condition_amplification": amp, "tolerance_p0": tol * amp}
        # iteration totals: the ranks add their partial dot products in rank order, one GPU adds them in its own order;
        # in single precision a convergence test that sits on the threshold can flip by one iteration between the two
        # (seen: 221 against 220 adjoint iterations over 20 solves at 8 GPUs).  The oracle tests demand equal counts at
        # the sizes the oracle runs; here the bar is the fields plus "no more than one iteration per 20 solves apart".
        its_gap = max(abs(int(ks) - int(ks1)), abs(int(ka) - int(ka1)))
        par["its_gap"] = its_gap
        par["its_gap_allowed"] = max(1, nsolves // 40)
        par["ok"] = bool(par["rel_l2_cT"] < tol and par["rel_l2_p0"] < tol * amp and its_gap <= par["its_gap_allowed"])
        line["parity"] = par
        line["config"]["parity"] = par
        line["config"]["strong_scaling_base"] = {k: line["strong_scaling_base"][k] for k in
                                                 ("n_gpus", "value", "speedup", "efficiency", "same_iterations_as_slab_run")}
    barrier()
    if rank == 0:
        print(json.dumps(line), flush=True)
        if "parity" in line and not line["parity"]["ok"]:
            dist.destroy_process_group()
            raise SystemExit("PARITY FAILURE of the slab-decomposed solve: " + json.dumps(line["parity"]))
    dist.destroy_process_group()


def run_ensemble(a):
    """Config 5 of BASELINE.json: 64 independent 128^3 forward solves, (kappa, rho) on an 8 x 8 grid
    (kappa in [0.005, 0.05] log-spaced, rho in [4, 15]), spread over the ranks (8 per GPU at
    N = 8).  A rank's members ride in ONE ensemble handle (glia_rd_create_batch): the kernels carry a
    member index, every member keeps its own coefficients, PCG state and iteration count, one set of
    launches advances all of them.  `--concurrency` > 1 splits a rank's members over that many handles
    driven from host threads (the round-1 arrangement, kept for comparison).  Replicas only: no collective
    on the data path; the value is aggregate member-time-steps per second."""
    import threading
    import torch
    from glia_b200.rd import RDHandle

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    atlas, c0, dtype = make_inputs(a)
    put = lambda x: torch.from_numpy(x).to(dev)
    wm, gm, csf, c0d = put(atlas["wm"]), put(atlas["gm"]), put(atlas["csf"]), put(c0)
    fsum = float(atlas["filter"].astype(np.float64).sum())
    kappas = np.exp(np.linspace(np.log(0.005), np.log(0.05), 8))
    rhos = np.linspace(4.0, 15.0, 8)
    members = [(float(k), float(r)) for k in kappas for r in rhos][: a.members]
    mine = members[rank::world]
    conc = max(1, min(a.concurrency, len(mine)))
    groups = [mine[j::conc] for j in range(conc)]          # one ensemble handle per group
    handles = [RDHandle(a.n, a.precision, device=local, dt_ctx=a.dt, nbatch=len(g)) for g in groups]
    c0b = [c0d.unsqueeze(0).repeat(len(g), 1, 1, 1).contiguous() for g in groups]
    outs = [torch.empty_like(x) for x in c0b]
    for h in handles:
        h.resize_history(a.nt, a.dt)
    torch.cuda.synchronize()
    its_total = [0] * conc
    ms_thread = [0.0] * conc

    def worker(j, timed):
        torch.cuda.set_device(local)
        h, g = handles[j], groups[j]
        if timed:
            h.timer_start()
        # coefficient set-up is part of a member's cost (it differs per member), as in round 1
        h.set_coefficients_batch(wm, gm, csf, [m[0] for m in g], 0.0, 0.0, fsum, [m[1] for m in g], 0.0, 0.0)
        h.prec_factor()
        its = h.solve_state(c0b[j], outs[j], 0)
        if timed:
            ms_thread[j] = h.timer_stop_ms()
        its_total[j] = its

    def sweep(timed):
        if conc == 1:
            worker(0, timed)
            return
        th = [threading.Thread(target=worker, args=(j, timed)) for j in range(conc)]
        [t.start() for t in th]
        [t.join() for t in th]

    for _ in range(a.warmup):
        sweep(False)
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    l0 = sum(h.launch_count for h in handles)
    ms = 0.0
    for _ in range(a.steps):
        sweep(True)
        ms += max(ms_thread)     # the streams start together; the slowest one ends the step
    barrier()
    launches = sum(h.launch_count for h in handles) - l0
    clocks = sampler.stop()
    ms = max_over_ranks(ms)
    value = len(members) * a.nt * a.steps / (ms * 1e-3)
    nsolves = 2 * a.nt * len(mine)
    F = float(np.dtype(dtype).itemsize) * a.n ** 3
    peak, peak_src = peaks()
    model_bytes = step_bytes_model(F, sum(its_total), nsolves, a.nt * len(mine)) - 7.0 * F * a.nt * len(mine)
    per_member_its = [h.batch_iterations(True) for h in handles]
    line = {
        "metric": "rd_ensemble_member_time_steps_per_sec", "value": value, "unit": "member-time-steps/s",
        "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms / a.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": a.precision, "data": "synthetic",
        "config": {"workload": f"config 5: {len(members)} independent {a.n}^3 RD forward solves, (kappa, rho) on an "
                               "8 x 8 grid, replicas only (no data-path collective)", "n": a.n, "nt": a.nt,
                   "dt": a.dt, "members": len(members), "members_per_gpu": len(mine),
                   "ensemble_handles_per_gpu": conc, "members_per_handle": [len(g) for g in groups],
                   "batching": "members ride in the kernels' member index (glia_rd_create_batch): one set of launches "
                               "per PCG iteration for all members of a handle",
                   "value_per_gpu": value / world,
                   "l2": "a 128^3 field is 8 MB; a handle's members together exceed the 126 MB L2 from 3 members on "
                         "(k, rho, 5 PCG vectors, histories per member)"},
        "pcg_iterations": {"state": int(sum(its_total)), "solves": nsolves,
                           "mean_per_solve": sum(its_total) / max(nsolves, 1),
                           "per_member_min_max": [int(min(min(x) for x in per_member_its)),
                                                  int(max(max(x) for x in per_member_its))]},
        "gpu_launches": int(launches) * world, "clocks": clocks,
        "roofline": {"bound": "hbm", "whole_step": {"alg_bytes_per_gpu": model_bytes,
                                                     "achieved": model_bytes * a.steps / (ms * 1e-3) / 1e9,
                                                     "frac": model_bytes * a.steps / (ms * 1e-3) / 1e9 / peak},
                     "peak": peak, "unit": "GB/s", "peak_source": peak_src,
                     "note": "forward only: A_min model F*[sum_solves(33+30 m_i) + 3 nt] per member, with every "
                             "member's own iteration counts"},
    }
    if rank == 0:
        print(json.dumps(line), flush=True)
    for h in handles:
        h.close()
    if dist is not None:
        dist.destroy_process_group()


def main():
    # NCCL prints its version banner (NCCL_DEBUG=VERSION on the bench boxes) to stdout; the contract is ONE
    # JSON line there, so send NCCL's own log to stderr
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    elif a.workload == "slab512":
        run_slab(a)
    elif a.workload == "ensemble":
        run_ensemble(a)
    else:
        run_b200(a)


if __name__ == "__main__":
    main()

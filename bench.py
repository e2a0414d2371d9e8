#!/usr/bin/env python
"""bench.py -- full PT P(k, mu) model evaluations per second (BASELINE.json metric).

One "evaluation" = rebuild of EVERY perturbation-theory table for a new linear P(k) (16 I_mn, 6 J_mn,
13 K_mn, the four one-loop spectra on 1000 k, the 7 x 3 ImnOneLoop integrals, sigma_v^2(k), X(r), Y(r),
Zel'dovich P00/P01/P11) plus one GalaxySpectrum.power on the BASELINE C2 grid (1000 k x 40 mu).
One "step" = a batch of `--batch` such evaluations per GPU (distinct synthetic cosmologies).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl reference]

N > 1 is launched by torchrun (one rank per GPU); cosmologies are sharded over ranks with no data-path
collective, the per-walker result (mu-averaged P(k), 1000 doubles) is all-gathered over NCCL inside the
timed region.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "full PT P(k,mu) model evals/sec"
UNIT = "evals/s"
from pyrsd_b200.synthetic import NK_OUT, NMU_OUT, SEED, synthetic_batch  # noqa: E402  (workload generator, no GPU needed)


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons DURING the timed region (B200_PROFILING.md): NVML polled every few ms (the timed
    region of a default run lasts tens of ms), `nvidia-smi` as a fallback"""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        threading.Thread.__init__(self, daemon=True)
        self.index, self.samples, self.stop_flag, self.armed = index, [], False, False
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def run(self):
        while not self.stop_flag:
            if not self.armed:
                time.sleep(0.001)
                continue
            try:
                if self.nvml is not None:
                    n = self.nvml
                    mhz = float(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM))
                    r = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.h)) if hasattr(n, 'nvmlDeviceGetCurrentClocksEventReasons') \
                        else int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                    flags = [bool(r & 0x8), bool(r & 0x40), bool(r & 0x20), bool(r & 0x4)]   # hw_slowdown, hw_thermal, sw_thermal, sw_power_cap
                    self.samples.append([mhz, self.max_mhz] + flags)
                    time.sleep(0.002)
                else:
                    out = subprocess.run(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q, '--format=csv,noheader,nounits'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, timeout=5).stdout.strip()
                    if out:
                        f = [x.strip() for x in out.split(',')]
                        self.samples.append([float(f[0]), float(f[1])] + [x.lower().startswith('active') for x in f[2:6]])
                    time.sleep(0.05)
            except Exception:
                time.sleep(0.05)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no clock samples (NVML and nvidia-smi unavailable)"]}
        sm = sorted(s[0] for s in self.samples)
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [names[j] for j in range(4) if any(s[2 + j] for s in self.samples)]
        return {"sm_mhz": sm[len(sm)//2], "sm_max_mhz": self.samples[0][1], "reasons": reasons, "samples": len(sm),
                "source": "nvml" if self.nvml is not None else "nvidia-smi"}


# ------------------------------------------------------------------------------- GPU arm
def run_gpu(args):
    import ctypes as C
    import torch
    import torch.distributed as dist
    from pyrsd_b200 import _native as N, parallel, pygcl, rsd

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    L = rsd._L()
    vp, dp = C.c_void_p, N._dp
    L.prsd_ctx_profile.argtypes = [vp, C.c_int]
    L.prsd_ctx_profile_read.argtypes = [vp, dp, np.ctypeslib.ndpointer(dtype=np.int64, flags='C_CONTIGUOUS')]   # 8 families
    L.prsd_fp64_peak.argtypes = [vp, C.c_int, C.POINTER(C.c_double)]
    L.prsd_spectra_reload_device.argtypes = [vp, vp, vp, vp]
    L.prsd_ctx_stream.argtypes = [vp, C.POINTER(vp)]

    ctx = N.default_context(local)
    sptr = vp()
    N.check(L.prsd_ctx_stream(ctx.ptr, C.byref(sptr)))
    stream = torch.cuda.ExternalStream(sptr.value, device=torch.device('cuda', local))

    B = args.batch
    data = synthetic_batch(B, rank)
    eng = rsd.PTEngine(context=ctx)
    info = eng.info()
    gal = eng.galaxy(data['kmodel'], rsd.galaxy_config())
    npts = len(data['kk'])

    # pinned host staging (e2e) and device-resident copies (value)
    def pinned(a):
        t = torch.empty(a.shape, dtype=torch.float64, pin_memory=True)
        t.numpy()[...] = a
        return t
    h_k, h_T, h_pars = pinned(data['k']), pinned(data['T']), pinned(data['pars'])
    h_wpar, h_stoch = pinned(data['wpar']), pinned(data['stoch'])
    h_kk, h_mm = pinned(data['kk']), pinned(data['mm'])
    h_out = torch.empty((B, npts), dtype=torch.float64, pin_memory=True)
    dev = torch.device('cuda', local)
    d_k, d_T, d_pars = h_k.to(dev), h_T.to(dev), h_pars.to(dev)
    d_kk, d_mm = h_kk.to(dev), h_mm.to(dev)
    d_out = torch.empty((B, npts), dtype=torch.float64, device=dev)
    # L2 flush buffer (> 126 MB): written between timed iterations
    flush = torch.empty(64*1024*1024, dtype=torch.float32, device=dev)
    torch.cuda.synchronize()

    sp = eng.spectra(data['k'], data['T'], data['pars'][:, 0], data['pars'][:, 1], data['pars'][:, 2], data['pars'][:, 3],
                     data['pars'][:, 4], data['pars'][:, 5])
    res = rsd._new_result()

    def step_device():
        """inputs resident in HBM, outputs left in HBM"""
        N.check(L.prsd_spectra_reload_device(sp.ptr, d_k.data_ptr(), d_T.data_ptr(), d_pars.data_ptr()))
        eng.run(sp, result=res)
        N.check(L.prsd_galaxy_power_device(gal.ptr, sp.ptr, res.ptr, B, None, h_wpar.numpy(), h_stoch.numpy(), npts,
                                           d_kk.data_ptr(), d_mm.data_ptr(), d_out.data_ptr()))
        if world > 1:
            # the ensemble step's exchange: every rank receives the per-walker result of all walkers
            torch.cuda.current_stream().wait_stream(stream)
            parallel.gather_rows(d_out.view(B, NK_OUT, NMU_OUT).mean(-1), world*B)
            stream.wait_stream(torch.cuda.current_stream())

    def step_e2e():
        """the public call path with HOST buffers: H2D of the step's inputs and D2H of P(k, mu) inside"""
        sp2 = eng.spectra(h_k.numpy(), h_T.numpy(), h_pars.numpy()[:, 0], h_pars.numpy()[:, 1], h_pars.numpy()[:, 2],
                          h_pars.numpy()[:, 3], h_pars.numpy()[:, 4], h_pars.numpy()[:, 5])
        eng.run(sp2, result=res)
        N.check(L.prsd_galaxy_power(gal.ptr, sp2.ptr, res.ptr, B, None, h_wpar.numpy(), h_stoch.numpy(), npts,
                                    h_kk.numpy(), h_mm.numpy(), h_out.numpy()))
        sp2.close()
        return h_out

    sampler = ClockSampler(local)

    def timed(fn, steps, warmup, profile=False):
        for _ in range(warmup):
            fn()
            flush.zero_()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        if profile:
            N.check(L.prsd_ctx_profile(ctx.ptr, 1))
        sampler.armed = True
        l0 = ctx.launches()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for e0, e1 in evs:
            flush.zero_()                       # L2 flush between timed iterations (outside the event pair)
            torch.cuda.synchronize()
            e0.record(stream)
            fn()
            e1.record(stream)
        torch.cuda.synchronize()
        sampler.armed = False
        if world > 1:
            dist.barrier()
        ms = sum(e0.elapsed_time(e1) for e0, e1 in evs)
        launches = ctx.launches() - l0
        prof = None
        if profile:
            pms, pc = np.zeros(8), np.zeros(8, dtype=np.int64)
            N.check(L.prsd_ctx_profile_read(ctx.ptr, pms, pc))
            N.check(L.prsd_ctx_profile(ctx.ptr, 0))
            prof = (pms, pc)
        ms = parallel.max_over_ranks(ms, device=dev)
        return ms, launches, prof

    # FP64 pipe peaks, measured once outside the timed region (MEASURED_PEAKS.json has no FP64 entry)
    pk = C.c_double()
    N.check(L.prsd_fp64_peak(ctx.ptr, 0, C.byref(pk)))
    peak_dfma = pk.value
    N.check(L.prsd_fp64_peak(ctx.ptr, 1, C.byref(pk)))
    peak_dmma = pk.value

    if rank == 0:
        sampler.start()
    ms_dev, launches, prof = timed(step_device, args.steps, args.warmup, profile=True)
    ms_e2e, _, _ = timed(step_e2e, max(2, args.steps//2), 1)
    sampler.stop_flag = True
    e2e_steps = max(2, args.steps//2)

    # quick in-run parity guard: the first cosmology of the batch equals an independent one-off evaluation
    out0 = d_out[0].cpu().numpy()
    assert np.all(np.isfinite(out0))       # (the PT model itself may go negative at k mu ~ 0.5 for extreme synthetic walkers)
    assert np.allclose(step_e2e()[0].numpy(), out0, rtol=1e-12)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    # ---- roofline of the mode-coupling operators: algorithmic FP64 flops = declared nodes x
    #      cosmologies x flops per (node, cosmology) of the implemented rule, FMA = 2
    pms, pc = prof
    # per (node, cosmology): 4-point table read of every spectrum (7 flop each) + products/accumulation + the Lagrange
    # coefficients formed from the stencil coordinate (13 flop per node, shared by the 4 cosmologies of a one-loop CTA)
    FLOP_ML, FLOP_DIAG, FLOP_IK = 109, 20.25, 67
    kern = {
        'k_nodal_apply<SpecML>': dict(fam=6, flops=info['nodes_ml_half']*B*FLOP_ML, pipe='fp64 DFMA'),
        'k_nodal_apply<SpecDiag>': dict(fam=7, flops=info['nodes_oneloop']*B*FLOP_DIAG, pipe='fp64 DFMA'),
        'k_bilinear_apply<4>': dict(fam=0, flops=info['nodes_ik']*8*((B + 7)//8)*FLOP_IK, pipe='fp64 DMMA m8n8k4'),
    }
    traffic, limiter = {}, {}
    try:
        prof_json = json.load(open(os.path.join(ROOT, 'profiles', 'r01_ncu_summary.json')))
        traffic = prof_json.get('dram_bytes_per_launch', {})
        for kname, kd in prof_json.get('kernels', {}).items():
            limiter[kname] = {m.split('.')[0]: float(kd[m]['value']) for m in (
                'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
                'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
                'sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active',
                'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed') if m in kd}
    except Exception:
        pass
    for name, kd in kern.items():
        n = max(int(pc[kd['fam']]), 1)
        kd['ms'] = pms[kd['fam']]/n
        kd['launches_per_step'] = n//args.steps
        kd['flops'] = float(kd['flops'])
        kd['achieved'] = kd['flops']/(kd['ms']*1e-3)/1e12 if kd['ms'] > 0 else None
        kd['share'] = pms[kd['fam']]/ms_dev
    top = max(kern, key=lambda k: kern[k]['share'])
    peak = peak_dmma if 'DMMA' in kern[top]['pipe'] else peak_dfma
    fam = ['ik_dmma', 'linear_ops', 'zeldovich', 'galaxy', 'spline_builds', 'fftlog', 'imn_oneloop_nodal', 'oneloop_tables_nodal']
    shares = {fam[i]: round(pms[i]/ms_dev, 4) for i in range(8)}
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    value = world*B*args.steps/(ms_dev*1e-3)
    e2e_value = world*B*e2e_steps/(ms_e2e*1e-3)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_dev/args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": dict(workload_config(B), quadrature_nodes_per_cosmology=info['nodes'],
                       l2="flushed between timed iterations (256 MB write)",
                       parallelism="walkers sharded over %d rank(s), NCCL all-gather of the mu-averaged P(k)" % world),
        "e2e": {"value": e2e_value, "unit": UNIT,
                "h2d_bytes_per_step": int(h_k.numel()*8 + h_T.numel()*8 + h_pars.numel()*8 + h_wpar.numel()*8 + h_stoch.numel()*8
                                          + h_kk.numel()*8 + h_mm.numel()*8),
                "d2h_bytes_per_step": int(h_out.numel()*8), "ms_per_step": ms_e2e/e2e_steps},
        "gpu_launches": int(launches),
        "roofline": {"bound": "tensor", "achieved": kern[top]['achieved'], "peak": peak, "unit": "TFLOP/s",
                     "frac": (kern[top]['achieved']/peak) if (kern[top]['achieved'] and peak) else None,
                     "traffic": traffic.get(top),
                     "ncu_pct_of_peak": limiter.get(top),     # from profiles/r01_ncu_summary.json: the binding unit is the LSU data pipe
                     "kernel": top, "pipe": kern[top]['pipe'] + " (B200 runs DFMA and FP64 DMMA on the same pipe: measured peaks "
                                                                "%.1f / %.1f TFLOP/s)" % (peak_dfma, peak_dmma),
                     "peak_source": "FP64 microbenchmark (csrc/microbench.cu) run in this process before the timed region; "
                                    "MEASURED_PEAKS.json has no FP64 entry (hbm_gbs there = %s)" % peaks.get('hbm_gbs'),
                     "algorithmic_flops_per_launch": kern[top]['flops'], "kernel_ms_per_launch": kern[top]['ms'],
                     "launches_per_step": kern[top]['launches_per_step'],
                     "flops_per_node_per_cosmology": {"k_nodal_apply<SpecML>": FLOP_ML, "k_nodal_apply<SpecDiag>": FLOP_DIAG,
                                                      "k_bilinear_apply<4>": FLOP_IK},
                     "nodes": {"imn_oneloop": info['nodes_ml_half'], "oneloop_tables": info['nodes_oneloop'], "ik": info['nodes_ik']},
                     "other_kernels": {k: {"achieved": v['achieved'], "frac": (v['achieved']/(peak_dmma if 'DMMA' in v['pipe'] else peak_dfma))
                                           if v['achieved'] else None, "ms_per_launch": v['ms'], "traffic": traffic.get(k)}
                                       for k, v in kern.items() if k != top},
                     "kernel_share_of_step": shares},
        "clocks": sampler.summary(),
    }
    if not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline_sample()
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------ CPU reference (oracle)
def _cpu_eval(seed, thin=1):
    """ONE evaluation of the hot path on the reference's own C++ (oracle/_ref) at the reference's default tolerances:
    every PT table of rsd/pt_integrals.py (16 I, 6 J, 13 K, sigma_v^2(k) on the 250-point k_interp; the four
    one-loop spectra on 1000 k; the 7 x 3 ImnOneLoop integrals), X(r), Y(r) on 1024 r, and the Zel'dovich P00/P01/P11
    on the 200-point model grid at both normalisations.  `thin` > 1 evaluates every thin-th point of each grid and
    scales the time (a bounded sample); returns the seconds of the full evaluation."""
    from oracle import gcl_ref as R
    g = np.load(os.path.join(ROOT, 'tests', 'golden', 'golden_pt.npz'))
    cosmo = R.golden_cosmology(g)
    P = R.RefLinearPS(cosmo)
    off = seed % thin
    kint = np.ascontiguousarray(np.logspace(np.log10(5e-6), 0, 250)[off::thin])
    k1 = np.ascontiguousarray(R.logspace(1e-5, 10, 1000)[off::thin])
    t0 = time.perf_counter()
    for m in range(4):
        for n in range(4):
            R.Imn(P, kint, m, n, epsrel=1e-4)
    for (m, n) in [(0, 0), (0, 1), (0, 2), (1, 0), (1, 1), (2, 0)]:
        R.Jmn(P, kint, m, n, epsrel=1e-4)
    for (m, n, t, p) in [(0, 0, 0, 0), (0, 1, 0, 0), (0, 0, 1, 0), (0, 1, 1, 0), (0, 2, 1, 0), (1, 0, 0, 0), (1, 1, 0, 0),
                         (1, 0, 1, 0), (1, 1, 1, 0), (2, 0, 0, 0), (2, 0, 0, 1), (2, 0, 1, 0), (2, 0, 1, 1)]:
        R.Kmn(P, kint, m, n, bool(t), p, epsrel=1e-3)
    P.VelocityDispersion(kint, 0.5)
    for (m, n) in [(0, 0), (0, 1), (1, 1), (2, 3), (3, 2), (3, 3)]:
        R.Imn(P, k1, m, n, epsrel=1e-4)
    for (m, n) in [(0, 0), (0, 1), (1, 1)]:
        R.Jmn(P, k1, m, n, epsrel=1e-4)
    # ImnOneLoop on the shipped one-loop tables
    ol = {w: R.RefOneLoop(P, w, 1e-4, g['oneloop_P%s_0' % w]) for w in ('dd', 'dv', 'vv')}
    for (p1, p2, m, n) in [('vv', 'dd', 0, 1), ('vv', 'dd', 0, 2), ('dv', None, 0, 3), ('dv', None, 0, 4),
                           ('vv', None, 2, 3), ('vv', None, 3, 2), ('vv', None, 3, 3)]:
        for which in ('linear', 'cross', 'oneloop'):
            R.ImnOneLoop(ol[p1], None if p2 is None else ol[p2], kint, m, n, which, 1e-4)
    # Zel'dovich set-up X(r), Y(r) and P00/P01/P11 on the 200-point model grid
    r = np.ascontiguousarray((10.**(1.5 + (np.arange(1, 1025) - 512.5)*(7./1024)))[off::thin])
    P.X_Zel(r)
    P.Y_Zel(r)
    state = (float(g['zel_sigma8_z']), float(g['zel_k0_low']), float(g['zel_sigmasq']), g['zel_X0'], g['zel_XX'], g['zel_YY'])
    z = R.RefZeldovich(cosmo, False, state)
    km = np.ascontiguousarray(np.logspace(-3, np.log10(0.5), 200)[off::thin])
    for name in ('P00', 'P01', 'P11'):
        z(name, km)
        z(name, km)                                               # second normalisation (Jennings fits at sigma8_z / D)
    return (time.perf_counter() - t0)*thin


def cpu_baseline_sample():
    """reference C++ (oracle/_ref), 1 core (its packaged build has no OpenMP, setup.py:188): bounded sample = every
    second point of every grid of one evaluation (~10-25 s of CPU work), scaled by 2"""
    from oracle import gcl_ref as R
    if not R.available():
        return {"value": None, "unit": UNIT, "cores": 1, "kind": "reference", "sample": "oracle/_ref/libgcl_ref.so missing"}
    t0 = time.perf_counter()
    est = _cpu_eval(SEED, thin=2)
    wall = time.perf_counter() - t0
    return {"value": 1.0/est, "unit": UNIT, "cores": 1, "kind": "reference",
            "sample": "every second point of every grid of ONE evaluation on the reference C++ at its default epsrel (all 16 I, 6 J, 13 K "
                      "and sigma_v^2(k) on 250 k, the four one-loop spectra on 1000 k, 7 x 3 ImnOneLoop integrals, X(r)/Y(r) on 1024 r, "
                      "Zel'dovich P00/P01/P11 on the 200-point model grid at two normalisations): %.1f s of CPU work, x 2 = %.1f s per "
                      "evaluation; the Python P(k,mu) assembly of the reference (<1 s, docs/source/gal-power-overview.rst:68-69) is "
                      "NOT included" % (wall, est),
            "seconds_per_eval": est}


def _ref_worker(seed):
    return _cpu_eval(seed, thin=4)


def workload_config(B):
    """the workload both arms are measured on (BASELINE.json configs[1])"""
    return {"workload": "configs[1]: GalaxySpectrum P(k,mu) full model (all one-loop + Zel'dovich terms) on 1000 k x 40 mu, "
                        "every PT table rebuilt per cosmology; batch of %d distinct synthetic cosmologies per GPU per step" % B,
            "batch_per_gpu": B, "k_interp": 250, "oneloop_k": 1000, "model_k": 200, "out_grid": [NK_OUT, NMU_OUT]}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation on every host core (one process per core, one
    cosmology each: how rsdfit's MPI pool uses it); each step is a bounded sample (every 4th point of each grid)"""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    from oracle import gcl_ref as R
    if not R.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libgcl_ref.so missing on this box"}))
        return
    import multiprocessing as mp
    cores = len(os.sched_getaffinity(0)) if hasattr(os, 'sched_getaffinity') else (os.cpu_count() or 1)
    cores = max(1, min(cores, 256))
    pool = mp.get_context('fork').Pool(cores)
    for w in range(min(args.warmup, 1)):
        pool.map(_ref_worker, [SEED + i for i in range(cores)])
    t0 = time.perf_counter()
    ests = []
    for s in range(args.steps):
        ests += pool.map(_ref_worker, [SEED + 100*s + i for i in range(cores)])
    wall = time.perf_counter() - t0
    pool.close()
    # every worker processed a bounded sample standing for one full evaluation of `est` seconds
    est = float(np.mean(ests))
    value = cores/est
    sample = ("%d processes x %d steps; each step of a process = every 4th point of every grid of one full evaluation "
              "(as in cpu_baseline), time scaled by 4; mean %.1f s per evaluation per core; wall %.1f s"
              % (cores, args.steps, est, wall))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3*wall/max(args.steps, 1), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": dict(workload_config(args.batch),
                           reference_arm="full PT table rebuild per cosmology by the reference C++ at its default epsrel, one "
                                         "cosmology per host core; the reference's Python P(k,mu) assembly (< 1 s) is not included"),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "reference", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=8)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--batch', type=int, default=256, help='cosmologies per GPU per step')
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    args = ap.parse_args()
    if args.impl == 'reference':
        if args.steps > 3:
            args.steps = 3
        run_reference(args)
    else:
        if args.warmup < 3:
            args.warmup = 3
        run_gpu(args)


if __name__ == '__main__':
    main()

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
def bind_to_gpu_numa(index):
    """run this process (and first-touch its page-locked buffers) on the NUMA node the GPU hangs off: with every
    rank on node 0 the device-to-host copies of the far GPUs cross the socket interconnect (round 1: 64 GB/s aggregate
    for 8 ranks)"""
    try:
        import pynvml
        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(index)).busId
        bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()
        if len(bus.split(':')[0]) == 8:
            bus = bus[4:]
        node = int(open('/sys/bus/pci/devices/%s/numa_node' % bus).read())
        if node < 0:
            return None
        cpus = []
        for part in open('/sys/devices/system/node/node%d/cpulist' % node).read().strip().split(','):
            a, _, b = part.partition('-')
            cpus += list(range(int(a), int(b or a) + 1))
        allowed = set(os.sched_getaffinity(0)) & set(cpus)
        if allowed:
            os.sched_setaffinity(0, allowed)
            return node
    except Exception:
        pass
    return None


def run_gpu(args):
    import ctypes as C
    import torch
    import torch.distributed as dist
    from pyrsd_b200 import _native as N, parallel, pygcl, rsd

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    numa = bind_to_gpu_numa(local)
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    L = rsd._L()
    vp, dp, ci = C.c_void_p, N._dp, C.c_int
    L.prsd_ctx_profile.argtypes = [vp, ci]
    L.prsd_ctx_profile_read.argtypes = [vp, dp, np.ctypeslib.ndpointer(dtype=np.int64, flags='C_CONTIGUOUS')]   # 8 families
    L.prsd_fp64_peak.argtypes = [vp, ci, C.POINTER(C.c_double)]
    L.prsd_spectra_reload_device.argtypes = [vp, vp, vp, vp]
    L.prsd_gridop_create.argtypes = [vp, ci, ci, ci, dp, C.POINTER(vp)]
    L.prsd_gridop_apply.argtypes = [vp, ci, vp, vp, ci]
    L.prsd_gridop_destroy.argtypes = [vp]
    L.prsd_fftlog_device.argtypes = [vp, ci, ci, C.c_double, C.c_double, C.c_double, C.c_double, ci, vp, vp, vp]
    # private function objects (CDLL.__getitem__ does not cache) whose walker arguments are raw pinned-memory addresses
    gp_dev = L['prsd_galaxy_power_device']
    gp_dev.argtypes, gp_dev.restype = [vp, vp, vp, ci, vp, vp, vp, ci, vp, vp, vp], ci

    t_cold0 = time.perf_counter()
    ctx = N.default_context(local)
    sptr = vp()
    N.check(L.prsd_ctx_stream(ctx.ptr, C.byref(sptr)))
    stream = torch.cuda.ExternalStream(sptr.value, device=dev)

    B = args.batch
    data = synthetic_batch(B, rank)
    t_plan0 = time.perf_counter()
    eng = rsd.PTEngine(context=ctx)
    info = eng.info()
    gal = eng.galaxy(data['kmodel'], rsd.galaxy_config())
    plan_wall_s = time.perf_counter() - t_plan0
    npts = len(data['kk'])

    def pinned(a, dtype=torch.float64):
        t = torch.empty(a.shape, dtype=dtype, pin_memory=True)
        t.numpy()[...] = a
        return t
    h_wpar, h_stoch = pinned(data['wpar']), pinned(data['stoch'])
    d_k, d_T, d_pars = [torch.from_numpy(data[n]).to(dev) for n in ('k', 'T', 'pars')]
    d_kk, d_mm = torch.from_numpy(data['kk']).to(dev), torch.from_numpy(data['mm']).to(dev)
    d_out = torch.empty((B, npts), dtype=torch.float64, device=dev)
    # the per-walker result every rank receives at the end of a step: mu-averaged P(k), written by the library's own
    # grid-reduction kernel straight into a persistent send buffer; persistent receive buffer, no per-step allocation
    gop = vp()
    N.check(L.prsd_gridop_create(ctx.ptr, 1, NK_OUT, NMU_OUT, np.full((1, NK_OUT, NMU_OUT), 1.0/NMU_OUT), C.byref(gop)))
    d_send = torch.empty((B, NK_OUT), dtype=torch.float64, device=dev)
    d_recv = torch.empty((world*B, NK_OUT), dtype=torch.float64, device=dev)
    flush = torch.empty(64*1024*1024, dtype=torch.float32, device=dev)      # L2 flush buffer (256 MB > 126 MB)
    torch.cuda.synchronize()
    # the exchange over peer memory (NVLink / NVSwitch, CUDA IPC between the ranks); --exchange nccl or a box without peer
    # access between its GPUs: the all-gather of torch.distributed
    peer_x, exchange_kind = None, "none (one rank)"
    if world > 1:
        exchange_kind = "nccl all-gather (torch.distributed)"
        if args.exchange == 'peer':
            try:
                peer_x = parallel.PeerExchange(ctx, B*NK_OUT)      # raises on every rank or on none
            except Exception as e:                                    # noqa: BLE001 -- reported; NCCL all-gather instead
                peer_x = None
                if rank == 0:
                    print("peer exchange unavailable (%s): NCCL all-gather instead" % e, file=sys.stderr)
            if peer_x is not None:
                exchange_kind = "grid-reduction kernel -> own slot of the receive buffer, push to every peer's slot over NVLink peer memory + flag exchange (csrc/exchange.cu)"

    sp = eng.spectra(data['k'], data['T'], data['pars'][:, 0], data['pars'][:, 1], data['pars'][:, 2], data['pars'][:, 3],
                     data['pars'][:, 4], data['pars'][:, 5])
    res = rsd._new_result()

    def step_device(exchange=True, marks=None):
        """inputs resident in HBM, outputs left in HBM"""
        def mark():
            if marks is not None:
                e = torch.cuda.Event(enable_timing=True)
                e.record(stream)
                marks.append(e)
        mark()
        N.check(L.prsd_spectra_reload_device(sp.ptr, d_k.data_ptr(), d_T.data_ptr(), d_pars.data_ptr()))
        eng.run(sp, result=res, sync=False)
        N.check(gp_dev(gal.ptr, sp.ptr, res.ptr, B, None, h_wpar.data_ptr(), h_stoch.data_ptr(), npts,
                    d_kk.data_ptr(), d_mm.data_ptr(), d_out.data_ptr()))
        mark()
        if world > 1 and exchange:
            # the ensemble step's exchange, enqueued behind the kernels on the library's own stream
            if peer_x is not None:
                # the grid-reduction kernel writes into this rank's slot of its receive buffer, a push kernel stores the slot
                # into every peer's buffer over NVLink peer memory, a one-CTA flag exchange follows (csrc/exchange.cu)
                peer_x.apply(gop, B, C.c_void_p(d_out.data_ptr()))
            else:
                N.check(L.prsd_gridop_apply(gop, B, d_out.data_ptr(), d_send.data_ptr(), 1))
                mark()
                with torch.cuda.stream(stream):
                    dist.all_gather_into_tensor(d_recv, d_send)
        mark()

    first0 = time.perf_counter()
    step_device()
    torch.cuda.synchronize()
    first_call_s = time.perf_counter() - first0
    exchange_check = None
    if peer_x is not None:
        # untimed: what arrived over peer memory against the library all-gather of the same step
        got = peer_x.received(check=True).reshape(world*B, NK_OUT).clone()
        N.check(L.prsd_gridop_apply(gop, B, d_out.data_ptr(), d_send.data_ptr(), 1))
        with torch.cuda.stream(stream):
            dist.all_gather_into_tensor(d_recv, d_send)
        torch.cuda.synchronize()
        same = torch.tensor([1 if torch.equal(got, d_recv) else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
        assert int(same.item()) == 1, "peer exchange differs from the NCCL all-gather"
        exchange_check = "bit-identical to the NCCL all-gather of the same step on every rank"

    # ---- end to end through the public streaming API: every step copies its inputs host -> pinned -> device and its
    #      P(k, mu) device -> pinned host; the copy of step i runs under the PT kernels of step i + 1
    st = rsd.PTStream(eng, data['kmodel'], rsd.galaxy_config(), B, data['k'].shape[1], data['kk'], data['mm'], depth=2)

    def run_e2e_pipelined(steps):
        prev, last = None, None
        for i in range(steps):
            t = st.submit(data['k'], data['T'], data['pars'], data['wpar'], data['stoch'])
            if prev is not None:
                last = st.result(prev)             # step i - 1 is on the host while step i computes
            prev = t
        last = st.result(prev)
        st.drain()
        return last

    h_out_blk = np.empty((B, npts))

    def step_e2e_blocking():
        """the blocking public calls with host (pageable NumPy) buffers, one step at a time"""
        sp2 = eng.spectra(data['k'], data['T'], data['pars'][:, 0], data['pars'][:, 1], data['pars'][:, 2],
                          data['pars'][:, 3], data['pars'][:, 4], data['pars'][:, 5])
        r2 = eng.run(sp2, result=res)
        N.check(L.prsd_galaxy_power(gal.ptr, sp2.ptr, r2.ptr, B, None, data['wpar'], data['stoch'], npts,
                                    data['kk'], data['mm'], h_out_blk))
        sp2.close()
        return h_out_blk

    sampler = ClockSampler(local)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed_steps(fn, steps, warmup, profile=False, flush_l2=True):
        """`steps` calls of fn, each bracketed by CUDA events on the library stream, L2 flushed in between"""
        for _ in range(warmup):
            fn()
            flush.zero_()
        barrier()
        if profile:
            N.check(L.prsd_ctx_profile(ctx.ptr, 1))
        sampler.armed = True
        l0 = ctx.launches()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for e0, e1 in evs:
            if flush_l2:
                flush.zero_()                   # L2 flush between timed iterations (outside the event pair)
            torch.cuda.synchronize()
            if world > 1:
                # every timed step starts on all ranks together (the step ends in a collective: without this the event pair
                # of a rank would also hold the other ranks' host-side drift through the L2 flush above)
                dist.barrier()
                torch.cuda.synchronize()
            e0.record(stream)
            fn()
            e1.record(stream)
        torch.cuda.synchronize()
        sampler.armed = False
        barrier()
        ms = sum(e0.elapsed_time(e1) for e0, e1 in evs)
        timed_steps.last = [round(e0.elapsed_time(e1), 4) for e0, e1 in evs]
        launches = ctx.launches() - l0
        prof = None
        if profile:
            pms, pc = np.zeros(8), np.zeros(8, dtype=np.int64)
            N.check(L.prsd_ctx_profile_read(ctx.ptr, pms, pc))
            N.check(L.prsd_ctx_profile(ctx.ptr, 0))
            prof = (pms, pc)
        return parallel.max_over_ranks(ms, device=dev), launches, prof

    def timed_block(fn, warm):
        """one event pair around a whole (pipelined) run"""
        warm()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sampler.armed = True
        e0.record(stream)
        out = fn()
        e1.record(stream)
        torch.cuda.synchronize()
        sampler.armed = False
        ms = parallel.max_over_ranks(e0.elapsed_time(e1), device=dev)
        barrier()
        return ms, out

    # FP64 pipe peaks, measured once outside the timed region (MEASURED_PEAKS.json has no FP64 entry)
    pk = C.c_double()
    N.check(L.prsd_fp64_peak(ctx.ptr, 0, C.byref(pk)))
    peak_dfma = pk.value
    N.check(L.prsd_fp64_peak(ctx.ptr, 1, C.byref(pk)))
    peak_dmma = pk.value

    if rank == 0:
        sampler.start()
    # the timed steps run without instrumentation; a second pass of the same steps carries the library's per-kernel-family
    # CUDA events (ProfScope: an event pair around every family on its stream, ~0.1 ms per step of overhead)
    ms_dev, launches, _ = timed_steps(step_device, args.steps, args.warmup)
    if peer_x is not None:
        # a wait for a peer that ran into its 2 s bound (a rank held up by the host) invalidates the steps: agreed on by all
        # ranks, the measurement is then repeated with the all-gather of torch.distributed
        bad = torch.zeros(1, dtype=torch.int32, device=dev)
        try:
            peer_x.received(check=True)
        except RuntimeError:
            bad[0] = 1
        dist.all_reduce(bad, op=dist.ReduceOp.MAX)
        if int(bad.item()):
            if rank == 0:
                print("peer exchange: a wait timed out; repeating the timed steps with the NCCL all-gather", file=sys.stderr)
            peer_x, exchange_kind = None, "nccl all-gather (torch.distributed) after a peer-exchange time-out"
            ms_dev, launches, _ = timed_steps(step_device, args.steps, args.warmup)
    steps_rank0 = list(timed_steps.last)
    ms_instr, _, prof = timed_steps(step_device, args.steps, 1, profile=True)
    multi = None
    if world > 1:
        # where the N-GPU step differs from the 1-GPU one: the same steps without the exchange, per rank (GPUs of one box
        # differ by a few per cent under load; the step with the exchange runs at the pace of the slowest)
        ms_nox, _, _ = timed_steps(lambda: step_device(False), args.steps, 1)
        own = torch.tensor([0.0], dtype=torch.float64, device=dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        per = []
        for _ in range(args.steps):
            flush.zero_()
            torch.cuda.synchronize()
            e0.record(stream); step_device(False); e1.record(stream)
            torch.cuda.synchronize()
            per.append(e0.elapsed_time(e1))
        own[0] = float(np.mean(per))
        # where inside a step with the exchange the time goes (rank 0): kernels | [reduction |] exchange
        split = []
        for _ in range(4):
            flush.zero_()
            torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
            marks = []
            step_device(True, marks)
            torch.cuda.synchronize()
            split.append([round(marks[i].elapsed_time(marks[i + 1]), 4) for i in range(len(marks) - 1)])
        allr = torch.empty(world, dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(allr, own)
        multi = {"ms_per_step_without_exchange_max_over_ranks": ms_nox/args.steps,
                 "ms_per_step_without_exchange_per_rank": [round(float(x), 4) for x in allr.cpu().numpy()],
                 "exchange": "%s; %d x %d doubles per rank and step" % (exchange_kind, B, NK_OUT),
                 "exchange_check": exchange_check, "ms_of_every_timed_step_rank0": steps_rank0,
                 "ms_split_of_4_steps_rank0": split}
        multi["exchange"] = "%s; %d x %d doubles per rank and step" % (exchange_kind, B, NK_OUT)
    e2e_steps = max(8, 2*args.steps)        # the final drain (one un-hidden copy) is amortised over the run
    ms_e2e, last = timed_block(lambda: run_e2e_pipelined(e2e_steps), lambda: run_e2e_pipelined(3))
    ms_blk, _, _ = timed_steps(step_e2e_blocking, max(2, args.steps//2), 1, flush_l2=True)
    blk_steps = max(2, args.steps//2)

    # the box's device -> host ceiling for this step's result: every rank copies its 82 MB of P(k, mu) into its own
    # page-locked buffer at the same time, nothing else running (one plain cudaMemcpyAsync per step).  On one GPU this is
    # hidden under the next step's kernels; with 8 ranks on one host it is longer than a step and bounds e2e.
    h_ceiling = torch.empty((B, npts), dtype=torch.float64, pin_memory=True)

    def d2h_only(steps=4):
        with torch.cuda.stream(stream):
            for _ in range(steps):
                h_ceiling.copy_(d_out, non_blocking=True)
    ms_d2h, _ = timed_block(lambda: d2h_only(4), lambda: d2h_only(1))
    ms_d2h /= 4
    sampler.stop_flag = True

    # in-run guards: the streaming and the blocking host paths return what the device-resident path returns
    out_dev = d_out.cpu().numpy()
    assert np.all(np.isfinite(out_dev))    # (the PT model itself may go negative at k mu ~ 0.5 for extreme synthetic walkers)
    assert np.array_equal(np.asarray(last), out_dev) and np.array_equal(h_out_blk, out_dev)
    if world > 1:
        assert np.allclose(d_recv[rank*B:(rank + 1)*B].cpu().numpy(), out_dev.reshape(B, NK_OUT, NMU_OUT).mean(-1), rtol=1e-13)

    extra = other_configs(args, world, rank, dev, stream, eng, gal, ctx, L, N, parallel, data, barrier)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    # ---- roofline of the mode-coupling operators: EXECUTED FP64 flops = declared nodes x
    #      cosmologies x flops per (node, cosmology) of the implemented rule, FMA = 2.  (SURVEY 8d's nominal 538 / 210 /
    #      622 flop per node assumed on-the-fly kernel evaluation; the operator formulation evaluates the kernels once
    #      per plan, so the nominal count no longer describes the per-cosmology work.)
    pms, pc = prof
    FLOP_ML, FLOP_DIAG, FLOP_IK = 109, 20.25, 67
    kern = {
        'k_nodal_apply<SpecML>': dict(fam=6, flops=info['nodes_ml_half']*B*FLOP_ML, pipe='fp64 DFMA'),
        'k_nodal_apply<SpecDiag>': dict(fam=7, flops=info['nodes_oneloop']*B*FLOP_DIAG, pipe='fp64 DFMA'),
        'k_bilinear_apply<4>': dict(fam=0, flops=info['nodes_ik']*8*((B + 7)//8)*FLOP_IK, pipe='fp64 DMMA m8n8k4'),
    }
    traffic, limiter, prof_file = {}, {}, None
    for cand in ('r02_ncu_summary.json', 'r01_ncu_summary.json'):
        try:
            prof_json = json.load(open(os.path.join(ROOT, 'profiles', cand)))
            traffic = prof_json.get('dram_bytes_per_launch', {})
            for kname, kd in prof_json.get('kernels', {}).items():
                limiter[kname] = {m.split('.')[0]: float(kd[m]['value']) for m in (
                    'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
                    'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
                    'sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active',
                    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed') if m in kd}
            prof_file = 'profiles/' + cand
            break
        except Exception:
            continue
    for name, kd in kern.items():
        n = max(int(pc[kd['fam']]), 1)
        kd['ms'] = pms[kd['fam']]/n
        kd['launches_per_step'] = n//args.steps
        kd['flops'] = float(kd['flops'])
        kd['achieved'] = kd['flops']/(kd['ms']*1e-3)/1e12 if kd['ms'] > 0 else None
        kd['share'] = pms[kd['fam']]/ms_instr
    top = max(kern, key=lambda k: kern[k]['share'])
    peak = peak_dmma if 'DMMA' in kern[top]['pipe'] else peak_dfma
    fam = ['ik_dmma', 'linear_ops', 'zeldovich', 'galaxy', 'spline_builds', 'fftlog', 'imn_oneloop_nodal', 'oneloop_tables_nodal']
    shares = {fam[i]: round(pms[i]/ms_instr, 4) for i in range(8)}
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    value = world*B*args.steps/(ms_dev*1e-3)
    e2e_value = world*B*e2e_steps/(ms_e2e*1e-3)
    h2d, d2h = st.bytes_per_step()
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_dev/args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": dict(workload_config(B), quadrature_nodes_per_cosmology=info['nodes'],
                       l2="value: L2 flushed between timed iterations (256 MB write); e2e: consecutive pipelined steps, each "
                          "streams 0.36 GB of quadrature operators + 82 MB of output (> 126 MB L2)",
                       parallelism="walkers sharded over %d rank(s), no data-path collective; the mu-averaged P(k) of every "
                                   "walker reaches every rank inside the timed step (%s)" % (world, exchange_kind),
                       numa_node_of_rank0=numa),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "ms_per_step": ms_e2e/e2e_steps, "steps": e2e_steps,
                "d2h_only_ms_per_step": ms_d2h,
                "d2h_only_note": "the device -> host copy of one step's P(k, mu) alone, all %d rank(s) copying at once (max over "
                                 "ranks): the floor of the e2e step on this box whatever the kernels do" % world,
                "how": "rsd.PTStream (prsd_spectra_reload / prsd_plan_run_async / prsd_galaxy_power_async): every step copies its "
                       "inputs (spectra T(k), cosmological and walker parameters, stochasticity) from NumPy arrays to page-locked "
                       "slots and to the device and its P(k, mu) back to page-locked host memory; the device-to-host copy of step i "
                       "runs on a second stream under the PT kernels of step i + 1 and the input copies of step i + 1 on a third "
                       "stream under the kernels of step i; the fixed (k, mu) evaluation grid goes up once with the first batch; one "
                       "CUDA-event pair around all steps including the final drain",
                "blocking": {"value": world*B*blk_steps/(ms_blk*1e-3), "ms_per_step": ms_blk/blk_steps,
                             "how": "prsd_spectra_create / prsd_plan_run / prsd_galaxy_power with pageable host buffers, "
                                    "one step at a time (host waits for every result)"}},
        "gpu_launches": int(launches),
        "multi_gpu": multi,
        "roofline": {"bound": "fp64 (DFMA pipe; ncu: the binding unit is the LSU / shared-memory data pipe, see ncu_pct_of_peak)",
                     "achieved": kern[top]['achieved'], "peak": peak, "unit": "TFLOP/s",
                     "frac": (kern[top]['achieved']/peak) if (kern[top]['achieved'] and peak) else None,
                     "traffic": traffic.get(top), "flops_basis": "executed (declared nodes x cosmologies x flops of the implemented "
                                                                 "rule; SURVEY 8d's nominal per-node counts assumed on-the-fly kernel "
                                                                 "evaluation, which the operator formulation hoists into the plan)",
                     "ncu_pct_of_peak": limiter.get(top), "ncu_source": prof_file,
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
                     "kernel_share_of_step": shares,
                     "instrumented_ms_per_step": ms_instr/args.steps,
                     "instrumentation": "kernel times and shares come from a second pass of the same steps with the library's "
                                        "per-family CUDA-event pairs switched on; the timed steps behind `value` run without them"},
        "clocks": sampler.summary(),
        "cold_start": {"plan_build_s": info['build_seconds'], "engine_and_galaxy_ctor_wall_s": plan_wall_s,
                       "first_step_s": first_call_s, "process_start_to_first_result_s": first0 + first_call_s - t_cold0},
        "configs": extra,
    }
    if not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline_sample()
        line["oracle_check"] = oracle_check(eng, res, data, B)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def other_configs(args, world, rank, dev, stream, eng, gal, ctx, L, N, parallel, data, barrier):
    """the other BASELINE.json configurations and the single-evaluation latency, each timed with CUDA events on the
    library stream (max over ranks); small step counts: these are reported next to the headline, not instead of it"""
    import ctypes as C
    import torch
    import torch.distributed as dist
    from pyrsd_b200 import pygcl, rsd
    from pyrsd_b200.synthetic import base_cosmology
    out = {}
    if args.no_configs:
        return out

    def ev_time(fn, reps, warm=1):
        for _ in range(warm):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            fn()
        e1.record(stream)
        torch.cuda.synchronize()
        ms = parallel.max_over_ranks(e0.elapsed_time(e1)/reps, device=dev)
        barrier()
        return ms

    # ---- C1 (configs[0]): pygcl OneLoopP00 (= P_L + one-loop P_dd) + ZeldovichP00 at z = 0.55, 200 k in [0.005, 0.5]
    if rank == 0:
        c = base_cosmology()
        params = dict(h=c['h'], Omega_b=c['Omega0_b'], Omega_cdm=c['Omega0_m'] - c['Omega0_b'], n_s=c['n_s'], T_cmb=c['Tcmb'],
                      growth={0.0: 1.0, 0.55: 0.7486}, f={0.0: 0.5, 0.55: 0.7602}, H={0.0: 100.*c['h'], 0.55: 133.86*c['h']})
        k200 = np.logspace(np.log10(0.005), np.log10(0.5), 200)
        t0 = time.perf_counter()
        cosmo = pygcl.Cosmology(params, 0, c['sigma8'], c['k'], c['T'])
        PL = pygcl.LinearPS(cosmo, 0.55)
        a = pygcl.OneLoopPdd(PL).EvaluateFull(k200)
        b = pygcl.ZeldovichP00(cosmo, 0.55)(k200)
        t_first = time.perf_counter() - t0

        def c1():
            PLz = pygcl.LinearPS(cosmo, 0.55)
            pygcl.OneLoopPdd(PLz).EvaluateFull(k200)
            pygcl.ZeldovichP00(cosmo, 0.55)(k200)
        t0 = time.perf_counter()
        for _ in range(3):
            c1()
        out['C1_pygcl_oneloop_plus_zeldovich_200k'] = {
            "ms_per_call": 1e3*(time.perf_counter() - t0)/3, "first_call_s": t_first, "timer": "host wall clock (blocking one-off pygcl calls)",
            "finite": bool(np.all(np.isfinite(a)) and np.all(np.isfinite(b)))}
    # ---- C3 (configs[2]): Imn / Jmn / Kmn + one-loop table rebuild for 1024 variants (sharded over the ranks)
    n3 = 1024//world
    d3 = synthetic_batch(n3, rank + 100, nk_out=8, nmu_out=2)
    sp3 = eng.spectra(d3['k'], d3['T'], d3['pars'][:, 0], d3['pars'][:, 1], d3['pars'][:, 2], d3['pars'][:, 3], d3['pars'][:, 4], d3['pars'][:, 5])
    r3 = rsd._new_result()
    dk3, dT3, dp3 = [torch.from_numpy(d3[n]).to(dev) for n in ('k', 'T', 'pars')]

    def c3():
        N.check(L.prsd_spectra_reload_device(sp3.ptr, dk3.data_ptr(), dT3.data_ptr(), dp3.data_ptr()))
        eng.run(sp3, result=r3, sync=False)
    ms = ev_time(c3, 3)
    out['C3_pt_tables_1024_variants'] = {"ms_per_rebuild": ms, "variants_per_s": 1024/(ms*1e-3), "variants_per_rank": n3,
                                         "what": "16 I, 6 J, 13 K, 4 one-loop tables x 1000 k, 21 ImnOneLoop parts, sigma_v^2(k), X(r), Y(r)"}
    sp3.close()
    # ---- C4 (configs[3]): FFTLog Hankel transforms, 4096-point log grid, 16384 sequences sharded over the ranks
    n4, N4 = 16384//world, 4096
    a4 = torch.rand((n4, N4), dtype=torch.float64, device=dev)
    o4 = torch.empty_like(a4)
    torch.cuda.synchronize()

    def c4():
        N.check(L.prsd_fftlog_device(ctx.ptr, n4, N4, 0.5, 0.0, np.log(1e7)/N4, 1.0, 1, a4.data_ptr(), o4.data_ptr(), None))
    ms = ev_time(c4, 5, warm=2)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    gbs = 16.0*N4*n4/(ms*1e-3)/1e9
    flops = (5*N4*np.log2(N4) + 43*N4)*n4/(ms*1e-3)/1e12
    out['C4_fftlog_16384x4096'] = {"ms": ms, "transforms_per_s": 16384/(ms*1e-3), "per_rank": n4, "hbm_gbs_per_gpu": gbs,
                                   "hbm_frac": gbs/peaks['hbm_gbs'] if peaks.get('hbm_gbs') else None, "fp64_tflops_per_gpu": flops,
                                   "algorithmic_bytes": "16 N per transform (read + write)", "inputs": "resident in HBM (3 x L2)"}
    del a4, o4
    # ---- C5 (configs[4]): emcee-style ensemble, 2048 walkers in total (STRONG scaling): every walker = a new cosmology,
    #      all PT tables rebuilt, multipoles on 40 k x (l = 0, 2, 4), chi^2 against a data vector; chi^2 gathered over NCCL
    n5 = 2048//world
    d5 = synthetic_batch(n5, rank + 200, nk_out=8, nmu_out=2)
    sp5 = eng.spectra(d5['k'], d5['T'], d5['pars'][:, 0], d5['pars'][:, 1], d5['pars'][:, 2], d5['pars'][:, 3], d5['pars'][:, 4], d5['pars'][:, 5])
    r5 = rsd._new_result()
    dk5, dT5, dp5 = [torch.from_numpy(d5[n]).to(dev) for n in ('k', 'T', 'pars')]
    kd = np.linspace(0.02, 0.4, 40)
    nd = 3*len(kd)
    rng = np.random.default_rng(5)
    A = rng.normal(size=(nd, nd))*0.01
    cinv = np.eye(nd) + A@A.T
    datav = rng.uniform(1e3, 1e4, nd)
    wp5, st5 = torch.from_numpy(d5['wpar']).pin_memory(), torch.from_numpy(d5['stoch']).pin_memory()
    chi_all = torch.empty(2048, dtype=torch.float64, device=dev)
    poles_dev = L['prsd_galaxy_poles_device']
    poles_dev.argtypes = [C.c_void_p]*3 + [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
    poles_dev.restype = C.c_int
    d_kd = torch.from_numpy(kd).to(dev)
    d_poles = torch.empty((n5, 3, len(kd)), dtype=torch.float64, device=dev)
    ells = np.array([0, 2, 4], dtype=np.int32)
    chi_host = np.empty(n5)
    d_chi = torch.empty(n5, dtype=torch.float64, device=dev)

    def c5():
        N.check(L.prsd_spectra_reload_device(sp5.ptr, dk5.data_ptr(), dT5.data_ptr(), dp5.data_ptr()))
        eng.run(sp5, result=r5, sync=False)
        N.check(poles_dev(gal.ptr, sp5.ptr, r5.ptr, n5, None, wp5.data_ptr(), st5.data_ptr(), len(kd),
                          d_kd.data_ptr(), 3, ells.ctypes.data_as(C.c_void_p), 41, d_poles.data_ptr()))
        N.check(L.prsd_chi2(ctx.ptr, n5, nd, C.c_void_p(d_poles.data_ptr()), datav, cinv, chi_host, 1))
        with torch.cuda.stream(stream):
            d_chi.copy_(torch.from_numpy(chi_host))
            if world > 1:
                dist.all_gather_into_tensor(chi_all, d_chi)
    ms = ev_time(c5, 2)
    out['C5_ensemble_2048_walkers'] = {"ms_per_ensemble_step": ms, "walker_evals_per_s": 2048/(ms*1e-3), "walkers_per_rank": n5,
                                       "scaling": "strong (2048 walkers fixed)", "finite": bool(np.all(np.isfinite(chi_host))),
                                       "what": "per walker: all PT tables rebuilt + Zel'dovich + P_0, P_2, P_4 on 40 k + chi^2; chi^2 all-gathered"}
    sp5.close()
    # ---- the north-star latency target: ONE cosmology, all tables rebuilt, P(k, mu) on the full 1000 x 40 grid, host in / host out
    if rank == 0:
        d1 = synthetic_batch(1, 0)
        out1 = np.empty((1, len(d1['kk'])))

        def one():
            sp1 = eng.spectra(d1['k'], d1['T'], d1['pars'][:, 0], d1['pars'][:, 1], d1['pars'][:, 2], d1['pars'][:, 3], d1['pars'][:, 4], d1['pars'][:, 5])
            r1 = eng.run(sp1)
            N.check(L.prsd_galaxy_power(gal.ptr, sp1.ptr, r1.ptr, 1, None, d1['wpar'], d1['stoch'], len(d1['kk']), d1['kk'], d1['mm'], out1))
            sp1.close()
        for _ in range(3):
            one()
        ts = []
        for _ in range(10):
            t0 = time.perf_counter()
            one()
            ts.append(time.perf_counter() - t0)
        out['latency_single_eval_ms'] = {"median": 1e3*float(np.median(ts)), "min": 1e3*float(np.min(ts)),
                                         "what": "one cosmology: every PT table + Zel'dovich + P(k, mu) on 1000 k x 40 mu, NumPy in, NumPy "
                                                 "out, blocking calls, host wall clock", "target_ms": 1.0}
    barrier()
    return out


def oracle_check(eng, res, data, B):
    """three cosmologies of THIS run's batch (the first two and the one with the largest tilt) against the reference C++
    (oracle/_ref) at tight tolerance on a few k: I00, I22, K11 (relative to max(|value|, P_L)), J02 -- untimed"""
    from oracle import gcl_ref as R
    if not R.available():
        return {"status": "oracle/_ref/libgcl_ref.so missing"}
    from pyrsd_b200.synthetic import base_cosmology
    c = base_cosmology()
    I, J, K = eng.get(res, 'I'), eng.get(res, 'J'), eng.get(res, 'K')
    kint = eng.k_interp
    sel = np.array([100, 150, 190, 215, 232, 244])
    worst = 0.0
    which = sorted(set([0, min(1, B - 1), int(np.argmax(np.abs(data['delta'])))]))
    for i in which:
        T = c['T']*(c['k']/0.05)**(0.5*data['delta'][i])
        rc = R.RefCosmology(c['k'], T, c['sigma8'], c['n_s'], c['h'], c['Omega0_m'], c['Omega0_b'], c['Tcmb'])
        PL = R.RefLinearPS(rc)
        A = (rc.delta_H()/c['delta_H'])**2            # the oracle normalises to sigma8, the batch keeps delta_H fixed
        k = np.ascontiguousarray(kint[sel])
        floor = PL(k)
        for mine, ref in ((A*A*I[i, 0][sel], R.Imn(PL, k, 0, 0, 1e-8)), (A*A*I[i, 10][sel], R.Imn(PL, k, 2, 2, 1e-8)),
                          (A*A*K[i, 6][sel], R.Kmn(PL, k, 1, 1, False, 0, 1e-8))):
            worst = max(worst, float(np.max(np.abs(mine - ref)/np.maximum(np.abs(ref), floor))))
        worst = max(worst, float(np.max(np.abs(A*J[i, 2][sel]/R.Jmn(PL, k, 0, 2, 1e-10) - 1))))
    return {"cosmologies_of_this_batch": which, "k": [float(x) for x in kint[sel]], "terms": ["I00", "I22", "K11", "J02"],
            "max_err": worst, "tolerance": 1e-5, "pass": bool(worst < 1e-5)}


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
    ap.add_argument('--exchange', default='peer', choices=['peer', 'nccl'],
                    help='N > 1: how the per-walker results reach every rank (peer-memory stores, or torch.distributed all-gather)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-configs', action='store_true', help='skip the other BASELINE configurations')
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

#!/usr/bin/env python
"""bench.py -- log-L evals/s (sample x source x epoch) of the astrometric likelihood.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[4], the one the metric is quoted on): 2^22 parameter
vectors PER GPU x 8 sources x 1024 epochs, full model (shared parallax / proper motion
/ EFAC, source 0 a binary with Omega_asc + inclination: "variant C", P = 22).
One step = one pass of the fused likelihood kernel over the batch (+ the NCCL gather
of the log-L values when N > 1).  Weak scaling: per-GPU batch fixed.

Prints ONE JSON line on rank 0 (see DESIGN.md "Measurement" for every key).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = 'log-L evals/sec (sample x source x epoch)'
UNIT = 'evals/s'
FLOPS_PER_EVAL = {'A': 30.0, 'B': 42.0, 'C': 43.5}      # algorithmic FP64 flops, SURVEY.md 8(d)
LOG2_N = 22


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    QUERY = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
             'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
             'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.gpu), '--query-gpu=' + self.QUERY,
                                          '--format=csv,noheader,nounits', '-lms', '200'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons, power = [], [], set(), []
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for line in self.lines:
            f = [x.strip() for x in line.split(',')]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        if not sm:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['no samples']}
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2], 'sm_max_mhz': max(smax), 'reasons': sorted(reasons),
                'samples': len(sm), 'power_w_max': max(power)}


def measured_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    try:
        with open(path) as f:
            return json.load(f), 'measured'
    except Exception:
        return {'hbm_gbs': 6650.0}, 'fallback'


# --------------------------------------------------------------------------------------
# reference arm: the CPU restatement of the reference's algorithm on all host cores
# --------------------------------------------------------------------------------------
def run_reference(args, rank, world):
    if rank != 0:
        return 0
    from oracle import cpu_bench                      # the one place bench.py executes oracle/
    cores = os.cpu_count() or 1
    rows = cpu_rows(args, cpu_bench, cores, target_s=10.0)      # includes the warm-up call
    walls = []
    for _ in range(args.steps):
        v, w, c = cpu_bench.time_c_oracle('C', rows=rows, threads=cores)
        walls.append(w)
    value = args.steps * rows * 8 * 1024 / sum(walls)
    sample = ('%d parameter vectors x 8 sources x 1024 epochs per step (of the 2^%d-vector workload); C restatement '
              'of simulate.py:362-384 in the reference\'s operation order (oracle/c/sterne_oracle.c), OpenMP over %d '
              'host threads' % (rows, LOG2_N, cores))
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * sum(walls) / args.steps,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': workload_config(args.gpus),
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
        'note': 'dingswin/sterne itself cannot be installed here (astropy, bilby, novas/DE421 absent, no '
                'network); this is the oracle port of its algorithm, not a measurement of sterne.',
    }
    print(json.dumps(line))
    return 0


def cpu_rows(args, cpu_bench, cores, target_s):
    """Size of the bounded CPU sample: --cpu-rows if given, else enough parameter vectors for about
    `target_s` seconds on this host (calibrated with a short run that doubles as the warm-up)."""
    if args.cpu_rows > 0:
        return args.cpu_rows
    cpu_bench.time_c_oracle('C', rows=cores * 64, threads=cores)
    v, _, _ = cpu_bench.time_c_oracle('C', rows=cores * 2048, threads=cores)
    rows = int(v * target_s / (8 * 1024))
    return max(cores * 256, min(rows, 1 << LOG2_N))


def workload_config(n_gpus):
    return {'workload': 'config5-C: 2^%d parameter vectors per GPU x 8 sources x 1024 epochs, P=22 '
                        '(shared px, mu_a, mu_d, efac; source 0 binary with incl, om_asc)' % LOG2_N,
            'samples_per_gpu': 1 << LOG2_N, 'sources': 8, 'epochs_per_source': 1024, 'params': 22,
            'parallelism': 'dp%d (sample batch sharded, epoch tables replicated)' % n_gpus,
            'l2': 'theta is 738 MB per GPU, larger than the 126 MB L2: no flush needed between steps',
            'kernel_mode': 'fast'}


# --------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------
def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the headline kernel, read from the newest
    committed ncu summary of this command (profiles/rNN_ncu_loglike_fast_C.txt); (bytes, file) or (None, None)."""
    import glob
    import re
    files = sorted(glob.glob(os.path.join(ROOT, 'profiles', 'r*_ncu_loglike_fast_C.txt')))
    if not files:
        return None, None
    text = open(files[-1]).read()
    tot = 0.0
    for key in ('dram__bytes_read.sum', 'dram__bytes_write.sum'):
        m = re.search(r'^%s\s+(\w+)\s+([0-9.eE+-]+)' % re.escape(key), text, re.M)
        if not m:
            return None, None
        unit = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(m.group(1))
        if unit is None:
            return None, None
        tot += float(m.group(2)) * unit
    return int(tot), os.path.relpath(files[-1], ROOT)


class SharedResult(object):
    """One (n,) float64 host array shared by all ranks of the box (POSIX shared memory), pinned in every
    rank: each GPU DMAs its slice of log L straight into it, so the result of an N-GPU evaluation is
    assembled in ONE host array without a collective.  Falls back to None when /dev/shm is too small."""

    def __init__(self, n, rank, dist, group):
        import numpy as np
        from multiprocessing import shared_memory
        from sterne_b200 import host_register
        self.shm = None
        name = [None]
        if rank == 0:
            try:
                st = os.statvfs('/dev/shm')
                if st.f_bavail * st.f_frsize > 2 * n * 8 + (64 << 20):
                    self.shm = shared_memory.SharedMemory(create=True, size=n * 8)
                    name[0] = self.shm.name
            except Exception:
                self.shm = None
        dist.broadcast_object_list(name, src=0, group=group)
        self.array = None
        if name[0] is None:
            return
        if rank != 0:
            self.shm = shared_memory.SharedMemory(name=name[0])
            try:                                   # only the creating rank may unlink it (Python < 3.13 tracks attachers too)
                from multiprocessing import resource_tracker
                resource_tracker.unregister(self.shm._name, 'shared_memory')
            except Exception:
                pass
        self.array = np.ndarray((n,), dtype=np.float64, buffer=self.shm.buf)
        self.pin = host_register(self.array)
        self.owner = rank == 0

    def close(self):
        if self.shm is None:
            return
        try:
            self.pin.release()
        except Exception:
            pass
        self.array = None
        self.pin = None
        try:
            self.shm.close()
            if self.owner:
                self.shm.unlink()
        except Exception:
            pass


def run_ours(args, rank, local_rank, world):
    import numpy as np
    import torch
    import torch.distributed as dist
    from sterne_b200 import synthetic, fp64_peak
    from sterne_b200.sharding import ShardedLikelihood, shard_bounds

    # CPU baseline first (rank 0, N = 1 only), before this process touches CUDA
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import cpu_bench                  # cpu_baseline leg: checker timed, never shipped
        cores = os.cpu_count() or 1
        args.cpu_rows = cpu_rows(args, cpu_bench, cores, target_s=15.0)
        v, w, c = cpu_bench.time_c_oracle('C', rows=args.cpu_rows, threads=cores)
        nv, nw, nc = cpu_bench.time_vectorised('C', rows_per_core=1024, cores=cores)
        sv, sw = cpu_bench.time_scalar('C', rows=2)
        cpu_baseline = {
            'value': v, 'unit': UNIT, 'cores': c, 'kind': 'port',
            'sample': '%d parameter vectors of the same 8-source x 1024-epoch problem (%.1f s wall on %d threads); '
                      'C restatement in the reference\'s operation order, OpenMP' % (args.cpu_rows, w, c),
            'numpy_vectorised_value': nv, 'numpy_vectorised_cores': nc,
            'numpy_vectorised_sample': '1024 rows/core, one process per core (%.1f s)' % nw,
            'scalar_value': sv, 'scalar_cores': 1,
            'scalar_sample': '2 rows through the reference-shaped per-vector Python loop (%.1f s)' % sw}

    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device; sterne_b200 has no CPU path')
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    cpu_group = None
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
        cpu_group = dist.new_group(backend='gloo')         # host-side barriers that keep the GPUs idle

    def host_barrier():
        if world > 1:
            dist.barrier(group=cpu_group)

    # weak scaling (default, the contract's): 2^22 vectors per GPU.  --scaling strong: 2^22 vectors in
    # total, 2^22 / world per GPU (SURVEY 8e's reading of the 1 -> 8 GPU target).
    N = (1 << LOG2_N) if args.scaling == 'weak' else (1 << LOG2_N) // world
    case = synthetic.config5('C')
    like = case.likelihood(device=local_rank)
    ctx = like._context(None)
    E = case.evals_per_sample
    P = case.n_params
    theta = synthetic.device_theta(case, N, seed=1234 + rank, device=dev)
    out = torch.empty(N, dtype=torch.float64, device=dev)
    sharded = ShardedLikelihood(like) if world > 1 else None
    # the gather fused into the kernel needs the ranks' buffers mapped into each other (CUDA IPC); where that is
    # impossible every rank falls back to one NCCL all_gather per step, and the line says so
    fused_gather = bool(sharded.fused_available((1 << LOG2_N) * world)) if world > 1 else False

    def sync_all():
        torch.cuda.synchronize(dev)
        host_barrier()
        torch.cuda.synchronize(dev)

    def timed_device_steps(th, n_global, steps, warm):
        """K steps of: fused likelihood kernel over this rank's slice (+ for N > 1 its epilogue stores into
        every rank's result buffer over NVLink, then the stream-ordered barrier).  Returns
        (total ms max over ranks, mean kernel-only ms on this rank, this rank's result tensor, launches)."""
        n_local = int(th.shape[0])
        plan = ctx.auto_plan(n_global)
        res = out[:n_local]

        def one(evs=None):
            nonlocal res
            if evs is not None:
                evs[0].record()
            if world == 1:
                like.log_likelihood_batch(th, out=res, plan=plan)
            elif fused_gather:
                res = sharded.evaluate_scatter(th, n_global, plan=plan, barrier=False)
            else:
                like.log_likelihood_batch(th, out=out[:n_local], plan=plan)
            if evs is not None:
                evs[1].record()
            if world > 1 and fused_gather:
                sharded.barrier()
            elif world > 1:
                res = sharded.all_gather(out[:n_local], n_global)

        for _ in range(warm):
            one()
        sync_all()
        l0 = ctx.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        kev = []
        e0.record()
        for _ in range(steps):
            kev.append((torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)))
            one(kev[-1])
        e1.record()
        sync_all()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        kernel_ms = sum(a.elapsed_time(b) for a, b in kev) / steps
        return t.item(), kernel_ms, res, ctx.launch_count() - l0

    warm = max(args.warmup, 3)
    timed_device_steps(theta, N * world, 1, warm)                 # warm-up (also builds the peer mappings)
    # FP64 roofline denominator, measured with the GPU already warm (37 ms per launch)
    peak_tf, _ = fp64_peak(local_rank, iters=1 << 17, reps=5)
    sync_all()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_total, kernel_ms, res_dev, launches = timed_device_steps(theta, N * world, args.steps, 0)
    clocks = sampler.stop() if rank == 0 else None
    dev_result = res_dev.clone()                                   # (N * world,) for N > 1, (N,) for N = 1

    # ---- end to end: HOST buffers through the public API, copies inside the timed region.  N = 1: one call
    # of Gaussianlikelihood.log_likelihood_batch(numpy) (pinned, then pageable).  N > 1 (one process per GPU):
    # every rank makes that call on its slice and DMAs its slice of the result into ONE host array shared by
    # all ranks (POSIX shared memory pinned in every rank): the result is assembled in one place, no collective.
    def timed_host_steps(th_np, n_global, lo, out_np, steps):
        plan = ctx.auto_plan(n_global)
        n_local = th_np.shape[0]
        for _ in range(2):
            like.log_likelihood_batch(th_np, out=out_np[lo:lo + n_local], plan=plan)
        sync_all()
        t0 = time.perf_counter()
        for _ in range(steps):
            like.log_likelihood_batch(th_np, out=out_np[lo:lo + n_local], plan=plan)   # synchronous: result is on the host
            host_barrier()                                                               # ... of every rank
        ms = (time.perf_counter() - t0) * 1e3
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    e2e_steps = max(2, min(args.steps, 5))
    theta_host = torch.empty((N, P), dtype=torch.float64, pin_memory=True)
    theta_host.copy_(theta)
    th_np = theta_host.numpy()
    lo_w = rank * N
    shared = SharedResult(N * world, rank, dist, cpu_group) if world > 1 else None
    if shared is not None and shared.array is not None:
        out_np, assembled = shared.array, 'one host array shared by all ranks (POSIX shm, pinned in each rank)'
    else:
        out_host = torch.empty(N * world, dtype=torch.float64, pin_memory=True)
        out_np, assembled = out_host.numpy(), ('rank-local pinned array' if world > 1 else 'the caller\'s pinned array')
    e2e_ms = timed_host_steps(th_np, N * world, lo_w, out_np, e2e_steps)
    if shared is not None and shared.array is not None:
        same = bool(np.array_equal(out_np, dev_result.cpu().numpy())) if rank == 0 else True   # the WHOLE assembled result
    else:
        same = bool(np.array_equal(out_np[lo_w:lo_w + N], dev_result[lo_w:lo_w + N].cpu().numpy()))
    e2e_pageable = None
    if world == 1:
        th_page = np.array(th_np, copy=True)                       # ordinary (pageable) numpy arrays, as a sampler has
        out_page = np.empty(N)
        pg_ms = timed_host_steps(th_page, N, 0, out_page, e2e_steps)
        e2e_pageable = {'value': N * E * e2e_steps / (pg_ms * 1e-3), 'unit': UNIT, 'ms_per_step': pg_ms / e2e_steps,
                        'matches_device_path': bool(np.array_equal(out_page, dev_result.cpu().numpy())),
                        'note': 'pageable numpy in / out: staged through the library\'s pinned ring (stn_loglike_host)'}
        del th_page, out_page

    # ---- N > 1 extras: bits independent of the GPU count; strong scaling (2^22 vectors in total); ONE process
    # driving all GPUs
    sharded_bits_match, strong, single_process = None, None, None
    if world > 1:
        n_probe = 1 << 16
        probe = synthetic.device_theta(case, n_probe, seed=99, device=dev)            # the same batch on every rank
        one_gpu = like.log_likelihood_batch(probe, plan=ctx.auto_plan(n_probe))
        fused = sharded.log_likelihood_batch(probe)
        nccl = sharded.log_likelihood_batch(probe, fused=False)
        ok = torch.tensor([float(torch.equal(fused, one_gpu) and torch.equal(nccl, one_gpu))], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        sharded_bits_match = bool(ok.item() == 1.0)

        n_tot = 1 << LOG2_N
        lo_s, hi_s = shard_bounds(n_tot, world, rank)
        th_s = theta[:hi_s - lo_s]
        s_ms, s_kernel_ms, s_res, _ = timed_device_steps(th_s, n_tot, args.steps, 2)
        s_dev = s_res.clone()
        s_e2e_ms = timed_host_steps(th_np[:hi_s - lo_s], n_tot, lo_s, out_np, e2e_steps)
        if shared is not None and shared.array is not None:
            s_same = bool(np.array_equal(out_np[:n_tot], s_dev.cpu().numpy())) if rank == 0 else True
        else:
            s_same = bool(np.array_equal(out_np[lo_s:hi_s], s_dev[lo_s:hi_s].cpu().numpy()))
        strong = {'scaling': 'strong', 'vectors_total': n_tot, 'vectors_per_gpu': hi_s - lo_s,
                  'value': n_tot * E * args.steps / (s_ms * 1e-3), 'unit': UNIT, 'ms_per_step': s_ms / args.steps,
                  'kernel_ms': s_kernel_ms,
                  'e2e': {'value': n_tot * E * e2e_steps / (s_e2e_ms * 1e-3), 'unit': UNIT,
                          'ms_per_step': s_e2e_ms / e2e_steps, 'h2d_bytes_per_step': n_tot * P * 8,
                          'd2h_bytes_per_step': n_tot * 8, 'result_assembled_in': assembled,
                          'matches_device_path': s_same}}
        # one process, all GPUs (what a single emcee / dynesty process gets): rank 0 drives every device of the
        # box through Gaussianlikelihood(device=[0..N-1]) while the other ranks wait on a host-side barrier
        sync_all()
        if rank == 0:
            sp_like = case.likelihood(device=list(range(world)))
            n_tot = int(th_np.shape[0])                               # rank 0's whole host batch (2^22 in the weak run)
            sp_out = torch.empty(n_tot, dtype=torch.float64, pin_memory=True).numpy()
            for _ in range(2):
                sp_like.log_likelihood_batch(th_np, out=sp_out)
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                sp_like.log_likelihood_batch(th_np, out=sp_out)
            sp_ms = (time.perf_counter() - t0) * 1e3
            ref_one = like.log_likelihood_batch(theta, plan=ctx.auto_plan(n_tot)).cpu().numpy()
            th_page = np.array(th_np, copy=True)
            sp_page = np.empty(n_tot)
            sp_like.log_likelihood_batch(th_page, out=sp_page)
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                sp_like.log_likelihood_batch(th_page, out=sp_page)
            sp_page_ms = (time.perf_counter() - t0) * 1e3
            single_process = {
                'api': 'Gaussianlikelihood(device=[0..%d]).log_likelihood_batch(numpy (2^22, 22)) -> '
                       'stn_multi_loglike_host' % (world - 1),
                'vectors_total': n_tot, 'e2e_value': n_tot * E * e2e_steps / (sp_ms * 1e-3), 'unit': UNIT,
                'ms_per_step': sp_ms / e2e_steps, 'bits_match_one_gpu': bool(np.array_equal(sp_out, ref_one)),
                'pageable_e2e_value': n_tot * E * e2e_steps / (sp_page_ms * 1e-3),
                'pageable_ms_per_step': sp_page_ms / e2e_steps,
                'pageable_bits_match_one_gpu': bool(np.array_equal(sp_page, ref_one))}
            sp_like.close()
            del th_page, sp_page
        sync_all()

    # ---- variants A and B of the same kernel (kernel-only, rank 0, N = 1)
    variants = {}
    if world == 1 and not args.no_variants:
        for vname in ('A', 'B'):
            vc = synthetic.config5(vname)
            vl = vc.likelihood(device=local_rank)
            vt = synthetic.device_theta(vc, N, seed=1234, device=dev)
            vo = torch.empty(N, dtype=torch.float64, device=dev)
            for _ in range(3):
                vl.log_likelihood_batch(vt, out=vo)
            torch.cuda.synchronize(dev)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(args.steps):
                vl.log_likelihood_batch(vt, out=vo)
            b.record()
            torch.cuda.synchronize(dev)
            ms = a.elapsed_time(b) / args.steps
            ev = N * vc.evals_per_sample / (ms * 1e-3)
            variants[vname] = {'params': vc.n_params, 'ms_per_step': ms, 'evals_per_s': ev,
                               'fp64_tflops': ev * FLOPS_PER_EVAL[vname] / 1e12,
                               'fp64_frac': ev * FLOPS_PER_EVAL[vname] / 1e12 / peak_tf}
            vl.close()
            del vt, vo

    # ---- batch-size sweep of the headline variant (BASELINE config 5: N = 2^16 .. 2^22), kernel-only
    sweep = {}
    if world == 1 and not args.no_variants:
        for l2 in (16, 17, 18, 19, 20, 21):
            n = 1 << l2
            so = torch.empty(n, dtype=torch.float64, device=dev)
            reps = max(args.steps, 1 << max(0, 20 - l2))
            for _ in range(3):
                like.log_likelihood_batch(theta[:n], out=so)
            torch.cuda.synchronize(dev)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(reps):
                like.log_likelihood_batch(theta[:n], out=so)
            b.record()
            torch.cuda.synchronize(dev)
            ms = a.elapsed_time(b) / reps
            sweep['2^%d' % l2] = {'ms_per_step': ms, 'evals_per_s': n * E / (ms * 1e-3), 'plan': ctx.auto_plan(n)}

    line = None
    if rank == 0:
        peaks, peaks_src = measured_peaks()
        evals_per_step_rank = N * E
        value = world * evals_per_step_rank * args.steps / (ms_total * 1e-3)
        k_evals = evals_per_step_rank / (kernel_ms * 1e-3)
        achieved_tf = k_evals * FLOPS_PER_EVAL['C'] / 1e12
        alg_bytes = N * (8.0 * P + 8.0)
        traffic, traffic_file = ncu_traffic()
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
            'warmup': warm, 'ms_per_step': ms_total / args.steps, 'higher_is_better': True,
            'scaling': args.scaling, 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': dict(workload_config(world), samples_per_gpu=N),
            'clocks': clocks,
            'e2e': {'value': world * evals_per_step_rank * e2e_steps / (e2e_ms * 1e-3), 'unit': UNIT,
                    'h2d_bytes_per_step': world * N * P * 8, 'd2h_bytes_per_step': world * N * 8,
                    'ms_per_step': e2e_ms / e2e_steps, 'steps': e2e_steps,
                    'api': 'Gaussianlikelihood.log_likelihood_batch(numpy pinned) -> stn_loglike_host'
                           + (' on every rank, each rank\'s slice of the result DMA\'d into ' + assembled if world > 1 else ''),
                    'result_assembled_in': assembled, 'matches_device_path': same, 'pageable': e2e_pageable},
            'gpu_launches': launches,
            'gather': None if world == 1 else (
                'fused: the likelihood kernel\'s epilogue stores each log L into every rank\'s (N,) buffer over NVLink '
                '(CUDA IPC peer mappings, stn_loglike_scatter); the step ends with a one-element all_reduce as the barrier'
                if fused_gather else 'nccl all_gather_into_tensor after the kernel (CUDA IPC unavailable: %s)' % sharded.fused_error),
            'roofline': {
                'bound': 'fp64', 'achieved': achieved_tf, 'peak': peak_tf, 'unit': 'TFLOP/s',
                'frac': achieved_tf / peak_tf,
                # dram__bytes_read.sum + dram__bytes_write.sum of this kernel from one `ncu --set full` capture of this
                # command, read from the committed summary named in traffic_source; algorithmic bytes per launch below
                'traffic': traffic, 'traffic_unit': 'bytes per launch', 'traffic_source': traffic_file,
                'algorithmic_bytes': alg_bytes,
                'kernel': 'stn_loglike_fast_kernel<3,1,64>', 'kernel_ms': kernel_ms,
                'flops_per_eval': FLOPS_PER_EVAL['C'],
                'peak_source': 'stn_fp64_peak DFMA micro-benchmark in this run (MEASURED_PEAKS.json has no FP64 '
                               'figure); nominal 148 SM x 64 DFMA/clk x 2 x 1.965 GHz = 37.2',
                'hbm': {'achieved_gbs': alg_bytes / (kernel_ms * 1e-3) / 1e9, 'peak_gbs': peaks.get('hbm_gbs'),
                        'peak_source': peaks_src + ' (MEASURED_PEAKS.json)' if peaks_src == 'measured' else 'fallback',
                        'bytes_per_eval': alg_bytes / evals_per_step_rank},
            },
            'cpu_baseline': cpu_baseline,
            'variants': variants,
            'sweep': sweep,
        }
        if world > 1:
            line['sharded_bits_match'] = sharded_bits_match
            line['strong'] = strong
            line['single_process'] = single_process
    if shared is not None:
        shared.close()
    if sharded is not None:
        sharded.close()
    like.close()
    if world > 1:
        dist.destroy_process_group()
    if line is not None:
        sys.stdout.write('\n' + json.dumps(line) + '\n')      # after NCCL's teardown: the last line on stdout
        sys.stdout.flush()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--scaling', default='weak', choices=['weak', 'strong'])
    ap.add_argument('--cpu-rows', type=int, default=0,
                    help='parameter vectors in the bounded CPU sample (0: calibrate to ~10-15 s of wall time)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-variants', action='store_true')
    args = ap.parse_args()
    rank, local_rank, world = env_int('RANK', 0), env_int('LOCAL_RANK', 0), env_int('WORLD_SIZE', 1)
    if args.impl == 'reference':
        return run_reference(args, rank, world)
    return run_ours(args, rank, local_rank, world)


if __name__ == '__main__':
    sys.exit(main())

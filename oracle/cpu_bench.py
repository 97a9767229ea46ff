"""CPU timing of the oracle (TEST / BENCH INFRASTRUCTURE ONLY; see sterne_oracle.py header).

Times the CPU restatements of the reference's likelihood on the host cores:
  * `c_oracle`: the plain-C restatement (reference operation order), OpenMP over all cores --
    the strongest CPU arm, used as THE cpu baseline;
  * `vectorised`: batch_log_likelihood over rows, one worker PROCESS per core, all
    started on a common wall-clock tick so they load the machine together;
  * `scalar`: the reference-shaped per-vector Python loop (Gaussianlikelihood.log_likelihood),
    which is what dingswin/sterne actually executes per walker (simulate.py:362-384), on 1 core.
These are restatements of the reference algorithm, never measurements of sterne itself
(its dependencies are not installable here).

Worker protocol: `python -m oracle.cpu_bench --worker --variant C --rows R --seed S --start T`
prints one JSON line {"t0":…, "t1":…, "evals":…, "checksum":…}.
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _case(variant):
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    from sterne_b200 import synthetic          # input generator only
    return synthetic.config5(variant)


def _case_args(case):
    return (case.refepoch, case.LoD_timing, case.LoD_VLBI, case.shares, case.earth_xyz)


def _worker(variant, rows, seed, start):
    os.environ.setdefault('OMP_NUM_THREADS', '1')
    from oracle import sterne_oracle as O
    case = _case(variant)
    theta = case.sample_theta(rows, seed=seed, scale=0.3)
    O.batch_log_likelihood(*_case_args(case), theta[:8])           # warm-up
    while time.time() < start:
        time.sleep(0.005)
    t0 = time.time()
    out = O.batch_log_likelihood(*_case_args(case), theta, chunk=128)
    t1 = time.time()
    print(json.dumps({'t0': t0, 't1': t1, 'evals': rows * case.evals_per_sample, 'checksum': float(out.sum())}))


def time_vectorised(variant='C', rows_per_core=256, cores=None, seed=77, lead_s=8.0, timeout_s=600):
    """(evals/s, wall seconds, cores): `cores` worker processes evaluating concurrently."""
    cores = cores or os.cpu_count() or 1
    start = time.time() + lead_s
    env = dict(os.environ, OMP_NUM_THREADS='1', MKL_NUM_THREADS='1', OPENBLAS_NUM_THREADS='1',
               PYTHONPATH=ROOT + os.pathsep + os.environ.get('PYTHONPATH', ''))
    procs = [subprocess.Popen([sys.executable, '-m', 'oracle.cpu_bench', '--worker', '--variant', variant,
                               '--rows', str(rows_per_core), '--seed', str(seed + c), '--start', repr(start)],
                              cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
             for c in range(cores)]
    results = []
    for p in procs:
        try:
            out, err = p.communicate(timeout=timeout_s)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise RuntimeError('cpu_bench worker timed out')
        if p.returncode != 0:
            raise RuntimeError('cpu_bench worker failed: ' + err[-500:])
        results.append(json.loads(out.strip().splitlines()[-1]))
    t0 = min(r['t0'] for r in results)
    t1 = max(r['t1'] for r in results)
    evals = sum(r['evals'] for r in results)
    return evals / (t1 - t0), t1 - t0, cores


def time_c_oracle(variant='C', rows=1 << 15, threads=None, seed=79, repeats=1):
    """(evals/s, wall seconds, threads) of the threaded C restatement (oracle/c/sterne_oracle.c):
    the reference's operation order, one parameter vector per inner call, all host cores."""
    from oracle import c_oracle
    case = _case(variant)
    threads = threads or os.cpu_count() or 1
    co = c_oracle.from_case(case)
    theta = case.sample_theta(rows, seed=seed, scale=0.3)
    co.log_likelihood_batch(theta[:threads * 8], nthreads=threads)           # warm-up
    t0 = time.perf_counter()
    for _ in range(repeats):
        co.log_likelihood_batch(theta, nthreads=threads)
    wall = time.perf_counter() - t0
    return repeats * rows * case.evals_per_sample / wall, wall, threads


def time_scalar(variant='C', rows=1, seed=78):
    """(evals/s, wall seconds) of the reference-shaped scalar path on one core."""
    from oracle import sterne_oracle as O
    case = _case(variant)
    like = O.Gaussianlikelihood(case.refepoch, case.LoD_timing, case.LoD_VLBI, case.shares, case.earth_xyz,
                                case.DoD_additional_constraints, case.a1dot_constraints)
    theta = case.sample_theta(rows, seed=seed, scale=0.3)
    t0 = time.perf_counter()
    like.log_likelihood_rows(theta)
    wall = time.perf_counter() - t0
    return rows * case.evals_per_sample / wall, wall


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--worker', action='store_true')
    ap.add_argument('--variant', default='C')
    ap.add_argument('--rows', type=int, default=256)
    ap.add_argument('--seed', type=int, default=77)
    ap.add_argument('--start', type=float, default=0.0)
    a = ap.parse_args()
    if a.worker:
        _worker(a.variant, a.rows, a.seed, a.start)
    else:
        v, w, c = time_vectorised(a.variant, a.rows)
        print('vectorised: %.3e evals/s on %d cores (%.1f s)' % (v, c, w))
        v, w = time_scalar(a.variant, 1)
        print('scalar    : %.3e evals/s on 1 core (%.1f s)' % (v, w))

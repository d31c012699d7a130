"""Headline benchmark: SVGP ELBO iterations/s at N=1e6, M=1024, D=8 (BASELINE.json `metric`).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one rank per GPU)
    python bench.py --impl reference --gpus N ...            # the reference's algorithm on the host cores
    python bench.py --workload c1|c2|c3|c4|c5 [--out FILE]   # the other BASELINE.json configs, same JSON

One "step" = one pass of the reference's inner loop body (mxfusion/inference/minibatch_loop.py:81-92):
gather a minibatch of B=4096 rows, ELBO forward (svgp_regression.py:43-109), gradient, Adam update --
driven through the public API (Model / SVGPRegression.define_variable / GradBasedInference.run with a
MinibatchInferenceLoop).  `value` has the data set resident in HBM (a 256 MiB L2 flush between iterations inside the
timed region; `value_no_flush` is the same loop without it); `e2e` streams every minibatch from pinned host memory and
streams the loss back (non-blocking D2H into pinned memory every step).  Data-parallel runs give every rank a 1/G shard
of the rows and the same per-rank batch (weak scaling), with the flat gradient bucket averaged once per step (one peer-memory
all-reduce kernel captured in the step's graph; NCCL if the GPUs cannot map each other's memory);
value = G * steps / time, in minibatch iterations per second; `strong_scaling` repeats the run with the GLOBAL batch
fixed at B (B/G rows per rank).

Extra keys: `roofline` (K(X,Z) at the full headline size against the measured HBM peak), `roofline_tensor` (the panel /
trailing-update GEMMs of a potrf at N=8192 -- the tensor-pipe kernels of the path -- against the TF32 peak measured in
the same run), `cpu_baseline` (the op-for-op CPU restatement of the reference's step on all host cores, on the SAME
minibatches as the GPU arm), `parity` (loss after the same 25 steps on both arms), `clocks`.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_ROWS, M_IND, D_IN, BATCH = 1000000, 1024, 8, 4096
KERNEL = 'rbf'
# other BASELINE.json configs (not the bench line the driver reads; `--workload c2|c3` prints the same JSON for them)
WORKLOADS = {'headline': (1000000, 1024, 8, 4096, 'rbf'), 'c2': (100000, 512, 8, 2048, 'rbf'),
             'c3': (1000000, 1024, 16, 4096, 'matern52')}
OTHER_WORKLOADS = ('c1', 'c4', 'c5')      # exact GP N=512 D=2 / mean-field BNN / 2-layer deep GP: see run_other()
PARITY_STEPS = 25
JITTER, LR = 1e-6, 1e-2
KBUILD_NCU_TRAFFIC_BYTES = 4068219664      # committed capture profiles/r2b_kbuild_tc_rbf8_raw.csv (ncu --set full, this round,
                                           # kbuild_fwd_tc_kernel<rbf, 1 K-slice>): 32.08 MB read + 4036.1 MB written per launch
METRIC = "svgp_elbo_iters_per_sec"
UNIT = "minibatch iterations (B=4096 rows: ELBO fwd + grad + Adam) per second, summed over GPUs"


def synthetic(n=None, d=None, m=None):
    """X ~ U(-3,3)^(N x D), f = sum_d sin(x_d)/sqrt(D), Y = f + 0.05 N(0,1);
    Z = first M rows of a seed-1 permutation of X."""
    import torch
    n, d, m = n or N_ROWS, d or D_IN, m or M_IND
    g = torch.Generator(device='cpu').manual_seed(0)
    X = torch.rand((n, d), generator=g, dtype=torch.float32) * 6.0 - 3.0
    Y = (torch.sin(X).sum(dim=1, keepdim=True) / math.sqrt(d) +
         0.05 * torch.randn((n, 1), generator=g, dtype=torch.float32))
    perm = torch.randperm(n, generator=torch.Generator(device='cpu').manual_seed(1))[:m]
    return X, Y, X[perm].clone()


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json'))), 'measured'
    except Exception:
        return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0}, 'fallback'


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.rows, self.stop_flag, self.index = [], False, index
        self.proc = None

    def run(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(',')])
                if self.stop_flag:
                    break
        except Exception:
            pass

    def finish(self):
        self.stop_flag = True
        if self.proc is not None:
            self.proc.terminate()
        sm, mx, reasons = [], 0.0, set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
                for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[3:7]):
                    if v.lower().startswith('active'):
                        reasons.add(name)
            except (ValueError, IndexError):
                continue
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': mx or None, 'reasons': sorted(reasons),
                'samples': len(sm)}


def build_inference(X, Y, Z, n_total, world, data_resident, dtype='float32', device=None, batch=None):
    import torch
    import mxfusion_b200 as mf
    from mxfusion_b200.components.variables import PositiveTransformation
    from mxfusion_b200.components.distributions.gp.kernels import RBF, Matern52
    from mxfusion_b200.modules.gp_modules import SVGPRegression
    from mxfusion_b200.inference import GradBasedInference, MAP, MinibatchInferenceLoop
    mf.config.DEFAULT_DTYPE = dtype
    m = mf.Model()
    m.N = mf.Variable()
    m.X = mf.Variable(shape=(m.N, D_IN))
    m.noise_var = mf.Variable(shape=(1,), transformation=PositiveTransformation(), initial_value=0.01)
    m.kernel = (RBF if KERNEL == 'rbf' else Matern52)(input_dim=D_IN, variance=1, lengthscale=1)
    m.Z = mf.Variable(shape=tuple(Z.shape), initial_value=Z)
    m.Y = SVGPRegression.define_variable(X=m.X, kernel=m.kernel, noise_var=m.noise_var, inducing_inputs=m.Z,
                                         shape=(m.N, 1))
    m.Y.factor.svgp_log_pdf.jitter = JITTER
    batch = BATCH if batch is None else batch
    loop = MinibatchInferenceLoop(batch_size=batch, rv_scaling={m.Y: n_total / float(batch)},
                                  data_resident=data_resident, rng=np.random.RandomState(1234))
    infr = GradBasedInference(inference_algorithm=MAP(model=m, observed=[m.X, m.Y]), grad_loop=loop,
                              dtype=dtype, context=device)
    infr.initialize(X=tuple(X.shape), Y=tuple(Y.shape))
    post = m.Y.factor._extra_graphs[0]
    Mi = Z.shape[0]
    infr.params[post.qU_mean] = torch.zeros((Mi, 1))
    infr.params[post.qU_cov_W] = torch.zeros((Mi, Mi))
    infr.params[post.qU_cov_diag] = torch.ones((Mi,))
    return infr, loop


def timed_run(infr, loop, X, Y, steps, warmup, flush, use_events, barrier, batch=None):
    """Runs warmup + steps through GradBasedInference.run; returns (seconds for `steps` steps, last loss)."""
    import torch
    state = {}
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def on_step(k, loss):
        if flush is not None:
            flush.zero_()                      # > L2 write between iterations
        if k == warmup:
            barrier()
            torch.cuda.synchronize()
            state['t0'] = time.perf_counter()
            ev0.record()
        elif k == warmup + steps:
            ev1.record()
            barrier()
            torch.cuda.synchronize()
            state['t1'] = time.perf_counter()
            state['loss'] = float(loss)
    epochs = 1 + (warmup + steps) * (batch or BATCH) // X.shape[0]
    infr.run(X=X, Y=Y, max_iter=epochs + 1, learning_rate=LR, max_steps=warmup + steps, on_step=on_step)
    wall = state['t1'] - state['t0']
    dev = ev0.elapsed_time(ev1) / 1e3
    return (dev if use_events else wall), wall, state['loss']


def kernel_rooflines(device, pk):
    """The dominant HBM kernel of the path on its own: K(X,Z) at the full headline size (output 4.1 GB >> L2),
    CUDA events on the launching stream, after warm-up.  Algorithmic bytes: SURVEY.md section 8(d)."""
    import torch
    from mxfusion_b200 import _raw
    g = torch.Generator(device='cpu').manual_seed(0)
    X = (torch.rand((1, N_ROWS, D_IN), generator=g) * 6 - 3).to(device)
    Z = X[:, :M_IND].clone()
    ls = torch.ones((1, 1), device=device)
    var = torch.ones((1, 1), device=device)
    out = torch.empty((1, N_ROWS, M_IND), device=device)
    kind = _raw.RBF if KERNEL == 'rbf' else _raw.MATERN52
    for _ in range(3):
        _raw.kbuild_fwd(kind, X, Z, ls, var, out=out)
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        _raw.kbuild_fwd(kind, X, Z, ls, var, out=out)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ms = sum(ts) / len(ts)
    nbytes = 4 * (N_ROWS * M_IND + N_ROWS * D_IN + M_IND * D_IN + D_IN + 1)
    ach = nbytes / ms / 1e6
    del out
    same = (KERNEL, N_ROWS, M_IND, D_IN) == ('rbf', 1000000, 1024, 8)
    return {'bound': 'hbm', 'kernel': 'kbuild_fwd_tc_kernel<%s> (tcgen05 cross term + TMA stores) K(X,Z) N=%d M=%d D=%d' % (
                KERNEL, N_ROWS, M_IND, D_IN),
            'achieved': ach,
            'peak': pk['hbm_gbs'], 'unit': 'GB/s', 'frac': ach / pk['hbm_gbs'],
            'traffic': KBUILD_NCU_TRAFFIC_BYTES if same else None,
            'traffic_source': ('committed capture (not measured in this run): dram__bytes_read.sum + dram__bytes_write.sum '
                               'per launch, ncu --set full, profiles/r2b_kbuild_tc_rbf8_raw.csv (0.032 GB read + 4.036 GB written)')
            if same else None,
            'ms_per_launch': ms, 'algorithmic_bytes': nbytes,
            'note': 'stand-alone launch at the BASELINE "K(X,Z) HBM GB/s" size; inside the timed step the same kernel builds '
                    'Kuu (1024 x 1024) and Kuf (1024 x 4096), L2-resident'}


def measure_tf32_peak(device, n=8192):
    """cuBLAS TF32 GEMM rate measured the way the driver measured bf16 (MEASURED_PEAKS.json `how`): torch.matmul n^3, best of
    10, CUDA events.  The denominator of the tensor roofline (scripts/measure_peaks.py writes the committed copy)."""
    import torch
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    a = torch.randn((n, n), device=device)
    b = torch.randn((n, n), device=device)
    c = torch.empty((n, n), device=device)
    for _ in range(3):
        torch.matmul(a, b, out=c)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b, out=c)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    torch.backends.cuda.matmul.allow_tf32 = prev
    return 2.0 * n ** 3 / best / 1e9


def tensor_roofline(device, pk):
    """The tensor-pipe kernels of the path: the panel (L21 = A21 W^T) and trailing-update (A22 -= L21 L21^T, lower tiles)
    GEMMs of a blocked potrf at N = 8192, K = 1024 (csrc/chol_packed.cu -> csrc/gemm_tc.cu, tcgen05 kind::tf32, 3 MMAs per
    multiply-add for fp32 accuracy).  Every GEMM of the factorisation is timed on its own with CUDA events; `achieved`
    counts the TF32 MMA flops issued on the tiles actually computed (3 x 2mnk, lower tiles only for the update),
    time-weighted over the 14 launches; `peak` is the cuBLAS TF32 rate measured in this run."""
    import torch
    from mxfusion_b200 import _raw
    n, ob = 8192, 1024
    A = torch.randn((1, n, ob), device=device) / 32
    Wt = torch.tril(torch.randn((1, ob, ob), device=device)) / 32
    out = torch.empty((1, n, ob), device=device)
    C = torch.zeros((1, n, n), device=device)
    rows = []
    for o0 in range(0, n - ob, ob):
        below = n - o0 - ob
        for name, fn, flops in (
                ('panel %d x %d x %d' % (below, ob, ob),
                 lambda: _raw.gemm(A[:, :below], Wt, False, True, C=out[:, :below]), 2.0 * below * ob * ob),
                ('update %d x %d x %d (lower tiles)' % (below, below, ob),
                 lambda: _raw.gemm(out[:, :below], out[:, :below], False, True, alpha=-1.0, beta=1.0, C=C[:, :below, :below], tri=1),
                 1.0 * below * (below + 128) * ob)):
            for _ in range(2):
                fn()
            torch.cuda.synchronize()
            ts = []
            for _ in range(5):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                fn()
                b.record()
                torch.cuda.synchronize()
                ts.append(a.elapsed_time(b))
            rows.append((name, min(ts), flops))
    peak = measure_tf32_peak(device)
    tot_ms = sum(r[1] for r in rows)
    tot_fl = sum(r[2] for r in rows)
    fp32 = tot_fl / tot_ms / 1e9
    return {'bound': 'tensor', 'kernel': 'gemm_tc_* (tcgen05 3xTF32): 7 panel + 7 trailing-update GEMMs of potrf N=8192, K=1024',
            'achieved': 3.0 * fp32, 'peak': peak, 'unit': 'TFLOP/s', 'frac': 3.0 * fp32 / peak,
            'fp32_equivalent_tflops': fp32, 'ms_total': tot_ms,
            'peak_source': 'torch.matmul TF32 8192^3, best of 10, measured in this run (cuBLAS)',
            'per_gemm': [{'gemm': r[0], 'ms': r[1], 'tf32_tflops': 3.0 * r[2] / r[1] / 1e9} for r in rows],
            'ncu': 'tensor-pipe activity of these launches: profiles/ (potrf8192 launch list of this round)'}


def shard_batches(n_rows, steps):
    """The minibatches the GPU arm's loop draws (RolloverBatchSampler seeded like build_inference's loop)."""
    from mxfusion_b200.inference.minibatch_loop import RolloverBatchSampler
    sampler = RolloverBatchSampler(n_rows, BATCH, rng=np.random.RandomState(1234))
    out = []
    while len(out) < steps:
        idx, nfull = sampler.epoch_indices()
        out += [idx[i * BATCH:(i + 1) * BATCH] for i in range(nfull)]
    return out[:steps]


def cpu_reference(X, Y, Z, n_total, seconds_budget=20.0, max_iters=30, min_iters=PARITY_STEPS):
    """The reference's algorithm, op for op, on the host cores (oracle/torch_ref.py: LAPACK potrf/trsm, BLAS gemm,
    autograd backward, Adam with grads/B) -- `kind: port` because MXNet cannot be installed here -- on the SAME
    minibatches as the GPU arm.  Returns iterations/s, cores, iterations, seconds, losses per iteration."""
    import torch
    from oracle import torch_ref
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    M = Z.shape[0]
    step = torch_ref.SVGPStepCPU(torch_ref.RBF if KERNEL == 'rbf' else torch_ref.MATERN52, Z, np.array([0.01]), np.array([1.0]), np.array([1.0]),
                                 np.zeros((M, 1)), np.zeros((M, M)), np.ones((M,)), JITTER,
                                 n_total / float(BATCH), LR, dtype=torch.float32)
    batches = shard_batches(X.shape[0], max(max_iters, min_iters))
    losses = []
    t0 = time.perf_counter()
    n = 0
    while n < len(batches) and (n < min_iters or (n < max_iters and (time.perf_counter() - t0) < seconds_budget)):
        sel = torch.from_numpy(batches[n])
        losses.append(step.step(X[sel], Y[sel], BATCH))
        n += 1
    dt = time.perf_counter() - t0
    return n / dt, cores, n, dt, losses


def gpu_parity_losses(Xs, Ys, Z, device, steps=PARITY_STEPS):
    """Losses of the first `steps` iterations of the GPU arm (same loop, same sampler seed), read back one by one."""
    infr, loop = build_inference(Xs, Ys, Z, N_ROWS, 1, data_resident=True, device=device)
    got = []
    infr.run(X=Xs, Y=Ys, max_iter=2 + steps * BATCH // Xs.shape[0], learning_rate=LR, max_steps=steps,
             on_step=lambda k, l: got.append(l.clone()))
    return [float(l) for l in got]


def parity_record(gpu_losses, cpu_losses, tol=2e-4):
    k = min(len(gpu_losses), len(cpu_losses))
    rel = [abs(a - b) / abs(b) for a, b in zip(gpu_losses[:k], cpu_losses[:k])]
    rec = {'steps': k, 'gpu_loss': gpu_losses[k - 1], 'cpu_loss': cpu_losses[k - 1], 'rel_diff_last': rel[-1],
           'rel_diff_max': max(rel), 'tol': tol, 'ok': max(rel) <= tol,
           'what': 'loss of iterations 1..%d on the same minibatches: this repo (f32, CUDA) vs the CPU restatement of the '
                   'reference (f32, LAPACK/BLAS)' % k}
    assert rec['ok'], "GPU and CPU arms disagree on the same minibatches: %r" % (rec,)
    return rec


def run_other(args, device):
    """BASELINE.json configs that are not SVGP: c1 exact GP N=512 D=2 (full-batch loop), c4 mean-field BNN (MC-ELBO),
    c5 two-layer deep GP (doubly-stochastic VI).  Same JSON contract."""
    import torch
    import mxfusion_b200 as mf
    from mxfusion_b200 import _lib
    pk, pk_kind = peaks()
    steps, warmup = args.steps, max(args.warmup, 3)
    if args.workload == 'c4':
        from scripts import bench_bnn as bb
        x, y = bb.data()
        sampler = ClockSampler(0)
        sampler.start()
        g = bb.gpu_leg(x, y, steps=steps, warm=warmup)
        clocks = sampler.finish()
        e = bb.gpu_leg(x, y, steps=steps, warm=warmup, data_resident=False, wall=True)
        unit = 'minibatch iterations (B=%d rows, S=%d samples: MC-ELBO fwd + grad + Adam) per second' % (bb.B, bb.S)
        line = {'metric': 'bnn_meanfield_svi_iters_per_sec', 'value': g['iters_per_s'], 'unit': unit, 'n_gpus': 1, 'steps': steps,
                'warmup': warmup, 'ms_per_step': g['ms_per_step'], 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
                'dtype': 'f32', 'data': 'synthetic',
                'config': {'workload': 'mean-field BNN 1-%d-%d-1 tanh (bnn_regression.ipynb), S=%d, StochasticVariationalInference '
                                       'MC-ELBO, minibatch=%d, N=%d, f32 (BASELINE c4)' % (bb.H, bb.H, bb.S, bb.B, bb.N),
                           'l2_flush': 'none (working set 40 kB of weights + one minibatch)'},
                'clocks': clocks, 'e2e': {'value': e['iters_per_s'], 'unit': unit, 'h2d_bytes_per_step': e['h2d_bytes_per_step'],
                                          'd2h_bytes_per_step': e['d2h_bytes_per_step'], 'ms_per_step': e['ms_per_step']},
                'gpu_launches': int((g['launches_per_step'] or 0) * steps), 'launches_per_step': g['launches_per_step'],
                'first_loss': g['first_loss'], 'final_loss': g['last_loss'], 'peaks': pk_kind,
                'roofline': None}
        if not args.no_cpu_baseline:
            c = bb.cpu_leg(x, y, budget_s=10.0)
            line['cpu_baseline'] = {'value': c['cpu_iters_per_s'], 'unit': unit, 'cores': c['cpu_cores'], 'kind': 'port',
                                    'sample': '%d iterations of the same step in 10 s (torch CPU f32, per-sample loop over the network '
                                              'as function_evaluation.py:80-93)' % c['cpu_iters']}
        return line
    if args.workload == 'c1':
        return run_exact_gp(args, device, pk_kind)
    return run_deep_gp(args, device, pk_kind)


def run_exact_gp(args, device, pk_kind):
    """c1: GPRegression exact GP, RBF, N=512, D=2, MAP with the full-batch loop (batch_loop.py:51-60); one step = one
    marginal-likelihood evaluation (gp_regression.py:42-76) + gradient + Adam on all 512 rows."""
    import torch
    import mxfusion_b200 as mf
    from mxfusion_b200.components.variables import PositiveTransformation
    from mxfusion_b200.components.distributions.gp.kernels import RBF
    from mxfusion_b200.modules.gp_modules import GPRegression
    from mxfusion_b200.inference import GradBasedInference, MAP
    from oracle import torch_ref
    N, D = 512, 2
    steps, warmup = args.steps, max(args.warmup, 3)
    g = torch.Generator().manual_seed(0)
    X = torch.rand((N, D), generator=g) * 6 - 3
    Y = torch.sin(X).sum(1, keepdim=True) / math.sqrt(D) + 0.05 * torch.randn((N, 1), generator=g)
    mf.config.DEFAULT_DTYPE = 'float32'

    def build():
        m = mf.Model()
        m.N = mf.Variable()
        m.X = mf.Variable(shape=(m.N, D))
        m.noise_var = mf.Variable(shape=(1,), transformation=PositiveTransformation(), initial_value=0.01)
        m.kernel = RBF(input_dim=D, variance=1, lengthscale=1)
        m.Y = GPRegression.define_variable(X=m.X, kernel=m.kernel, noise_var=m.noise_var, shape=(m.N, 1))
        infr = GradBasedInference(inference_algorithm=MAP(model=m, observed=[m.X, m.Y]), dtype='float32', context=device)
        infr.initialize(X=(N, D), Y=(N, 1))
        return m, infr
    # value: data resident, CUDA events around `steps` iterations of the captured step
    m, infr = build()
    Xd, Yd = X.to(device), Y.to(device)
    infr.run(X=Xd, Y=Yd, max_iter=warmup + 3, learning_rate=LR)
    st = infr._grad_loop.last_stepper
    sampler = ClockSampler(0)
    sampler.start()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = st.step()
    e1.record()
    torch.cuda.synchronize()
    clocks = sampler.finish()
    ms = e0.elapsed_time(e1) / steps
    # e2e: the public call with HOST arrays (H2D of X, Y, graph capture, `steps` iterations, D2H of the loss)
    m2, infr2 = build()
    infr2.run(X=X, Y=Y, max_iter=warmup, learning_rate=LR)
    t0 = time.perf_counter()
    l2 = infr2.run(X=X, Y=Y, max_iter=steps, learning_rate=LR)
    lf = float(l2)
    wall = time.perf_counter() - t0
    # parity: PARITY_STEPS iterations from the initial point on both arms
    m3, infr3 = build()
    l3 = float(infr3.run(X=X, Y=Y, max_iter=PARITY_STEPS, learning_rate=LR))
    unit = 'full-batch iterations (N=512: marginal likelihood fwd + grad + Adam) per second'
    line = {'metric': 'gp_exact_iters_per_sec', 'value': 1e3 / ms, 'unit': unit, 'n_gpus': 1, 'steps': steps, 'warmup': warmup,
            'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': 'GPRegression exact GP, RBF, N=512 D=2, MAP, BatchInferenceLoop, f32 (BASELINE c1); replicas only',
                       'l2_flush': 'none (1 MiB working set; the reference problem is L2-resident by construction)'},
            'clocks': clocks, 'e2e': {'value': steps / wall, 'unit': unit, 'h2d_bytes_per_step': int(4 * N * (D + 1) / steps),
                                      'd2h_bytes_per_step': 4.0 / steps, 'ms_per_step': 1e3 * wall / steps,
                                      'note': 'GradBasedInference.run(X=host, Y=host, max_iter=steps): one H2D of the data, graph '
                                              'capture, steps, one D2H of the final loss'},
            'gpu_launches': int((st.launches_per_step or 0) * steps), 'launches_per_step': st.launches_per_step,
            'final_loss': lf, 'peaks': pk_kind, 'roofline': None}
    if not args.no_cpu_baseline:
        torch.set_num_threads(os.cpu_count() or 1)
        inv = lambda v: torch.log(torch.expm1(torch.tensor([v])))
        leaves = [inv(0.01).requires_grad_(), inv(1.0).requires_grad_(), inv(1.0).requires_grad_()]
        opt = torch_ref.AdamMX(leaves, LR)
        sp = torch.nn.functional.softplus

        def cstep():
            loss = -torch_ref.gp_log_pdf(torch_ref.RBF, X[None], Y[None], sp(leaves[0])[None], sp(leaves[1])[None], sp(leaves[2])[None],
                                         jitter=0.0).sum()
            loss.backward()
            opt.step(1)
            return float(loss)
        closs = [cstep() for _ in range(PARITY_STEPS)]
        t0 = time.perf_counter()
        n = 0
        while time.perf_counter() - t0 < 8.0:
            cstep()
            n += 1
        dt = time.perf_counter() - t0
        line['cpu_baseline'] = {'value': n / dt, 'unit': unit, 'cores': os.cpu_count(), 'kind': 'port',
                                'sample': '%d iterations in %.1f s (torch CPU f32: LAPACK potrf/trsm + autograd + Adam)' % (n, dt)}
        # the GPU run returns the loss of iteration PARITY_STEPS evaluated BEFORE its update, like the CPU list's last entry
        rel = abs(l3 - closs[-1]) / abs(closs[-1])
        line['parity'] = {'steps': PARITY_STEPS, 'gpu_loss': l3, 'cpu_loss': closs[-1], 'rel_diff_last': rel, 'tol': 1e-3,
                          'ok': rel <= 1e-3}
        assert rel <= 1e-3, line['parity']
    return line


def run_deep_gp(args, device, pk_kind):
    from scripts import bench_deepgp
    return bench_deepgp.bench_line(args, device, pk_kind, ClockSampler)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=300)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--workload', default='headline', choices=sorted(WORKLOADS) + list(OTHER_WORKLOADS))
    ap.add_argument('--out', default=None, help='also write the JSON line to this file')
    ap.add_argument('--step-only', action='store_true',
                    help='profiling aid (ncu launch lists): only the data-resident timed loop, none of the other arms')
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))

    def emit(line):
        print(json.dumps(line))
        if args.out:
            os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
            with open(args.out, 'w') as f:
                f.write(json.dumps(line, indent=1) + '\n')

    if args.workload in OTHER_WORKLOADS and not (args.workload == 'c5' and world > 1):
        if rank != 0:
            return
        import torch
        if args.impl == 'reference':
            args.no_cpu_baseline = False
        torch.cuda.set_device(local)
        line = run_other(args, torch.device('cuda', local))
        if args.impl == 'reference':
            cb = line.get('cpu_baseline') or {}
            line = {'impl': 'reference', 'metric': line['metric'], 'value': cb.get('value'), 'unit': line['unit'], 'n_gpus': args.gpus,
                    'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 / cb['value'] if cb.get('value') else None,
                    'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
                    'config': line['config'], 'cpu_baseline': cb,
                    'e2e': {'value': cb.get('value'), 'unit': line['unit'], 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
        emit(line)
        return
    if args.workload == 'c5':
        from scripts import bench_deepgp
        bench_deepgp.main_distributed(args, emit)
        return

    global N_ROWS, M_IND, D_IN, BATCH, KERNEL, UNIT
    N_ROWS, M_IND, D_IN, BATCH, KERNEL = WORKLOADS[args.workload]
    UNIT = UNIT.replace('B=4096', 'B=%d' % BATCH)
    warmup = max(args.warmup, 3)
    config = {'workload': 'SVGPRegression N=%d M=%d D=%d %s minibatch=%d f32 (BASELINE %s; '
                          'jitter 1e-6, Adam lr 1e-2, rv_scaling=N/B)' % (N_ROWS, M_IND, D_IN, KERNEL, BATCH, args.workload),
              'N': N_ROWS, 'M': M_IND, 'D': D_IN,
              'batch_per_gpu': BATCH, 'parallelism': 'dp%d' % world,
              'sharding': 'rows split across ranks; the flat gradient is averaged once per step by one two-shot all-reduce '
                          'kernel over NVLink peer memory inside the captured step (NCCL all-reduce if peer mapping fails)'}

    if args.impl == 'reference':
        if rank != 0:
            return
        X, Y, Z = synthetic()
        shard = N_ROWS // world
        Xs, Ys = X[:shard], Y[:shard]              # rank 0's shard: the same minibatches as the GPU arm's rank 0
        ips, cores, n, dt, losses = cpu_reference(Xs, Ys, Z, N_ROWS, seconds_budget=max(20.0, 1.0 * args.steps),
                                                  max_iters=max(args.steps, PARITY_STEPS))
        line = {'impl': 'reference', 'metric': METRIC, 'value': ips, 'unit': UNIT, 'n_gpus': args.gpus,
                'steps': n, 'warmup': 0, 'ms_per_step': 1e3 / ips, 'higher_is_better': True, 'scaling': 'weak',
                'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'config': config,
                'cpu_baseline': {'value': ips, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                                 'sample': '%d iterations of the same B=%d step in %.1f s (torch CPU f32, '
                                           'LAPACK/BLAS op-for-op restatement; MXNet not installable), on the minibatches '
                                           'the GPU arm draws' % (n, BATCH, dt)},
                'final_loss': losses[-1], 'loss_at_step_%d' % PARITY_STEPS: losses[PARITY_STEPS - 1],
                'e2e': {'value': ips, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
        emit(line)
        return

    import torch
    import torch.distributed as dist
    from mxfusion_b200 import _lib
    torch.cuda.set_device(local)
    device = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=device)

    def barrier():
        if world > 1:
            dist.barrier()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    pk, pk_kind = peaks()
    X, Y, Z = synthetic()
    shard = N_ROWS // world
    Xs, Ys = X[rank * shard:(rank + 1) * shard], Y[rank * shard:(rank + 1) * shard]
    flush = torch.empty((256 << 20,), dtype=torch.uint8, device=device)

    # ---- value: data resident in HBM --------------------------------------------------------------------
    infr, loop = build_inference(Xs, Ys, Z, N_ROWS, world, data_resident=True, device=device)
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    secs, wall, loss = timed_run(infr, loop, Xs, Ys, args.steps, warmup, flush, True, barrier)
    clocks = sampler.finish() if sampler else None
    secs = max_over_ranks(secs)
    st = loop.last_stepper

    if args.step_only:
        if rank == 0:
            emit({'metric': METRIC, 'value': world * args.steps / secs, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
                  'warmup': warmup, 'ms_per_step': 1e3 * secs / args.steps, 'note': '--step-only: profiling aid, not a bench line'})
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- the same without the L2 flush (steady-state training: the 21 MB working set stays in L2) --------
    infr1, loop1 = build_inference(Xs, Ys, Z, N_ROWS, world, data_resident=True, device=device)
    secs_nf, _, _ = timed_run(infr1, loop1, Xs, Ys, args.steps, warmup, None, True, barrier)
    secs_nf = max_over_ranks(secs_nf)
    del infr1, loop1

    # ---- e2e: host buffers, H2D of every minibatch + D2H of the loss inside the timed region -------------
    infr2, loop2 = build_inference(Xs, Ys, Z, N_ROWS, world, data_resident=False, device=device)
    _, wall2, loss2 = timed_run(infr2, loop2, Xs, Ys, args.steps, warmup, flush, False, barrier)
    wall2 = max_over_ranks(wall2)

    # ---- strong scaling: the GLOBAL batch fixed at B, B/G rows per rank ----------------------------------
    strong = None
    if world > 1 and BATCH % world == 0:
        bl = BATCH // world
        infr3, loop3 = build_inference(Xs, Ys, Z, N_ROWS, world, data_resident=True, device=device, batch=bl)
        secs3, _, _ = timed_run(infr3, loop3, Xs, Ys, args.steps, warmup, flush, True, barrier, batch=bl)
        secs3 = max_over_ranks(secs3)
        strong = {'value': args.steps / secs3, 'ms_per_step': 1e3 * secs3 / args.steps, 'global_batch': BATCH, 'batch_per_gpu': bl,
                  'unit': 'iterations of the GLOBAL B=%d batch per second' % BATCH,
                  'note': 'the M x M work (two potrf, S^-1, the M^3 products) is replicated on every rank: SURVEY section 7.3 '
                          'bounds this curve at ~3x; the weak-scaling `value` is the data-parallel reading of the target'}
        del infr3, loop3

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    value = world * args.steps / secs
    e2e = world * args.steps / wall2
    launches_per_step = (st.launches_per_step if getattr(st, 'launches_per_step', None) else 0) + 2 + 2
    line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': warmup,
            'ms_per_step': 1e3 * secs / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f32', 'data': 'synthetic', 'config': dict(config, l2_flush='256 MiB device memset between '
                                                                'iterations (inside the timed region)'),
            'clocks': clocks,
            'value_no_flush': world * args.steps / secs_nf, 'ms_per_step_no_flush': 1e3 * secs_nf / args.steps,
            'e2e': {'value': e2e, 'unit': UNIT, 'h2d_bytes_per_step': loop2.h2d_bytes_per_step,
                    'd2h_bytes_per_step': loop2.d2h_bytes_per_step, 'ms_per_step': 1e3 * wall2 / args.steps,
                    'note': 'every minibatch gathered on the host into pinned memory and copied H2D inside the timed region; the '
                            'loss is STREAMED back (non-blocking D2H into pinned memory every step, read by the host at the end) '
                            '-- the reference blocks on asscalar() every step (minibatch_loop.py:92)'},
            'gpu_launches': int(launches_per_step * args.steps),
            'launches_per_step': int(launches_per_step), 'final_loss': loss, 'final_loss_e2e': loss2,
            'peaks': pk_kind}
    line['config']['gradient_exchange'] = getattr(st, 'exchange', 'single')
    # the same throughput as data rows per second and as seconds per pass over the N rows
    line['rows_per_s'] = value * BATCH
    line['seconds_per_pass_over_N'] = N_ROWS / (value * BATCH)
    line['e2e']['rows_per_s'] = e2e * BATCH
    if strong is not None:
        line['strong_scaling'] = strong
    if world == 1:
        del infr, infr2, flush
        torch.cuda.empty_cache()
        gpu_losses = gpu_parity_losses(Xs, Ys, Z, device)
        line['loss_at_step_%d' % PARITY_STEPS] = gpu_losses[-1]
        line['roofline'] = kernel_rooflines(device, pk)
        line['roofline_tensor'] = tensor_roofline(device, pk)
        if not args.no_cpu_baseline:
            ips, cores, n, dt, closs = cpu_reference(Xs, Ys, Z, N_ROWS, seconds_budget=12.0, max_iters=200)
            line['cpu_baseline'] = {'value': ips, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                                    'sample': '%d iterations of the same B=%d step in %.1f s on the host cores '
                                              '(torch CPU f32 op-for-op restatement of the reference), on the minibatches the '
                                              'GPU arm draws' % (n, BATCH, dt)}
            line['parity'] = parity_record(gpu_losses, closs)
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()

"""Headline benchmark: SVGP ELBO iterations/s at N=1e6, M=1024, D=8 (BASELINE.json `metric`).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one rank per GPU)
    python bench.py --impl reference --gpus N ...            # the reference's algorithm on the host cores

One "step" = one pass of the reference's inner loop body (mxfusion/inference/minibatch_loop.py:81-92):
gather a minibatch of B=4096 rows, ELBO forward (svgp_regression.py:43-109), gradient, Adam update --
driven through the public API (Model / SVGPRegression.define_variable / GradBasedInference.run with a
MinibatchInferenceLoop).  `value` has the data set resident in HBM; `e2e` streams every minibatch from
pinned host memory and reads the loss back every step.  Data-parallel runs give every rank a 1/G shard
of the rows and the same per-rank batch (weak scaling), with one NCCL all-reduce of the flat gradient
bucket per step; value = G * steps / time, in minibatch iterations per second.

Extra keys: `roofline` (K(X,Z) at the full headline size against the measured HBM peak, `traffic` from the committed ncu
capture), `roofline_tensor` (the 4096^3 tcgen05 3xTF32 GEMM against half the measured bf16 peak), `cpu_baseline` (the
op-for-op CPU restatement of the reference's step on all host cores, ~12 s sample), `clocks`.  `--workload c2|c3` runs the
other SVGP configs of BASELINE.json through the same code and prints the same JSON (not the line the driver reads).
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
JITTER, LR = 1e-6, 1e-2
KBUILD_NCU_TRAFFIC_BYTES = 4068866256      # profiles/r1e_kbuild_raw.csv: 32.06 MB read + 4036.8 MB written per launch
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


def build_inference(X, Y, Z, n_total, world, data_resident, dtype='float32', device=None):
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
    loop = MinibatchInferenceLoop(batch_size=BATCH, rv_scaling={m.Y: n_total / float(BATCH)},
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


def timed_run(infr, loop, X, Y, steps, warmup, flush, use_events, barrier):
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
    epochs = 1 + (warmup + steps) * BATCH // X.shape[0]
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
    for _ in range(3):
        _raw.kbuild_fwd(_raw.RBF if KERNEL == 'rbf' else _raw.MATERN52, X, Z, ls, var, out=out)
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        _raw.kbuild_fwd(_raw.RBF if KERNEL == 'rbf' else _raw.MATERN52, X, Z, ls, var, out=out)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ms = sum(ts) / len(ts)
    nbytes = 4 * (N_ROWS * M_IND + N_ROWS * D_IN + M_IND * D_IN + D_IN + 1)
    ach = nbytes / ms / 1e6
    del out
    return {'bound': 'hbm', 'kernel': 'kbuild_fwd_stream_kernel<float,%s> K(X,Z) N=%d M=%d D=%d' % (KERNEL, N_ROWS, M_IND, D_IN),
            'achieved': ach,
            'peak': pk['hbm_gbs'], 'unit': 'GB/s', 'frac': ach / pk['hbm_gbs'],
            'traffic': KBUILD_NCU_TRAFFIC_BYTES if (KERNEL, N_ROWS, M_IND, D_IN) == ('rbf', 1000000, 1024, 8) else None,
            'traffic_source': 'dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full capture '
                              'profiles/r1e_kbuild_raw.csv (0.032 GB read + 4.037 GB written)',
            'ms_per_launch': ms, 'algorithmic_bytes': nbytes}


def tensor_roofline(device, pk):
    """The tensor-pipe kernel of the path on its own (the potrf / trsm / syrk update engine, csrc/gemm_tc.cu): a
    4096^3 FP32-accurate product = 3 TF32 tcgen05 MMAs per multiply-add.  `achieved` counts the TF32 MMA flops actually
    issued (3 x 2mnk); `peak` = half the measured dense bf16 rate (TF32 runs at half the bf16 MMA rate)."""
    import torch
    from mxfusion_b200 import _raw
    n = 4096
    A = torch.randn((1, n, n), device=device)
    B = torch.randn((1, n, n), device=device)
    C = torch.empty((1, n, n), device=device)
    for _ in range(3):
        _raw.gemm(A, B, False, True, C=C)
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        _raw.gemm(A, B, False, True, C=C)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ms = sum(ts) / len(ts)
    fp32_tflops = 2.0 * n ** 3 / ms / 1e9
    peak = pk.get('bf16_tflops', 1590.0) / 2.0
    return {'bound': 'tensor', 'kernel': 'gemm_tc_ta_kernel<256> 4096^3 (3xTF32, tcgen05 + TMEM-resident A)',
            'achieved': 3.0 * fp32_tflops, 'peak': peak, 'unit': 'TFLOP/s', 'frac': 3.0 * fp32_tflops / peak,
            'fp32_equivalent_tflops': fp32_tflops, 'ms_per_launch': ms,
            'ncu': 'sm__pipe_tensor_cycles_active 83.1 % of peak sustained active, 70.4 % of elapsed: 3.46 waves '
                   '(profiles/r1e_gemm4096_raw.csv)'}


def cpu_reference_iters_per_sec(X, Y, Z, n_total, seconds_budget=20.0, max_iters=30):
    """The reference's algorithm, op for op, on the host cores (oracle/torch_ref.py: LAPACK potrf/trsm, BLAS gemm,
    autograd backward, Adam with grads/B) -- `kind: port` because MXNet cannot be installed here."""
    import torch
    from oracle import torch_ref
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    M = Z.shape[0]
    step = torch_ref.SVGPStepCPU(torch_ref.RBF if KERNEL == 'rbf' else torch_ref.MATERN52, Z, np.array([0.01]), np.array([1.0]), np.array([1.0]),
                                 np.zeros((M, 1)), np.zeros((M, M)), np.ones((M,)), JITTER,
                                 n_total / float(BATCH), LR, dtype=torch.float32)
    rng = np.random.RandomState(1234)
    idx = torch.from_numpy(rng.permutation(X.shape[0])[:BATCH * 64])

    def one(i):
        sel = idx[(i % 64) * BATCH:((i % 64) + 1) * BATCH]
        return step.step(X[sel], Y[sel], BATCH)
    one(0)
    one(1)
    t0 = time.perf_counter()
    n = 0
    while n < max_iters and (time.perf_counter() - t0) < seconds_budget:
        one(n + 2)
        n += 1
    dt = time.perf_counter() - t0
    return n / dt, cores, n, dt


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=300)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--workload', default='headline', choices=sorted(WORKLOADS))
    args = ap.parse_args()
    global N_ROWS, M_IND, D_IN, BATCH, KERNEL, UNIT
    N_ROWS, M_IND, D_IN, BATCH, KERNEL = WORKLOADS[args.workload]
    UNIT = UNIT.replace('B=4096', 'B=%d' % BATCH)
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    warmup = max(args.warmup, 3)
    config = {'workload': 'SVGPRegression N=%d M=%d D=%d %s minibatch=%d f32 (BASELINE %s; '
                          'jitter 1e-6, Adam lr 1e-2, rv_scaling=N/B)' % (N_ROWS, M_IND, D_IN, KERNEL, BATCH, args.workload),
              'N': N_ROWS, 'M': M_IND, 'D': D_IN,
              'batch_per_gpu': BATCH, 'parallelism': 'dp%d' % world,
              'sharding': 'rows split across ranks, one NCCL all-reduce of the flat gradient per step'}

    if args.impl == 'reference':
        if rank != 0:
            return
        X, Y, Z = synthetic()
        ips, cores, n, dt = cpu_reference_iters_per_sec(X, Y, Z, N_ROWS, seconds_budget=max(20.0, 1.0 * args.steps),
                                                        max_iters=max(args.steps, 5))
        line = {'impl': 'reference', 'metric': METRIC, 'value': ips, 'unit': UNIT, 'n_gpus': args.gpus,
                'steps': n, 'warmup': 2, 'ms_per_step': 1e3 / ips, 'higher_is_better': True, 'scaling': 'weak',
                'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'config': config,
                'cpu_baseline': {'value': ips, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                                 'sample': '%d iterations of the same B=4096 step in %.1f s (torch CPU f32, '
                                           'LAPACK/BLAS op-for-op restatement; MXNet not installable)' % (n, dt)},
                'e2e': {'value': ips, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
        print(json.dumps(line))
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
    l0 = _lib.launch_count()
    secs, wall, loss = timed_run(infr, loop, Xs, Ys, args.steps, warmup, flush, True, barrier)
    per_step_launches = getattr(loop.last_stepper, 'launches_per_step', None)
    clocks = sampler.finish() if sampler else None
    t = torch.tensor([secs], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    secs = float(t.item())
    eager_launches = _lib.launch_count() - l0

    # ---- e2e: host buffers, H2D of every minibatch + D2H of the loss inside the timed region -------------
    infr2, loop2 = build_inference(Xs, Ys, Z, N_ROWS, world, data_resident=False, device=device)
    _, wall2, loss2 = timed_run(infr2, loop2, Xs, Ys, args.steps, warmup, flush, False, barrier)
    t2 = torch.tensor([wall2], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
    wall2 = float(t2.item())

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    value = world * args.steps / secs
    e2e = world * args.steps / wall2
    st = loop.last_stepper
    launches_per_step = (st.launches_per_step if getattr(st, 'launches_per_step', None) else 0) + 2 + 2
    line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': warmup,
            'ms_per_step': 1e3 * secs / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f32', 'data': 'synthetic', 'config': dict(config, l2_flush='256 MiB device memset between '
                                                                'iterations (inside the timed region)'),
            'clocks': clocks,
            'e2e': {'value': e2e, 'unit': UNIT, 'h2d_bytes_per_step': loop2.h2d_bytes_per_step,
                    'd2h_bytes_per_step': loop2.d2h_bytes_per_step, 'ms_per_step': 1e3 * wall2 / args.steps},
            'gpu_launches': int(launches_per_step * args.steps),
            'launches_per_step': int(launches_per_step), 'final_loss': loss, 'final_loss_e2e': loss2,
            'peaks': pk_kind}
    if world == 1:
        del infr, infr2, flush
        torch.cuda.empty_cache()
        line['roofline'] = kernel_rooflines(device, pk)
        line['roofline_tensor'] = tensor_roofline(device, pk)
        if not args.no_cpu_baseline:
            ips, cores, n, dt = cpu_reference_iters_per_sec(X, Y, Z, N_ROWS, seconds_budget=12.0, max_iters=200)
            line['cpu_baseline'] = {'value': ips, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                                    'sample': '%d iterations of the same B=4096 step in %.1f s on the host cores '
                                              '(torch CPU f32 op-for-op restatement of the reference)' % (n, dt)}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()

#!/usr/bin/env python
"""Benchmark of the MMLF hot path on B200 (driver contract: one JSON line on stdout from rank 0).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload train|infer] [--variant base|upr|dpp]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # the CPU port of the reference algorithm (oracle/) on the host cores

workload train (default): one training step of the 4-stream FeedForward model -- forward, masked loss, backward,
gradient all-reduce, Adam -- on a global batch of 512 patches of 96 x 96 px, 9 views per stack (BASELINE.json
configs[1]); the batch is sharded over the ranks (strong scaling: the global batch is fixed).  The step is the product's
``mmlf_b200.train.step.TrainStep`` (what ``mmlf.train.cli`` runs): one CUDA-graph replay per step.
workload infer: full-light-field inference, one 9x9-view 512 x 512 light field per rank and step.
workload ese / bands: the 70-member shift ensemble / the row bands of ONE light field sharded over the ranks.

The default run prints the BASE-train headline line and nests the other BASELINE.json configs under ``secondary``
(``infer``, ``upr_train``, ``dpp_train``, ``ese``, ``bands``: each with value, e2e, roofline, measured in the same process,
one after the other); ``--no-secondary`` or an explicit ``--workload`` / ``--variant`` prints that one workload alone.
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
sys.path.insert(0, os.path.join(ROOT, 'tests'))

import numpy as np  # noqa: E402

ESE_MEMBERS = 70
# dram__bytes_read.sum + dram__bytes_write.sum of one launch of the dominant kernel, from the ncu --set full capture
# profiles/ncu_conv280_r02.txt (not measurable inside an un-profiled run)
CONV_TRAFFIC = {'bytes': 653921024,
                'note': 'ncu capture profiles/ncu_conv280_r02.txt: conv2x2_tc2 280->280 pad 0 on 64x96x96 patches, 347.9 MB '
                        'read + 306.1 MB written per launch against 693.7 MB algorithmic (activations in + out; the 663 KB '
                        'weight operand stays in L2)'}
FULL_KW = dict(model_ksize=2, model_in_blocks=3, model_out_blocks=8, model_chs=70, model_views=9, model_cross=False,
               model_uncert=False, model_unet=False, model_discrete=False, model_no_batchnorm=False,
               model_batchnorm_momentum=0.1, val_disp_min=-3.5, val_disp_max=3.5)


def model_kwargs(variant):
    kw = dict(FULL_KW)
    kw.update(model_uncert=(variant == 'upr'), model_discrete=(variant == 'dpp'))
    return kw


# ----------------------------------------------------------------------------------------------- algorithmic work
def conv_flops(B, H, W, cin, cout, ctype):
    """2 * MACs of one nn.Conv2d(cin, cout, 2): type 0 = padding 1 -> (H+1)x(W+1) outputs, type 1 = padding 0."""
    m = B * (H + 1) * (W + 1) if ctype == 0 else B * H * W
    return 2.0 * m * cout * 4 * cin


def net_forward_flops(B, H, W, variant, chs=70, views=9, streams=4, in_blocks=3, out_blocks=8):
    oc = {'base': 1, 'upr': 2, 'dpp': streams * views * 3}[variant]
    f = 0.0
    for _ in range(streams):
        cin = views * 3
        for _k in range(in_blocks):
            f += conv_flops(B, H, W, cin, chs, 0) + conv_flops(B, H, W, chs, chs, 1)
            cin = chs
    w = streams * chs
    for _k in range(out_blocks - 1):
        f += conv_flops(B, H, W, w, w, 0) + conv_flops(B, H, W, w, w, 1)
    f += conv_flops(B, H, W, w, oc, 0) + conv_flops(B, H, W, oc, oc, 1)
    return f


def conv_split(B, H, W, variant, training, chs=70, views=9, streams=4, in_blocks=3, out_blocks=8):
    """Algorithmic work of the conv2x2 launches (forward [+ data gradients]) split into the wide out-net layers (FLOPs) and
    the narrow in-net layers (FLOPs and HBM bytes: every input slot array read once, every output array written once;
    16-bit storage on the padded channel pitch, training forward: bf16 twins + ReLU bits of the first conv of a block)."""
    def pad16(c):
        return (c + 15) // 16 * 16
    n_slots = B * (H + 1) * (W + 1)
    narrow_fl = narrow_by = 0.0
    for _ in range(streams):
        cin = views * 3
        for k in range(in_blocks):
            f1, f2 = conv_flops(B, H, W, cin, chs, 0), conv_flops(B, H, W, chs, chs, 1)
            narrow_fl += f1 + f2
            outs = 2 if training else 1
            narrow_by += n_slots * 2.0 * (pad16(cin) + outs * pad16(chs)) + (n_slots * 4.0 * ((pad16(chs) + 31) // 32) if training else 0)
            narrow_by += n_slots * 2.0 * (pad16(chs) + pad16(chs))             # second conv: a1 in, z (or y) out
            if training:
                narrow_fl += f2 + (f1 if k > 0 else 0.0)                         # data gradients
                narrow_by += n_slots * 2.0 * 2 * pad16(chs) + n_slots * 4.0 * ((pad16(chs) + 31) // 32)
                if k > 0:
                    narrow_by += n_slots * 2.0 * (pad16(chs) + pad16(cin))
            cin = chs
    w = streams * chs
    wide_fl = (out_blocks - 1) * (conv_flops(B, H, W, w, w, 0) + conv_flops(B, H, W, w, w, 1))
    if training:
        wide_fl *= 2.0
    return wide_fl, narrow_fl, narrow_by


def train_step_flops(B, H, W, variant):
    """forward + weight gradients + data gradients (none for the first conv of each stream): SURVEY.md section 6."""
    fwd = net_forward_flops(B, H, W, variant)
    first = 4 * conv_flops(B, H, W, 27, 70, 0)
    return 3.0 * fwd - first


# ----------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.lines, self.proc = [], None
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(index), f'--query-gpu={self.Q}',
                                          '--format=csv,noheader,nounits', '-lms', '200'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return None
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], 0.0, set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for ts, line in self.lines:
            if ts < t0 or ts > t1 + 0.3:
                continue
            p = [x.strip() for x in line.split(',')]
            try:
                sm.append(float(p[0]))
                mx = max(mx, float(p[1]))
            except (ValueError, IndexError):
                continue
            for nm, val in zip(names, p[3:7]):
                if val.lower().startswith('active'):
                    reasons.add(nm)
        if not sm:
            return None
        return {'sm_mhz': float(np.median(sm)), 'sm_max_mhz': mx, 'reasons': sorted(reasons), 'samples': len(sm)}


# ----------------------------------------------------------------------------------------------- CPU baseline (oracle)
def cpu_train_step(variant, b, ps, threads):
    """One training step of the numpy port of the reference algorithm (oracle/): forward, loss, backward, Adam."""
    import torch
    import _fixtures as fx
    import oracle
    from oracle import losses as olosses
    torch.set_num_threads(threads)
    from oracle.init import default_state          # the reference's default init (torch.manual_seed(0)); no product code
    state = default_state(**model_kwargs(variant))
    net = oracle.FeedForwardOracle(state, model_uncert=(variant == 'upr'), model_discrete=(variant == 'dpp'))
    net.training = True
    rng = np.random.RandomState(0)
    views = [rng.uniform(0, 1, (b, 9, 3, ps, ps)).astype(np.float32) for _ in range(4)]
    gt = rng.uniform(-2, 2, (b, ps, ps)).astype(np.float32)
    mask = fx.synth_mask(1, b, ps, ps, margin=11)
    adam = {}

    def step(it):
        r = net.forward(*views, keep_tape=True)
        if variant == 'upr':
            _, g = olosses.improved_uncertainty_l1({'mean': r['mean'], 'logvar': r['logvar']}, gt, mask)
            gout = np.stack([g['mean'], g['logvar']], 1)
        elif variant == 'dpp':
            _, g = olosses.masked_cross_entropy({'scores': r['scores']}, olosses.reg_to_class(gt, -3.5, 3.5, 108), mask)
            gout = g['scores']
        else:
            _, g = olosses.masked_l1({'mean': r['mean']}, gt, mask)
            gout = g['mean'][:, None]
        grads = net.backward(gout)
        for k, gr in grads.items():
            m, v = adam.get(k, (np.zeros_like(net.p[k]), np.zeros_like(net.p[k])))
            net.p[k], m, v = oracle.adam_step(net.p[k], gr.reshape(net.p[k].shape), m, v, it + 1, 1e-3)
            adam[k] = (m, v)
    return step


def cpu_infer_step(variant, size, threads):
    import torch
    import oracle
    torch.set_num_threads(threads)
    from oracle.init import default_state
    state = default_state(**model_kwargs(variant))
    net = oracle.FeedForwardOracle(state, model_uncert=(variant == 'upr'), model_discrete=(variant == 'dpp'))
    rng = np.random.RandomState(0)
    views = [rng.uniform(0, 1, (1, 9, 3, size, size)).astype(np.float32) for _ in range(4)]
    return lambda it: net.forward(*views)


def cpu_ese_step(size, members, threads):
    import torch
    import oracle
    torch.set_num_threads(threads)
    from oracle.init import default_state
    state = default_state(**model_kwargs('upr'))
    net = oracle.FeedForwardOracle(state, model_uncert=True)
    rng = np.random.RandomState(0)
    views = [rng.uniform(0, 1, (1, 9, 3, size, size)).astype(np.float32) for _ in range(4)]
    step_size = 7.0 / members
    return lambda it: oracle.ensemble_forward(net, *views, -3.5, 3.5, step_size)


def run_cpu(args, steps, warmup):
    """Times the oracle port on the host cores on a bounded sample of the workload."""
    threads = os.cpu_count() or 1
    try:
        # torchrun exports OMP_NUM_THREADS=1 before the interpreter starts; the CPU arm uses all host cores
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=threads)
    except Exception:  # pragma: no cover
        pass
    if args.workload == 'train':
        b = args.cpu_batch
        step = cpu_train_step(args.variant, b, args.ps, threads)
        units, unit, sample = b, 'patches/s', f'{b} patches of {args.ps} px per step (fwd + loss + bwd + Adam), fp32 numpy'
    elif args.workload == 'ese':
        size, members = args.cpu_size, 2
        step = cpu_ese_step(size, members, threads)
        units, unit = size * size / 1e6 * members / ESE_MEMBERS, 'Mpx/s'
        sample = (f'{members} of the {ESE_MEMBERS} ensemble members (Shift + UPR forward + reduce) on one {size}x{size} '
                  f'crop per step, scaled x{members}/{ESE_MEMBERS}, fp32 numpy')
    else:
        size = args.cpu_size
        step = cpu_infer_step(args.variant, size, threads)
        units, unit = size * size / 1e6, 'Mpx/s'
        sample = f'one {size}x{size} crop of a 9x9-view light field per step, fp32 numpy'
    for i in range(warmup):
        step(i)
    t0 = time.time()
    for i in range(steps):
        step(warmup + i)
    dt = (time.time() - t0) / max(steps, 1)
    return {'value': units / dt, 'unit': unit, 'cores': threads, 'kind': 'port', 'sample': sample}, dt


# ----------------------------------------------------------------------------------------------- GPU arm
def describe(args, world):
    """metric / unit / config of one workload (BASELINE.json's metric and configs)."""
    metric = ('train patches/s (bs512, 96px)' if args.workload == 'train' else 'full-LF inference Mpx/s')
    unit = 'patches/s' if args.workload == 'train' else 'Mpx/s'
    wl = {'train': f'{args.variant.upper()} training step bs={args.bs} ps={args.ps}, 4-stream FeedForward '
                   f'(9 views, 70 ch, 108 bins), fwd+loss+bwd+allreduce+Adam',
          'infer': f'{args.variant.upper()} full-LF inference, one 9x9-view {args.size}x{args.size} light field per GPU '
                   f'and step',
          'bands': f'{args.variant.upper()} full-LF inference of ONE 9x9-view {args.size}x{args.size} light field per step, rows '
                   f'sharded over the GPUs in bands with an 11-px halo, gathered on every rank',
          'ese': f'ESE (--val_ensamble) full-LF shift-ensemble inference, {ESE_MEMBERS} UPR members (shift -3.5..3.4 step '
                 f'0.1) of one 9x9-view {args.size}x{args.size} light field per step, members sharded over the GPUs, '
                 f'incl. shifts + Laplace-mixture reduce'}[args.workload]
    strong = args.workload in ('train', 'ese', 'bands')
    config = {'workload': wl,
              'global_batch': args.bs if args.workload == 'train' else (1 if args.workload in ('ese', 'bands') else world),
              'parallelism': f'dp{world}', 'l2': 'inputs larger than L2 (>= 113 MB fp32 per step and GPU)',
              'activation_storage': args.precision, 'gradient_storage': 'bf16', 'accumulate': 'fp32'}
    return metric, unit, config, strong


def run_reference(args, world):
    """--impl reference: the CPU restatement of the reference algorithm (oracle/, numpy fp32) on the host cores, with the
    warm-up / step counts it was given, each step a bounded sample of the workload (stated in config.sample)."""
    metric, unit, config, strong = describe(args, world)
    base, dt = run_cpu(args, max(args.steps, 1), max(args.warmup, 0))
    config = {'workload': config['workload'], 'global_batch': config['global_batch'], 'parallelism': 'host cores',
              'sample': base['sample'], 'arithmetic': 'fp32 numpy (oracle/ port of the reference algorithm)'}
    return {'impl': 'reference', 'metric': metric, 'value': base['value'], 'unit': unit, 'n_gpus': world,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': dt * 1e3, 'higher_is_better': True,
            'scaling': 'strong' if strong else 'weak', 'vs_baseline': None, 'dtype': 'f32',
            'data': 'synthetic', 'config': config, 'cpu_baseline': base,
            'e2e': {'value': base['value'], 'unit': unit, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}


def run_gpu(args, rank, world, local, with_cpu_baseline, sample_clocks=True):
    """One workload on the GPUs -> the JSON line (dict) on rank 0, None elsewhere."""
    import gc
    import torch
    import torch.distributed as dist
    from mmlf_b200 import _lib, parallel
    from mmlf_b200.model.feed_forward import FeedForward
    from mmlf_b200.model import loss as L
    from mmlf_b200.optim import FusedAdam
    from mmlf_b200.train.step import TrainStep

    metric, unit, config, strong = describe(args, world)
    dev = torch.device('cuda', local)
    torch.manual_seed(0)
    model = FeedForward(**model_kwargs(args.variant)).to(dev)
    model.precision = args.precision
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    train_step = None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if args.workload == 'train':
        B = args.bs // world
        H = W = args.ps
        views = [torch.rand((B, 9, 3, H, W), device=dev, generator=gen) for _ in range(4)]
        gt = torch.rand((B, H, W), device=dev, generator=gen) * 4 - 2
        mask = L.create_mask_margin((B, H, W), 11).to(torch.int32).to(dev)
        opt = FusedAdam(model.parameters(), lr=1e-3)
        model.train()
        # the product's training step (mmlf.train.cli): forward + loss + backward + all-reduce + Adam, one graph replay;
        # DPP builds its one-hot class target inside the loss kernel (utils/dl.py:109-131)
        train_step = TrainStep(model, opt, {'base': 'l1', 'upr': 'upr', 'dpp': 'ce'}[args.variant], ce_from_gt=True)
        config['step'] = 'mmlf_b200.train.step.TrainStep (CUDA-graph replay)'

        def step(vs, gt_, mask_):
            return train_step(vs[0], vs[1], vs[2], vs[3], gt_, mask_)
        units_per_step = args.bs
        flops_per_step = train_step_flops(args.bs, H, W, args.variant)
        host = [t.cpu().pin_memory() for t in views + [gt, mask]]
    elif args.workload == 'ese':
        from mmlf_b200.model.ensamble import Ensamble
        B, H, W = 1, args.size, args.size
        gen = torch.Generator(device=dev).manual_seed(1234)          # every rank sees the same light field
        views = [torch.rand((B, 9, 3, H, W), device=dev, generator=gen) for _ in range(4)]
        model.eval()
        ens = Ensamble(model, -3.5, 3.5, 0.1)

        def step(vs, gt_=None, mask_=None):
            with torch.no_grad():
                return ens(*vs)['mean']
        units_per_step = H * W / 1e6
        flops_per_step = ESE_MEMBERS * net_forward_flops(1, H, W, 'upr')
        host = [t.cpu().pin_memory() for t in views]
        gt = mask = None
    else:
        B, H, W = 1, args.size, args.size
        views = [torch.rand((B, 9, 3, H, W), device=dev, generator=gen) for _ in range(4)]
        model.eval()

        if args.workload == 'bands':
            gen = torch.Generator(device=dev).manual_seed(1234)      # every rank holds the same light field
            views = [torch.rand((B, 9, 3, H, W), device=dev, generator=gen) for _ in range(4)]

        def step(vs, gt_=None, mask_=None):
            with torch.no_grad():
                if args.workload == 'bands':
                    if vs[0].shape[-2] != H:               # e2e: only this rank's band (+ halo) was uploaded
                        return parallel.banded_forward(model, vs, full_height=H)['mean']
                    return parallel.banded_forward(model, vs)['mean']
                out = model(*vs)
                _ = out['mean']
                if args.variant != 'base':
                    _ = out['posterior']
            return out['mean']
        n_lf = 1 if args.workload == 'bands' else world
        units_per_step = n_lf * H * W / 1e6
        flops_per_step = n_lf * net_forward_flops(1, H, W, args.variant)
        if args.workload == 'bands' and world > 1:
            # end to end, each rank reads only the rows of its band + halo from the host
            _lo, _hi, a_, b_ = parallel.band_rows(H, rank, world, 11)
            host = [t[..., a_:b_, :].contiguous().cpu().pin_memory() for t in views]
        else:
            host = [t.cpu().pin_memory() for t in views]
        gt = mask = None

    # ---------------- warm-up (the first training step also captures the graph)
    for _ in range(args.warmup):
        step(views, gt, mask)
    barrier()

    # ---------------- timed region: inputs resident in HBM, CUDA events, max over ranks
    # The timed region runs the product path without instrumentation (training and inference both replay CUDA graphs,
    # whose kernels cannot be bracketed one by one).  Per-kernel times come from a separate pass below: same work, one
    # stream, no graph, a CUDA-event pair around every launch.
    sampler = ClockSampler(local) if (rank == 0 and sample_clocks) else None
    _lib.launch_count = 0
    _lib.set_profile(False)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_wall0 = time.time()
    e0.record()
    for _ in range(args.steps):
        step(views, gt, mask)
    e1.record()
    host_ms_per_step = (time.time() - t_wall0) * 1e3 / args.steps      # time the host needs to enqueue one step
    barrier()
    t_wall1 = time.time()
    launches = _lib.launch_count
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_per_step = ms.item() / args.steps
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None

    # ---------------- end to end: host (pinned) inputs copied in every step, loss read back every step
    copy_stream = torch.cuda.Stream()
    bufs = [[torch.empty_like(t, device=dev) for t in host] for _ in range(2)]
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    h2d = sum(t.numel() * t.element_size() for t in host)

    def upload(slot):
        with torch.cuda.stream(copy_stream):
            for d, s_ in zip(bufs[slot], host):
                d.copy_(s_, non_blocking=True)
            ready[slot].record(copy_stream)

    e2e_steps = max(3, min(args.steps, 5))

    def run_e2e(upload_fn, step_fn):
        """One e2e pass of e2e_steps steps; returns the bytes read back per step.  Training reads its loss back with a host
        sync every step (the loop of train/cli.py logs it).  Inference keeps one step in flight: step i is enqueued, THEN the
        host waits for the read-back of step i - 1 (device copy of the static output buffer -> pinned host memory on the copy
        stream), so neither the read-back nor the host's enqueue time leaves the GPU idle; every step's result still
        reaches host memory inside the timed region."""
        cur = torch.cuda.current_stream()
        keep = host_out = None
        if args.workload != 'train':
            upload_fn(0)
            cur.wait_event(ready[0])
            o = step_fn(0)                                              # untimed: output shape, buffers of the pipeline
            keep = [torch.empty_like(o) for _ in range(2)]
            host_out = [torch.empty(o.shape, dtype=o.dtype).pin_memory() for _ in range(2)]
            torch.cuda.synchronize()
        done = [torch.cuda.Event(), torch.cuda.Event()]
        landed = [torch.cuda.Event(), torch.cuda.Event()]
        barrier()
        upload_fn(0)
        e0.record()
        nbytes = 0
        for i in range(e2e_steps):
            slot = i & 1
            cur.wait_event(ready[slot])
            res = step_fn(slot)                                         # enqueue only
            if args.workload == 'train':
                if i + 1 < e2e_steps:
                    # the other buffer set was last read by step i - 1, which that iteration's read-back synchronised:
                    # the next upload overlaps this step
                    upload_fn(slot ^ 1)
                _ = res.item()
                nbytes = 4
                continue
            keep[slot].copy_(res)
            done[slot].record(cur)
            if i > 0:
                landed[slot ^ 1].synchronize()                          # result of step i - 1 is in host memory;
            if i + 1 < e2e_steps:                                       # its input buffers are free again
                upload_fn(slot ^ 1)
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(done[slot])
                host_out[slot].copy_(keep[slot], non_blocking=True)
                landed[slot].record(copy_stream)
            nbytes = res.numel() * res.element_size()
        if args.workload != 'train':
            landed[(e2e_steps - 1) & 1].synchronize()
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item(), nbytes

    def step_f32(slot):
        cur = bufs[slot]
        return step(cur[:4], cur[4] if len(cur) > 4 else None, cur[5] if len(cur) > 5 else None)

    ms2_total, d2h = run_e2e(upload, step_f32)
    e2e = {'value': units_per_step / (ms2_total / e2e_steps / 1e3), 'unit': unit, 'h2d_bytes_per_step': h2d * world,
           'd2h_bytes_per_step': d2h * world, 'steps': e2e_steps,
           'note': 'pinned host inputs, double-buffered H2D on a copy stream overlapping the previous step; ' +
                   ('loss read back (one host sync) every step' if args.workload == 'train' else
                    'every result copied to pinned host memory, one step in flight while the previous result is awaited')}

    # ---------------- inference from the raw light field, as mmlf_b200.data.hci4d.HCI4D feeds it: of the 81 uint8 views a
    # scene holds (hci4d.py:151-193) only the 33 the four crosshair stacks read are copied in (25.9 MB instead of 113 MB
    # of float32 stacks), the extraction + u8 -> f32 runs on the GPU (mmlf_lf_extract_u8) in front of the forward
    e2e_u8 = None
    if args.workload == 'infer':
        try:
            from mmlf_b200.data import hci4d
            g8 = torch.Generator().manual_seed(99)
            us, vs, ids, dds = hci4d.view_indices((9, 9))
            need = sorted(set(us + vs + ids + dds))
            host8 = torch.randint(0, 256, (len(need), H, W, 3), dtype=torch.uint8, generator=g8).pin_memory()
            dbuf = [torch.zeros((81, H, W, 3), dtype=torch.uint8, device=dev) for _ in range(2)]

            def upload8(slot):
                with torch.cuda.stream(copy_stream):
                    for j, vi in enumerate(need):
                        dbuf[slot][vi].copy_(host8[j], non_blocking=True)
                    ready[slot].record(copy_stream)

            def step8(slot):
                hh, vv, ii, dd, _c = hci4d.extract_stacks(dbuf[slot], 9)
                return step([t.unsqueeze(0) for t in (hh, vv, ii, dd)], None, None)

            upload8(0)
            torch.cuda.current_stream().wait_event(ready[0])
            step8(0)                                                    # warm-up (graph capture for the new buffers)
            torch.cuda.synchronize()
            ms3_total, d2h8 = run_e2e(upload8, step8)
            e2e_u8 = {'value': units_per_step / (ms3_total / e2e_steps / 1e3), 'unit': unit,
                      'h2d_bytes_per_step': host8.numel() * world, 'd2h_bytes_per_step': d2h8 * world,
                      'steps': e2e_steps, 'note': 'the 33 uint8 views (H, W, 3) of the crosshair from pinned host memory (what '
                      'HCI4D.load_scene uploads), extraction (mmlf_lf_extract_u8) + forward on the GPU, every result copied '
                      'to pinned host memory (one step in flight)'}
        except Exception as ex:                                         # an extra measurement must not cost the bench line
            e2e_u8 = {'error': repr(ex)[:200]}
    del bufs

    # ---------------- profiling pass: un-graphed, one CUDA-event pair around every C-ABI call
    net = model
    prof_steps = (2 if args.workload == 'train' else 3) if args.profile_steps is None else args.profile_steps
    by_name, by_tag = {}, {}
    if prof_steps > 0:
        if train_step is not None:
            # the captured graph owns the activation memory of a whole step (88 GB at 512 patches): release it before the
            # eager pass allocates its own
            train_step._graphs.clear()
            gc.collect()
            torch.cuda.empty_cache()
            net.engine.overlap_wgrad = False
        else:
            net.use_cuda_graph = False
        _lib.set_profile(True)
        for i in range(prof_steps):
            if args.workload == 'train':
                step(views, gt, mask)             # the stream is always full at these sizes: pairs measure kernel time
                continue
            # a device-side sleep in front of every forward lets the host enqueue the whole forward before the first
            # kernel starts, so each event pair brackets exactly one kernel
            torch.cuda._sleep(20_000_000)
            with torch.no_grad():
                if args.workload == 'ese':
                    net.raw_forward(views, shift_disp=-3.5 + 0.1 * (7 * i + 3))
                else:
                    step(views, gt, mask)
        torch.cuda.synchronize()
        prof = _lib.set_profile(False)
        if train_step is None:
            net.use_cuda_graph = True
        barrier()
        for name, a, b in prof:
            base, _, tag = name.partition('#')           # the engine tags conv launches narrow / wide / head
            by_name.setdefault(base, []).append(a.elapsed_time(b))
            if tag:
                by_tag.setdefault((base, tag), []).append(a.elapsed_time(b))

    # ---------------- per-kernel shares and the roofline of the dominant kernel (the tcgen05 conv)
    shares = {k: sum(v) / prof_steps for k, v in by_name.items()}
    kernel_sum_ms = max(sum(shares.values()), 1e-9)           # summed kernel time of one profiled step / forward
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except OSError:
        pass
    peak_tf, peak_src = peaks.get('bf16_tflops_sustained'), 'MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)'
    if not peak_tf:
        peak_tf, peak_src = 1400.0, 'fallback (B200_PROFILING.md: sustained ~1.4 PFLOP/s)'
    conv_ms = by_name.get('mmlf_conv2x2', [])
    wgrad_ms = by_name.get('mmlf_conv2x2_wgrad', []) + by_name.get('mmlf_conv2x2_wgrad_canonical', [])
    roofline = None
    fwd = net_forward_flops(B, H, W, args.variant)
    if conv_ms:
        n_conv = len(conv_ms) / prof_steps
        # algorithmic flops of all conv2x2 launches of one step on this rank: forward convs + data gradients
        if args.workload == 'train':
            conv_flops_rank = 2.0 * fwd - 4 * conv_flops(B, H, W, 27, 70, 0)
            # the small head convs of BASE / UPR run partly on CUDA cores: negligible (< 0.1 %)
        elif args.workload == 'bands':
            lo_, hi_, a_, b_ = parallel.band_rows(H, rank, world, 11)
            conv_flops_rank = net_forward_flops(1, b_ - a_, W, args.variant)
        else:
            conv_flops_rank = fwd                            # ESE: the profiling pass times single members
        conv_time = sum(conv_ms) / prof_steps / 1e3
        achieved = conv_flops_rank / conv_time / 1e12
        roofline = {'kernel': 'conv2x2_tc_kernel (all forward + data-gradient launches of a step)', 'bound': 'tensor',
                    'achieved': achieved, 'peak': peak_tf, 'unit': 'TFLOP/s', 'frac': achieved / peak_tf,
                    'peak_source': peak_src, 'traffic': CONV_TRAFFIC['bytes'], 'traffic_note': CONV_TRAFFIC['note'],
                    'launches_per_step': n_conv, 'avg_launch_ms': sum(conv_ms) / len(conv_ms),
                    'share_of_step': conv_time * 1e3 / kernel_sum_ms,
                    'note': 'value: product path (CUDA-graph replay); kernel times: separate single-stream un-graphed pass '
                            'of %d step(s), CUDA events around every launch' % prof_steps}
        # the same launches split by the roof that binds them: the 280-channel layers against the
        # tensor peak, the 27 / 70-channel in-nets (67-122 FLOP per byte) against the HBM rate
        if by_tag and args.workload in ('train', 'infer'):
            hbm = peaks.get('hbm_gbs') or 6500.0
            wide_fl, narrow_fl, narrow_by = conv_split(B, H, W, args.variant, args.workload == 'train')
            split = {}
            t_w = sum(by_tag.get(('mmlf_conv2x2', 'wide'), [])) / prof_steps / 1e3
            t_n = sum(by_tag.get(('mmlf_conv2x2', 'narrow'), [])) / prof_steps / 1e3
            if t_w > 0:
                split['wide'] = {'bound': 'tensor', 'ms': t_w * 1e3, 'achieved': wide_fl / t_w / 1e12, 'unit': 'TFLOP/s',
                                 'frac': wide_fl / t_w / 1e12 / peak_tf}
            if t_n > 0:
                split['narrow'] = {'bound': 'hbm', 'ms': t_n * 1e3, 'achieved': narrow_by / t_n / 1e9, 'unit': 'GB/s',
                                   'peak': hbm, 'frac': narrow_by / t_n / 1e9 / hbm,
                                   'frac_of_tensor_peak': narrow_fl / t_n / 1e12 / peak_tf}
            roofline['by_layer_width'] = split
        if wgrad_ms:
            wg_time = sum(wgrad_ms) / prof_steps / 1e3
            roofline['wgrad'] = {'kernel': 'conv2x2_wgrad_kernel + reduce', 'achieved': fwd / wg_time / 1e12,
                                 'frac': fwd / wg_time / 1e12 / peak_tf, 'share_of_step': wg_time * 1e3 / kernel_sum_ms}
            if by_tag:
                wide_fl, narrow_fl, _ = conv_split(B, H, W, args.variant, False)
                for tag, fl in (('wide', wide_fl), ('narrow', narrow_fl)):
                    tt = sum(by_tag.get(('mmlf_conv2x2_wgrad_canonical', tag), [])) / prof_steps / 1e3
                    if tt > 0:
                        roofline['wgrad'][tag] = {'ms': tt * 1e3, 'achieved': fl / tt / 1e12, 'frac': fl / tt / 1e12 / peak_tf,
                                                  'frac_of_burst_peak': fl / tt / 1e12 / (peaks.get('bf16_tflops') or 1600.0)}
    step_tf = flops_per_step / world / (ms_per_step / 1e3) / 1e12
    if roofline is None:
        roofline = {'kernel': 'whole step (no per-kernel pass in this run)', 'bound': 'tensor', 'achieved': step_tf,
                    'peak': peak_tf, 'unit': 'TFLOP/s', 'frac': step_tf / peak_tf, 'peak_source': peak_src, 'traffic': None}
    roofline['step'] = {'achieved': step_tf, 'frac': step_tf / peak_tf,
                        'note': 'whole step, algorithmic FLOPs of SURVEY.md section 6 per GPU'}
    if shares:
        bn = sum(v for k, v in shares.items() if k.startswith('mmlf_bn_'))
        roofline['kernel_sum_ms'] = round(kernel_sum_ms, 3)
        # only meaningful for training: the inference / ESE passes time ONE un-graphed forward (ESE: one member of 70) with an
        # event pair per kernel, which is not comparable with the replayed step
        roofline['idle_frac_of_step'] = (round(max(0.0, 1.0 - kernel_sum_ms / ms_per_step), 4)
                                         if args.workload == 'train' else None)
        roofline['batchnorm_share_of_kernel_time'] = round(bn / kernel_sum_ms, 4)

    # free this workload's memory before the next one
    if train_step is not None:
        train_step.close()
    train_step = None
    del model, views, host
    gc.collect()
    torch.cuda.empty_cache()
    if rank != 0:
        return None
    cpu_base = None
    if world == 1 and with_cpu_baseline:
        cpu_base, _ = run_cpu(args, 2, 1)
    line = {'metric': metric, 'value': units_per_step / (ms_per_step / 1e3), 'unit': unit, 'n_gpus': world,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_per_step, 'higher_is_better': True,
            'scaling': 'strong' if strong else 'weak', 'vs_baseline': None,
            'dtype': args.precision + ' storage / fp32 accumulate', 'data': 'synthetic', 'config': config,
            'host_enqueue_ms_per_step': round(host_ms_per_step, 3), 'roofline': roofline, 'cpu_baseline': cpu_base,
            'e2e': e2e, 'gpu_launches': launches, 'clocks': clocks}
    if shares:
        line['kernel_ms_note'] = ('per training step, single-stream un-graphed profiling pass' if args.workload == 'train'
                                  else 'per single un-graphed forward (ESE: one member)')
        line['kernel_ms_per_step'] = {k: round(v, 3) for k, v in sorted(shares.items(), key=lambda kv: -kv[1])}
    if e2e_u8 is not None and 'value' in e2e_u8:
        # full-LF inference: the public data path is the HCI4D loader's (uint8 views in, extraction on the GPU); the
        # float32-stacks number (113 MB of H2D per light field: PCIe bound) is kept beside it
        line['e2e_f32_stacks'] = dict(e2e, note=e2e['note'] + '; inputs = the four float32 stacks')
        line['e2e'] = e2e_u8
    elif e2e_u8 is not None:
        line['e2e_u8_views'] = e2e_u8
    return line


# the other BASELINE.json configs, nested under `secondary` of the default line: (key, workload, variant, steps)
SECONDARY = [('infer', 'infer', 'base', 10), ('upr_train', 'train', 'upr', 4), ('dpp_train', 'train', 'dpp', 4),
             ('ese', 'ese', 'upr', 3), ('bands', 'bands', 'base', 10)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=8)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--workload', default=None, choices=['train', 'infer', 'ese', 'bands'])
    ap.add_argument('--variant', default=None, choices=['base', 'upr', 'dpp'])
    ap.add_argument('--bs', type=int, default=512, help='global batch (train)')
    ap.add_argument('--ps', type=int, default=96, help='patch size (train)')
    ap.add_argument('--size', type=int, default=512, help='light-field size (infer)')
    ap.add_argument('--precision', default='fp16', choices=['fp16', 'bf16', 'split'], help='activation storage format')
    ap.add_argument('--cpu-batch', type=int, default=2)
    ap.add_argument('--cpu-size', type=int, default=128)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-secondary', action='store_true', help='only the headline workload')
    ap.add_argument('--profile-steps', type=int, default=None, help='steps of the per-kernel pass (0 = skip it)')
    args = ap.parse_args()
    # the headline line (no --workload / --variant) also measures the other BASELINE configs
    secondary = args.workload is None and args.variant is None and not args.no_secondary and args.impl == 'b200'
    args.workload = args.workload or 'train'
    args.variant = args.variant or 'base'
    args.warmup = max(args.warmup, 3) if args.impl == 'b200' else args.warmup
    if args.workload == 'ese':
        args.variant = 'upr'                                 # --val_ensamble forces model_uncert (train/cli.py:68-69)

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if args.impl == 'reference':
        if rank == 0:
            print(json.dumps(run_reference(args, world)))
        return

    import copy
    import torch
    import torch.distributed as dist
    from mmlf_b200 import parallel
    rank, world, local = parallel.init_from_env('nccl')
    torch.cuda.set_device(local)
    line = run_gpu(args, rank, world, local, with_cpu_baseline=not args.no_cpu_baseline)
    if secondary:
        sec = {}
        for key, workload, variant, steps in SECONDARY:
            a = copy.copy(args)
            a.workload, a.variant, a.steps, a.warmup, a.profile_steps = workload, variant, steps, 3, 1
            try:
                res = run_gpu(a, rank, world, local, with_cpu_baseline=False, sample_clocks=False)
            except Exception as ex:                           # a secondary measurement must not cost the headline line
                res = {'error': repr(ex)[:300]}
                torch.cuda.empty_cache()
            if rank == 0 and res is not None:
                keep = ('metric', 'value', 'unit', 'ms_per_step', 'scaling', 'steps', 'warmup', 'config', 'e2e', 'roofline',
                        'gpu_launches', 'host_enqueue_ms_per_step', 'kernel_ms_per_step', 'e2e_u8_views', 'e2e_f32_stacks', 'error')
                sec[key] = {k: res[k] for k in keep if k in res}
        if rank == 0:
            line['secondary'] = sec
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()

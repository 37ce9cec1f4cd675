#!/usr/bin/env python
"""Benchmark of the logits -> loss -> metrics hot path (BASELINE.json metric: Mpix/s, % of B200 HBM peak).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--no-extras]

Headline workload (config.workload, BASELINE.json configs[1]): FCN-style head at Cityscapes shape — fp32 logits
(8,19,64,128) bilinearly resized to 512x1024 labels, cross-entropy with ignore_index=255, forward AND backward, plus
the in-loop top-1 accuracy. A "step" is one such batch per GPU; a pixel is a label-resolution pixel. Under
torchrun every rank runs the same per-GPU batch (weak scaling, images are independent) and the only exchange is the
all-reduce of the steps' 8-double statistics vectors (logging only), packed and issued once per ring cycle of 8 steps on a
side stream (--allreduce-every 1 issues it every step).

  value         device-resident throughput: CUDA-graph replays of the step over a ring of input sets larger than L2,
                timed with CUDA events on the launching stream, max over ranks.
  e2e           the same step through the public API with HOST inputs: pinned H2D of logits + labels every step,
                fused_resize_losses + backward, D2H read of the loss.
  roofline      the dominant kernel (up_gen_kernel) timed alone with CUDA events: algorithmic bytes / time vs the
                measured HBM copy peak (MEASURED_PEAKS.json).
  cpu_baseline  the reference's own resize / CrossEntropyLoss / accuracy (oracle/_ref: byte code compiled from its files
                by oracle/build_ref.py; kind "reference") on the host cores, bounded sample; the oracle port
                (oracle/oracle.py; kind "port") when oracle/_ref is absent.
  workloads     extra single-GPU results for BASELINE configs 3, 4 and 5 (HBM-bound shapes), each with its roofline;
                mirrored as flat scalars into roofline (c3_*, c4_*, c5i_*, c5ii_*, c5_resized_*, f1_*, lovasz_*, and
                under torchrun c4dp_* / c5i_sharded_*). c5i / c5ii are measured twice: through the list API the
                reference's intersect_and_union has (per-image areas, image table built inside the call: *_ms / *_frac)
                and over a prepared list with in-kernel totals (*_totals_ms / *_totals_frac: the kernel alone).

`--impl reference` times the reference's CPU implementation of the same workload on all host threads: its own files,
executed from oracle/_ref byte code (the sources stay in /root/reference; see oracle/build_ref.py), else the oracle port.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
import warnings

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = 'Mpix/s: logits resize+CE fwd/bwd and mIoU eval; % of B200 HBM peak'
C2 = dict(N=8, C=19, h=64, w=128, H=512, W=1024, ignore=255)
HBM_FALLBACK_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md, used only when MEASURED_PEAKS.json is absent


def hbm_peak():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    try:
        with open(p) as fh:
            return float(json.load(fh)['hbm_gbs']), 'measured'
    except Exception:
        return HBM_FALLBACK_GBS, 'fallback'


# ---------------------------------------------------------------------------------------------- synthetic data
def make_logits(shape, seed, dtype=torch.float32, device='cpu'):
    """randn*2, +1 on one class per pixel, quantised to 2^-6 (no soft-max rounding ties)."""
    g = torch.Generator(device=device).manual_seed(seed)
    n, c, h, w = shape
    x = torch.randn(shape, generator=g, device=device) * 2.0
    hot = torch.randint(0, c, (n, 1, h, w), generator=g, device=device)
    x.scatter_add_(1, hot, torch.ones((n, 1, h, w), device=device))
    x = torch.round(x * 64.0) / 64.0
    return x.to(dtype)


def make_labels(shape, num_classes, seed, ignore=255, frac=0.1, block=16, dtype=torch.int64, device='cpu'):
    """16x16-pixel constant blocks of random classes, ~10% of the blocks set to ignore_index."""
    g = torch.Generator(device=device).manual_seed(seed + 7)
    n, h, w = shape
    bh, bw = (h + block - 1) // block, (w + block - 1) // block
    y = torch.randint(0, num_classes, (n, bh, bw), generator=g, device=device)
    if ignore is not None and frac > 0:
        y[torch.rand((n, bh, bw), generator=g, device=device) < frac] = ignore
    y = y.repeat_interleave(block, 1).repeat_interleave(block, 2)[:, :h, :w].contiguous()
    return y.to(dtype)


# ---------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None
        self.t = None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.gpu), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '100'], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        self.t = threading.Thread(target=self._pump, daemon=True)
        self.t.start()

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in self.rows:
            f = [x.strip() for x in r.split(',')]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max(smax) if smax else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


# ---------------------------------------------------------------------------------------------- reference arm
_REF = None


def cpu_kind():
    """'reference' when the reference's own files can be executed here (oracle/_ref: byte code compiled from the sources
    under /root/reference by oracle/build_ref.py — it travels to the GPU box), else 'port' (oracle/oracle.py, which
    tests/test_reference_live.py pins bit for bit to those files)."""
    global _REF
    if _REF is None:
        from oracle import ref_loader
        try:
            _REF = (ref_loader.load(), ref_loader.origin()) if ref_loader.available() else (None, None)
        except Exception:
            _REF = (None, None)
    return 'reference' if _REF[0] is not None else 'port'


def cpu_step(x, y, ignore):
    x.grad = None
    if cpu_kind() == 'reference':
        # BaseDecodeHead.losses as the reference runs it (models/decode_heads/decode_head.py:261-295), on its own code
        R = _REF[0]
        if not hasattr(R, '_bench_ce'):
            R._bench_ce = R.CrossEntropyLoss()
        full = R.resize(input=x, size=y.shape[2:], mode='bilinear', align_corners=False)
        lab = y.squeeze(1)
        out = {'loss_ce': R._bench_ce(full, lab, weight=None, ignore_index=ignore),
               'acc_seg': R.accuracy(full, lab, ignore_index=ignore)}
    else:
        from oracle import oracle as O
        out = O.head_losses(x, y, [('ce', {}, 'loss_ce')], align_corners=False, ignore_index=ignore)
    out['loss_ce'].backward()
    return out


def cpu_kind_note():
    return ('the UNMODIFIED reference files (resize, CrossEntropyLoss, accuracy) executed from oracle/_ref byte code'
            if cpu_kind() == 'reference' else 'oracle port of the reference chain (oracle/oracle.py)')


def run_cpu(steps, warmup, n_images):
    """The reference's CPU path (its own files when oracle/_ref is present, else the oracle port) on a bounded sample:
    n_images of the C2 batch per step."""
    warnings.simplefilter('ignore')
    torch.set_num_threads(os.cpu_count() or 1)
    x = make_logits((n_images, C2['C'], C2['h'], C2['w']), 1234 + 100).requires_grad_(True)
    y = make_labels((n_images, C2['H'], C2['W']), C2['C'], 1234 + 100, C2['ignore']).unsqueeze(1)
    for _ in range(warmup):
        cpu_step(x, y, C2['ignore'])
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_step(x, y, C2['ignore'])
    dt = time.perf_counter() - t0
    px = n_images * C2['H'] * C2['W'] * steps
    return px / dt / 1e6, dt / steps * 1e3


def reference_main(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    n_img = C2['N']          # the full batch of the stated config (~0.2 s per step on 16 host cores)
    value, ms = run_cpu(args.steps, args.warmup, n_img)
    cores = torch.get_num_threads()
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': 'Mpix/s', 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic',
        'config': workload_config(),
        'cpu_baseline': {'value': value, 'unit': 'Mpix/s', 'cores': cores, 'kind': cpu_kind(),
                         'sample': 'full batch: %d of the %d images per step (resize + CE fwd/bwd + accuracy, torch %s CPU, '
                                   'os.cpu_count()=%s); %s' % (n_img, C2['N'], torch.__version__, os.cpu_count(), cpu_kind_note())},
        'e2e': {'value': value, 'unit': 'Mpix/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


def workload_config():
    return {'workload': 'C2: FCN-style head, Cityscapes shape — fp32 logits (8,19,64,128) bilinear-resized (align_corners=False) '
                        'to 512x1024, CE ignore_index=255 + top-1 accuracy, forward+backward',
            'batch_per_gpu': C2['N'], 'num_classes': C2['C'], 'logit_hw': [C2['h'], C2['w']], 'label_hw': [C2['H'], C2['W']],
            'label_dtype': 'int64 on the device for value / roofline (the reference feeds label.long()); uint8 host label maps for e2e',
            'pixels_per_step_per_gpu': C2['N'] * C2['H'] * C2['W']}


# ---------------------------------------------------------------------------------------------- b200 arm
def timed_events(fn, iters, stream=None):
    """ms per call of fn() over iters calls, CUDA events on the current stream."""
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    s.record()
    for i in range(iters):
        fn(i)
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


class GraphRing:
    """A step captured once per input set (the ring is larger than L2) and replayed; under torchrun the steps' 8-double
    statistics vectors (logging only: the gradient's denominator is known a priori) are all-reduced ONCE PER RING CYCLE —
    `every` steps packed into one buffer, one NCCL call on a side stream ordered by events, off the compute stream."""

    def __init__(self, step, n_sets, dev, world, dist, every=None):
        self.R, self.world, self.dist, self.dev = n_sets, world, dist, dev
        self.every = n_sets if every is None else max(1, min(int(every), n_sets))
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for i in range(n_sets):
                step(i)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graphs, self.outs = [], []
        for i in range(n_sets):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self.outs.append(step(i))
            self.graphs.append(g)
        self.stats = [o['_stats'] for o in self.outs]
        self.comm = torch.cuda.Stream(device=dev)
        self.ready = [torch.cuda.Event(), torch.cuda.Event()]
        self.done = [torch.cuda.Event(), torch.cuda.Event()]   # the all-reduce that last used packed buffer 0 / 1 has finished
        self.bufs = [torch.zeros(n_sets * self.stats[0].numel(), dtype=self.stats[0].dtype, device=dev) for _ in range(2)]
        self.cycle = 0
        self.pending = []                      # ring slots whose statistics have not been reduced yet
        self.collectives = 0

    def _reduce_pending(self):
        """Packs the pending steps' statistics on the COMPUTE stream (one small cat) and hands the packed buffer to the
        side stream for the all-reduce; two packed buffers alternate, so the compute stream only ever waits for a
        collective issued two cycles earlier."""
        if not self.pending:
            return
        cur = torch.cuda.current_stream()
        b = self.cycle & 1
        cur.wait_event(self.done[b])
        n = len(self.pending) * self.stats[0].numel()
        buf = self.bufs[b][:n]
        torch.cat([self.stats[i] for i in self.pending], out=buf)
        self.ready[b].record(cur)
        self.comm.wait_event(self.ready[b])
        with torch.cuda.stream(self.comm):
            self.dist.all_reduce(buf, op=self.dist.ReduceOp.SUM)
            self.done[b].record(self.comm)
        self.pending = []
        self.cycle += 1
        self.collectives += 1

    def run_step(self, k):
        i = k % self.R
        self.graphs[i].replay()
        if self.world > 1:
            self.pending.append(i)
            if len(self.pending) >= self.every:
                self._reduce_pending()

    def drain(self):
        if self.world > 1:
            self._reduce_pending()             # a partial cycle is reduced inside the timed region
        torch.cuda.current_stream().wait_stream(self.comm)

    def timed(self, K, W):
        """W warm-up steps, barrier, 3 more untimed steps (every rank is in steady state when its clock starts), then
        exactly K steps between two CUDA events on the launching stream; max over ranks. Returns ms per step."""
        for k in range(W):
            self.run_step(k)
        self.drain()
        torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        for k in range(3):
            self.run_step(W + k)
        s_ev, e_ev = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s_ev.record()
        for k in range(K):
            self.run_step(W + 3 + k)
        self.drain()
        e_ev.record()
        torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        t = torch.tensor([s_ev.elapsed_time(e_ev)], dtype=torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item()) / K


def b200_main(args):
    import torch.distributed as dist

    import image_segmentation_lab_b200 as B
    from image_segmentation_lab_b200 import _lib
    from image_segmentation_lab_b200 import distributed as D

    warnings.simplefilter('ignore')
    rank, local, world = D.init_from_env()
    if torch.cuda.is_available():
        torch.cuda.set_device(local)
    else:
        raise RuntimeError('bench.py needs a CUDA device: the B200 path has no CPU fallback')
    dev = torch.device('cuda', local)
    numa_bound = D.bind_to_gpu_numa_node(local) if world > 1 else False   # before any pinned allocation
    lib = B.load_library()
    peak, peak_kind = hbm_peak()
    K, W = args.steps, max(args.warmup, 3)
    N, Cc, h, w, H, Wd, ign = C2['N'], C2['C'], C2['h'], C2['w'], C2['H'], C2['W'], C2['ignore']
    px_step = N * H * Wd
    # the clock sampler is a subprocess: started BEFORE the warm-up so that its start-up never sits between the barrier
    # and the first timed step of rank 0 (the other ranks would wait for rank 0's all-reduce and report the wait)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()

    # ---- ring of input sets: 8 x (5.0 MB logits + 33.6 MB labels) = 308 MB > 126 MB L2
    R = 8
    seed0 = 1234 + 100 + rank
    xs = [make_logits((N, Cc, h, w), seed0 + 1000 * i, device=dev).requires_grad_(True) for i in range(R)]
    ys = [make_labels((N, H, Wd), Cc, seed0 + 1000 * i, ign, device=dev).unsqueeze(1) for i in range(R)]
    ce = B.CrossEntropyLoss()

    def step(i):
        xs[i].grad = None
        r = B.fused_resize_losses(xs[i], ys[i], ce, align_corners=False, ignore_index=ign, return_stats=True)
        r['loss_ce'].backward()
        return r

    # launches per step (host counter; graph replays do not pass through the host-side counter)
    torch.cuda.synchronize()
    c0 = B.launch_count()
    step(0)
    torch.cuda.synchronize()
    launches_per_step = B.launch_count() - c0

    ring = GraphRing(step, R, dev, world, dist, every=args.allreduce_every)
    ms_step = ring.timed(K, W)
    value = world * px_step / (ms_step * 1e-3) / 1e6
    ar_every = ring.every

    # ---- e2e: host inputs, H2D + step + D2H every step, through the public API. The copies of step k+1 are issued on a
    # copy stream before step k's kernels (double-buffered device inputs), as a DataLoader with pin_memory +
    # non_blocking copies does; every step still ends with the D2H read of its loss (parse_losses' .item()).
    e2e_steps = max(10, min(K, 50))

    def e2e_run(label_dtype):
        xh = [make_logits((N, Cc, h, w), seed0 + 77 + i).pin_memory() for i in range(2)]
        yh = [make_labels((N, H, Wd), Cc, seed0 + 77 + i, ign).unsqueeze(1).to(label_dtype).pin_memory() for i in range(2)]
        xd = [torch.empty((N, Cc, h, w), device=dev).requires_grad_(True) for _ in range(2)]
        yd = [torch.empty((N, 1, H, Wd), dtype=label_dtype, device=dev) for _ in range(2)]
        copy_stream = torch.cuda.Stream(device=dev)
        arrived = [torch.cuda.Event() for _ in range(2)]
        host_loss = []

        def issue_copy(k):
            j = k & 1
            with torch.cuda.stream(copy_stream):
                with torch.no_grad():
                    xd[j].copy_(xh[j], non_blocking=True)
                yd[j].copy_(yh[j], non_blocking=True)
                arrived[j].record(copy_stream)

        # the step itself is captured once per device buffer set through the public API (fused_resize_losses + backward),
        # as the device-resident loop does: per step the host issues two copies, one graph replay and one D2H read
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for j in range(2):
                xd[j].grad = None
                B.fused_resize_losses(xd[j], yd[j], ce, align_corners=False, ignore_index=ign)['loss_ce'].backward()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graphs, losses = [], []
        for j in range(2):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                xd[j].grad = None
                r = B.fused_resize_losses(xd[j], yd[j], ce, align_corners=False, ignore_index=ign)
                r['loss_ce'].backward()
            graphs.append(g)
            losses.append(r['loss_ce'])

        def run(steps):
            issue_copy(0)
            for k in range(steps):
                j = k & 1
                if k + 1 < steps:
                    issue_copy(k + 1)      # buffer (k+1)&1 was released by the .item() of step k-1
                torch.cuda.current_stream().wait_event(arrived[j])
                graphs[j].replay()
                host_loss.append(losses[j].item())  # D2H read of the step's result (synchronises, as parse_losses does)

        # untimed warm-up long enough for the PCIe link to leave its power-saving state (the device-resident phases before
        # this one move nothing over it: the first timed repeats of a cold link measured 13.5 / 15.1 / 16.1 Gpix/s against
        # 23.7 warm)
        for _ in range(4):
            run(e2e_steps)
        vals = []
        for _ in range(5):      # host-timed and PCIe-bound: five repeats, the median is reported (all are kept)
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            run(e2e_steps)
            torch.cuda.synchronize()
            tt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            vals.append(world * px_step * e2e_steps / float(tt.item()) / 1e6)
        return sorted(vals)[len(vals) // 2], xh[0].numel() * 4 + yh[0].numel() * yh[0].element_size(), vals

    # host label maps are uint8, as a segmentation pipeline delivers them (PNG masks; the reference casts with .long()
    # AFTER the copy, cross_entropy_loss.py:283) — the kernels read uint8 directly. The int64-host variant is kept beside it.
    e2e_value, h2d, e2e_runs = e2e_run(torch.uint8)
    e2e_i64_value, h2d_i64, _ = e2e_run(torch.int64)
    clocks = sampler.stop() if rank == 0 else None

    # ---- dominant kernel alone (C ABI, CUDA events on the launching stream)
    roof = None
    extras = {}
    cpu_base = None
    if rank == 0:
        roof = kernel_roofline(lib, _lib, xs, ys, N, Cc, h, w, H, Wd, ign, peak, peak_kind)
    del ring
    if rank == 0 and world == 1:
        if not args.no_extras:
            try:
                extras = extra_workloads(B, _lib, dev, peak, peak_kind)
            except Exception as ex:  # extras never invalidate the headline
                extras = {'error': repr(ex)}
        v, ms_cpu = run_cpu(6, 1, C2['N'])
        cpu_base = {'value': v, 'unit': 'Mpix/s', 'cores': torch.get_num_threads(), 'kind': cpu_kind(),
                    'sample': 'full C2 batch (8 images), 6 timed steps of resize + CE fwd/bwd + accuracy on the host CPU '
                              '(torch %s, os.cpu_count()=%s), %.0f ms/step; %s'
                              % (torch.__version__, os.cpu_count(), ms_cpu, cpu_kind_note())}
        # the same restatement on CUDA tensors: the reference's unfused ATen chain on THIS GPU ("the kernel to
        # beat on the same box") — a reported baseline like cpu_baseline, never on the product path
        try:
            xa = xs[0].detach().clone().requires_grad_(True)

            def aten_step(i):
                cpu_step(xa, ys[0], ign)

            for i in range(2):
                aten_step(i)
            ms_aten = timed_events(aten_step, 10)
            cpu_base['aten_same_gpu_mpix_s'] = px_step / ms_aten / 1e3
            cpu_base['aten_same_gpu_ms_per_step'] = ms_aten
            cpu_base['aten_same_gpu_note'] = ('F.interpolate -> F.cross_entropy -> weight_reduce_loss -> topk accuracy and '
                                              'their autograd backward on the same B200 and inputs')
            del xa
            torch.cuda.empty_cache()
        except Exception as ex:
            cpu_base['aten_same_gpu_error'] = repr(ex)

    if world > 1 and not args.no_extras:
        del xs, ys
        torch.cuda.empty_cache()
        try:
            c4 = c4_data_parallel(B, dist, dev, rank, world, peak, peak_kind, K, W, args.allreduce_every)
        except Exception as ex:
            c4 = {'error': repr(ex)}
        sharded = c5_sharded(B, D, dist, dev, rank, world, peak, peak_kind)
        if rank == 0:
            extras['C4_dp'] = c4
            extras['C5i_miou_label_maps_sharded'] = sharded

    if rank == 0:
        line = {
            'metric': METRIC, 'value': value, 'unit': 'Mpix/s', 'n_gpus': world, 'steps': K, 'warmup': W,
            'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
            'data': 'synthetic',
            'config': dict(workload_config(), l2='ring of %d input sets (%.0f MB) > 126 MB L2; one CUDA graph per set' % (
                R, R * (N * Cc * h * w * 4 + N * H * Wd * 8) / 1e6),
                cpu_affinity='GPU-local NUMA node (NVML)' if numa_bound else 'inherited',
                collective=('1 all_reduce per %d steps: the steps\' 8-double statistics vectors (logging only) packed into one '
                            'buffer, on a side stream' % ar_every) if world > 1 else 'none (single GPU)'),
            'clocks': clocks,
            'e2e': {'value': e2e_value, 'unit': 'Mpix/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': 4,
                    'steps': e2e_steps, 'host_label_dtype': 'uint8',
                    'repeats_mpix_s': [round(v, 1) for v in e2e_runs],
                    'int64_host_labels_value': e2e_i64_value, 'int64_host_labels_h2d_bytes_per_step': h2d_i64,
                    'note': 'pinned host fp32 logits + uint8 label maps (as a pipeline delivers masks; read directly by the '
                            'kernels) copied every step on a copy stream into double-buffered device inputs; the step '
                            '(fused_resize_losses + backward) replayed from a CUDA graph captured through the public API; loss read '
                            'back every step; median of 5 repeats after an untimed warm-up of the link. int64_host_labels_value = same loop with int64 host labels'},
            'gpu_launches': int(launches_per_step * K),
            'launches_per_step': int(launches_per_step),
            'roofline': roof,
        }
        if cpu_base is not None:
            line['cpu_baseline'] = cpu_base
        if extras:
            line['workloads'] = extras
            flatten_workloads(line, extras)
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def flatten_workloads(line, extras):
    """The driver keeps scalar keys of the known objects: every workload's ms / roofline fraction is mirrored into
    `roofline` as flat scalars (c3_*, c4_*, c5*_*), next to the nested `workloads` detail."""
    r = line.get('roofline')
    if not isinstance(r, dict):
        return

    def put(key, d, *path):
        try:
            for k in path:
                d = d[k]
            r[key] = d
        except Exception:
            pass

    put('c3_fwd_bwd_ms', extras, 'C3_ade20k_bf16_ce_dice', 'fwd_bwd', 'ms')
    put('c3_fwd_bwd_frac', extras, 'C3_ade20k_bf16_ce_dice', 'fwd_bwd', 'roofline', 'frac')
    put('c3_fwd_ms', extras, 'C3_ade20k_bf16_ce_dice', 'fwd', 'ms')
    put('c3_fwd_frac', extras, 'C3_ade20k_bf16_ce_dice', 'fwd', 'roofline', 'frac')
    put('c3_ce_only_fwd_bwd_frac', extras, 'C3_ade20k_bf16_ce_only', 'fwd_bwd', 'roofline', 'frac')
    put('c4_fwd_bwd_ms', extras, 'C4_voc_fp32_ce', 'fwd_bwd', 'ms')
    put('c4_fwd_bwd_frac', extras, 'C4_voc_fp32_ce', 'fwd_bwd', 'roofline', 'frac')
    put('c4_fwd_frac', extras, 'C4_voc_fp32_ce', 'fwd', 'roofline', 'frac')
    put('c2_ac_false_ms', extras, 'C2_align_corners_false', 'fwd_bwd', 'ms')
    put('c2_ac_true_ms', extras, 'C2_align_corners_true', 'fwd_bwd', 'ms')
    put('c2_c150_ns_per_px_class', extras, 'C2_like_c150_64to512', 'fwd_bwd', 'ns_per_px_class')
    put('c2_ac_true_ns_per_px_class', extras, 'C2_align_corners_true', 'fwd_bwd', 'ns_per_px_class')
    put('c2_c150_general_ms', extras, 'C2_like_c150_64to512', 'fwd_bwd', 'ms')
    put('c5i_ms', extras, 'C5i_miou_label_maps', 'ms')
    put('c5i_frac', extras, 'C5i_miou_label_maps', 'roofline', 'frac')
    put('c5i_totals_ms', extras, 'C5i_miou_label_maps', 'totals_prepared', 'ms')
    put('c5i_totals_frac', extras, 'C5i_miou_label_maps', 'totals_prepared', 'roofline', 'frac')
    put('c5ii_ms', extras, 'C5ii_miou_from_logits', 'ms')
    put('c5ii_frac', extras, 'C5ii_miou_from_logits', 'roofline', 'frac')
    put('c5ii_totals_ms', extras, 'C5ii_miou_from_logits', 'totals_prepared', 'ms')
    put('c5ii_totals_frac', extras, 'C5ii_miou_from_logits', 'totals_prepared', 'roofline', 'frac')
    put('c5_resized_ms', extras, 'C5_resized_lowres_logits', 'ms')
    put('c5_resized_frac', extras, 'C5_resized_lowres_logits', 'roofline', 'frac')
    put('f1_sigmoid_ce_fwd_bwd_ms', extras, 'F1_sigmoid_ce_2class', 'fwd_bwd', 'ms')
    put('f1_sigmoid_ce_fwd_bwd_frac', extras, 'F1_sigmoid_ce_2class', 'fwd_bwd', 'roofline', 'frac')
    put('lovasz_fwd_bwd_ms', extras, 'lovasz_softmax_cityscapes_shape', 'fwd_bwd', 'ms')
    put('c4dp_strong_ms', extras, 'C4_dp', 'strong', 'ms')
    put('c4dp_strong_frac', extras, 'C4_dp', 'strong', 'roofline', 'frac')
    put('c4dp_strong_mpix_s', extras, 'C4_dp', 'strong', 'mpix_s')
    put('c4dp_weak_ms', extras, 'C4_dp', 'weak', 'ms')
    put('c4dp_weak_frac', extras, 'C4_dp', 'weak', 'roofline', 'frac')
    put('c4dp_weak_mpix_s', extras, 'C4_dp', 'weak', 'mpix_s')
    put('c5i_sharded_ms', extras, 'C5i_miou_label_maps_sharded', 'ms')
    put('c5i_sharded_frac', extras, 'C5i_miou_label_maps_sharded', 'roofline', 'frac')
    put('c5i_sharded_mpix_s', extras, 'C5i_miou_label_maps_sharded', 'mpix_s')


def kernel_roofline(lib, _lib, xs, ys, N, Cc, h, w, H, Wd, ign, peak, peak_kind):
    """The resize-fused CE kernel (up_gen_kernel) alone: b200seg_loss_fused_fwdbwd(defer_combine=1) launches exactly it."""
    import ctypes as C
    dev = xs[0].device
    R = len(xs)
    nbytes = lib.b200seg_loss_fused_workspace_bytes(N, Cc, h, w, H, Wd, 0)
    pbs = [torch.empty(nbytes // 4, dtype=torch.float32, device=dev) for _ in range(2)]
    stats = torch.zeros(8, dtype=torch.int64, device=dev)
    descs = []
    for i in range(R):
        fu = _lib.LossFusedDesc()
        fd = fu.fwd
        fd.logits = xs[i].data_ptr(); fd.labels = ys[i].data_ptr()
        fd.logit_dtype = _lib.F32; fd.label_dtype = _lib.L_I64
        fd.N, fd.C, fd.h, fd.w, fd.H, fd.W = N, Cc, h, w, H, Wd
        fd.flags = _lib.WANT_CE | _lib.WANT_ACC
        fd.ignore_index = ign; fd.acc_has_ignore = 1; fd.acc_ignore_index = ign
        fd.dice_exponent = 2.0; fd.ce_loss_weight = 1.0
        fd.stats = stats.data_ptr()
        fu.grad_scale_host = 1.0
        fu.workspace = pbs[i & 1].data_ptr()
        fu.defer_combine = 1
        descs.append(fu)
    stream = _lib.stream_ptr(dev)

    def call(i):
        rc = lib.b200seg_loss_fused_fwdbwd(C.byref(descs[i % R]), stream)
        if rc:
            raise RuntimeError(_lib.last_error())

    for i in range(10):
        call(i)
    ms = timed_events(call, 200)
    s = 4
    algo = 2 * N * Cc * h * w * s + N * H * Wd * 8          # logits read + gradient written + int64 labels read
    achieved = algo / (ms * 1e-3) / 1e9
    traffic, warp_inst, src, pipe_pct = None, None, None, None
    for name in ('traffic_r2e.json', 'traffic_r2d.json'):   # DRAM bytes / warp instructions per launch from the committed ncu capture
        try:
            with open(os.path.join(ROOT, 'profiles', name)) as fh:
                ks = json.load(fh)['kernels']
            k = next(v for kk, v in ks.items() if kk.startswith('up_gen_kernel<float, 8, 1'))
            traffic, warp_inst, src, pipe_pct = k['dram_bytes'], k['warp_inst'], 'profiles/' + name, k.get('l1_data_pipe_pct')
            break
        except Exception:
            continue
    out = {'bound': 'hbm', 'kernel': 'up_gen_kernel<float,PXC=8,GRAD,int64 labels,32 threads>', 'achieved': achieved, 'peak': peak,
           'unit': 'GB/s', 'frac': achieved / peak, 'traffic': traffic,
           'traffic_source': (src + ' (ncu --set full capture of the same kernel and shape; not measured in this run)') if src else None,
           'peak_kind': peak_kind, 'ms_per_launch': ms,
           'algorithmic_bytes_per_launch': algo,
           'note': 'not HBM bound: with the logits at 1/8 resolution the only full-resolution tensor touched is the label map '
                   '(33.6 of the 43.5 MB), while every output pixel owes ~2.3 issue slots per class (packed fp32 math) plus ~47 of '
                   'per-pixel work. The unit it sits on is the L1 / shared-memory data pipe (l1_data_pipe_frac, from the committed '
                   'ncu capture) followed by issue (issue_frac); see DESIGN.md. The HBM-bound kernels of the path are the c3_/c4_/c5 keys'}
    if pipe_pct is not None:
        out['l1_data_pipe_frac'] = pipe_pct / 100.0
    if warp_inst:
        sm_clock = 1.965e9
        issue_peak = 148 * 4 * sm_clock          # warp instructions / s: 4 schedulers x 148 SMs at clocks.max.sm
        out['issue_frac'] = warp_inst / (ms * 1e-3) / issue_peak
        out['issue_warp_inst_per_launch'] = warp_inst
        out['issue_peak_warp_inst_s'] = issue_peak
    return out


def c4_data_parallel(B, dist, dev, rank, world, peak, peak_kind, K, W, ar_every=None):
    """BASELINE config 4 (PSPNet head at Pascal VOC shape: 21 classes, 512x512, fp32 CE fwd+bwd, batch 32) data-parallel
    over the ranks: STRONG = the 32 images split over the GPUs, WEAK = 32 images on every GPU. CUDA-graph replay over a ring
    of input sets larger than L2, one all-reduce of the statistics vector per step on a side stream."""
    out = {}
    Cn, Hh, Ww, ign = 21, 512, 512, 255
    ce = B.CrossEntropyLoss()
    for mode in ('strong', 'weak'):
        n_loc = max(32 // world, 1) if mode == 'strong' else 32
        set_bytes = n_loc * Cn * Hh * Ww * 4 * 2 + n_loc * Hh * Ww * 8
        R = max(2, min(8, int(400e6 // set_bytes) + 1))
        xs = [make_logits((n_loc, Cn, Hh, Ww), 4000 + 10 * i + rank, device=dev).requires_grad_(True) for i in range(R)]
        ys = [make_labels((n_loc, Hh, Ww), Cn, 4000 + 10 * i + rank, ign, device=dev).unsqueeze(1) for i in range(R)]

        def step(i):
            xs[i].grad = None
            r = B.fused_resize_losses(xs[i], ys[i], ce, ignore_index=ign, return_stats=True)
            r['loss_ce'].backward()
            return r

        ring = GraphRing(step, R, dev, world, dist, every=ar_every)
        ms = ring.timed(K, W)
        px = n_loc * world * Hh * Ww
        algo = 2 * n_loc * world * Cn * Hh * Ww * 4 + px * 8          # single pass: read + write logits, read labels
        a = algo / (ms * 1e-3) / 1e9
        out[mode] = {'images_per_gpu': n_loc, 'global_batch': n_loc * world, 'ms': ms, 'mpix_s': px / ms / 1e3, 'n_gpus': world,
                     'ring_sets': R,
                     'roofline': {'bound': 'hbm', 'achieved': a, 'peak': peak * world, 'unit': 'GB/s', 'frac': a / (peak * world),
                                  'peak_kind': peak_kind}}
        del ring, xs, ys
        torch.cuda.empty_cache()
    out['note'] = ('ce_bulk_kernel + finalize per step from a CUDA graph; the steps\' statistics vectors all-reduced once per ring '
                   'cycle on a side stream; strong = 32/G images per GPU, weak = 32 per GPU')
    return out


def c5_sharded(B, D, dist, dev, rank, world, peak, peak_kind):
    """BASELINE config 5 across ranks: the 500-image sweep sharded by image (ceil split). Per sweep and rank: ONE
    b200seg_confusion_labels launch that accumulates the (3, C) int64 totals of its image list inside the kernel, then ONE
    int64 all-reduce of those 456 bytes. The image table of the (unchanging) buffers is uploaded once. Timed with CUDA
    events between barriers, max over ranks; pixels counted over all 500 images."""
    Cn, n_img = 19, 500
    lo, hi = D.shard_range(n_img, rank, world)
    g = torch.Generator(device=dev).manual_seed(555 + rank)
    gts = [make_labels((1, 1024, 2048), Cn, 500 + i, 255, device=dev)[0].float() for i in range(4)]
    pred_base = [torch.randint(0, Cn, (1024, 2048), generator=g, device=dev) for _ in range(8)]
    preds = [pred_base[i % 8].clone() for i in range(lo, hi)]
    gt_all = [gts[i % 4].clone() for i in range(lo, hi)]
    table = B.prepare_images(preds, gt_all, Cn)

    def sweep():
        tot = B.area_totals_device(table, None, Cn, 255)
        return D.all_reduce_areas({'areas': tot})['areas']

    for _ in range(3):
        tot = sweep()
    torch.cuda.synchronize()
    dist.barrier()
    iters = 10
    s_ev, e_ev = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s_ev.record()
    for _ in range(iters):
        tot = sweep()
    e_ev.record()
    torch.cuda.synchronize()
    dist.barrier()
    t = torch.tensor([s_ev.elapsed_time(e_ev) / iters], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    px = n_img * 1024 * 2048
    label_px = int(tot[2].sum().item())       # all-reduced label-area total: every non-ignored pixel of all 500 images
    a = px * 12 / (ms * 1e-3) / 1e9
    return {'images': n_img, 'images_per_rank': hi - lo, 'pixels': px, 'ms': ms, 'mpix_s': px / ms / 1e3, 'n_gpus': world,
            'roofline': {'bound': 'hbm', 'achieved': a, 'peak': peak * world, 'unit': 'GB/s', 'frac': a / (peak * world),
                         'peak_kind': peak_kind},
            'all_reduced_label_pixels': label_px,
            'note': 'sharded by image over the ranks; per sweep one b200seg_confusion_labels launch per rank (in-kernel (3,C) '
                    'totals) + one int64 all-reduce of the area totals'}


def extra_workloads(B, _lib, dev, peak, peak_kind):
    """BASELINE configs 3, 4, 5 on one GPU: Mpix/s and HBM-roofline fraction of each kernel group."""
    out = {}

    def roof(algo_bytes, ms):
        a = algo_bytes / (ms * 1e-3) / 1e9
        return {'bound': 'hbm', 'achieved': a, 'peak': peak, 'unit': 'GB/s', 'frac': a / peak, 'peak_kind': peak_kind}

    def bench_losses(name, shape, dtype, losses, label_dtype=torch.int64, iters=20, single=False, plan='', graph=False):
        n, c, hh, ww = shape
        s = 4 if dtype == torch.float32 else 2
        xs = [make_logits(shape, 300 + i, dtype=dtype, device=dev).requires_grad_(True) for i in range(2)]
        ys = [make_labels((n, hh, ww), c, 300 + i, 255, device=dev, dtype=label_dtype).unsqueeze(1) for i in range(2)]
        L = ys[0].element_size()

        def fwd(i):
            with torch.no_grad():
                B.fused_resize_losses(xs[i & 1], ys[i & 1], losses, ignore_index=255)

        def fwdbwd(i):
            x = xs[i & 1]
            x.grad = None
            r = B.fused_resize_losses(x, ys[i & 1], losses, ignore_index=255)
            tot = None
            for k, v in r.items():
                if k.startswith('loss'):
                    tot = v if tot is None else tot + v
            tot.backward()

        for i in range(3):
            fwd(i); fwdbwd(i)
        if graph:   # steps of a few tens of microseconds: replayed from CUDA graphs (2 input sets), as the headline is
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for i in range(2):
                    fwd(i); fwdbwd(i)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            gf, gfb = [], []
            for i in range(2):
                for fn, lst in ((fwd, gf), (fwdbwd, gfb)):
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g):
                        fn(i)
                    lst.append(g)
            for i in range(4):
                gf[i & 1].replay(); gfb[i & 1].replay()
            ms_f = timed_events(lambda i: gf[i & 1].replay(), iters)
            ms_fb = timed_events(lambda i: gfb[i & 1].replay(), iters)
        else:
            ms_f = timed_events(fwd, iters)
            ms_fb = timed_events(fwdbwd, iters)
        px = n * hh * ww
        elems = n * c * hh * ww
        algo_f = elems * s + px * L
        algo_fb = (2 * elems * s + px * L) if single else (3 * elems * s + 2 * px * L)
        out[name] = {'shape': list(shape), 'dtype': str(dtype).replace('torch.', ''), 'pixels': px,
                     'fwd': dict(ms=ms_f, mpix_s=px / ms_f / 1e3, roofline=roof(algo_f, ms_f)),
                     'fwd_bwd': dict(ms=ms_fb, mpix_s=px / ms_fb / 1e3, roofline=roof(algo_fb, ms_fb),
                                     plan=plan),
                     'algorithmic_bytes': {'fwd': algo_f, 'fwd_bwd': algo_fb}}
        del xs, ys
        torch.cuda.empty_cache()

    cw = torch.linspace(0.5, 1.5, 150).tolist()
    bench_losses('C3_ade20k_bf16_ce_dice', (16, 150, 512, 512), torch.bfloat16,
                 [B.CrossEntropyLoss(class_weight=cw), B.DiceLoss(loss_weight=3.0)], iters=10,
                 plan='cs_fwd_kernel + finalize; cs_bwd_kernel (class-sliced tensor-map TMA pipeline: one read of the logits per direction)')
    bench_losses('C4_voc_fp32_ce', (32, 21, 512, 512), torch.float32, B.CrossEntropyLoss(), iters=20, single=True,
                 plan='ce_bulk_kernel (cp.async.bulk load warp / consumers / store warp): forward+backward in one pass; fwd = its forward-only form')
    bench_losses('C3_ade20k_bf16_ce_only', (16, 150, 512, 512), torch.bfloat16, B.CrossEntropyLoss(class_weight=cw), iters=10,
                 plan='ce_fwd_kernel (saves lse) + ce_bwd_kernel')
    ce2 = B.CrossEntropyLoss()
    ce2.single_pass = False
    bench_losses('C4_voc_fp32_ce_two_pass', (32, 21, 512, 512), torch.float32, ce2, iters=10,
                 plan='ce_bulk_kernel<forward only> (saves lse) + ce_bwd_kernel')

    # ---- row f1: the sigmoid path the shipped default config runs (use_sigmoid=True, 2 classes, configs/network/deeplabv3):
    # one-hot expansion + BCE-with-logits in one stream per direction (csrc/loss_bce.cu)
    try:
        bench_losses('F1_sigmoid_ce_2class', (32, 2, 512, 512), torch.float32, B.CrossEntropyLoss(use_sigmoid=True), iters=20,
                     single=True, graph=True,
                     plan='bce_kernel<fused>: loss, reduced scalar and gradient in one pass (one read and one write of the '
                          'logits); CUDA-graph replay')
    except Exception as e:
        out['F1_sigmoid_ce_2class'] = {'error': repr(e)}

    # ---- resize-fused CE beyond the headline shape: align_corners=True (thread-per-cell kernel, csrc/loss_upgen.cuh) and
    # 150 classes at 1/8 resolution (C > 32: the class-tiled plan of the same file); both deterministic — there is no
    # atomicAdd backward. Each step is replayed from a CUDA graph, as the headline is.
    def bench_up(name, shape, size, ac, iters=20, plan=''):
        n, c, hh, ww = shape
        xs = [make_logits(shape, 700 + i, device=dev).requires_grad_(True) for i in range(2)]
        ys = [make_labels((n,) + size, c, 700 + i, 255, device=dev).unsqueeze(1) for i in range(2)]
        ce = B.CrossEntropyLoss()

        def fb(i):
            x = xs[i & 1]
            x.grad = None
            B.fused_resize_losses(x, ys[i & 1], ce, align_corners=ac, ignore_index=255)['loss_ce'].backward()

        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for i in range(2):
                fb(i)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graphs = []
        for i in range(2):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                fb(i)
            graphs.append(g)
        for i in range(3):
            graphs[i & 1].replay()
        ms = timed_events(lambda i: graphs[i & 1].replay(), iters)
        px = n * size[0] * size[1]
        out[name] = {'logits': list(shape), 'label_hw': list(size), 'align_corners': ac, 'pixels': px,
                     'fwd_bwd': dict(ms=ms, mpix_s=px / ms / 1e3, ns_per_px_class=ms * 1e6 / (px * c), plan=plan)}
        del xs, ys
        torch.cuda.empty_cache()

    try:
        bench_up('C2_align_corners_false', (8, 19, 64, 128), (512, 1024), False,
                 plan='the headline shape measured the same way as the two lines below (2 input sets, CUDA-graph replay)')
        bench_up('C2_align_corners_true', (8, 19, 64, 128), (512, 1024), True,
                 plan='up_gen_kernel (thread per cell, any ratio) + up_combine + finalize')
        bench_up('C2_like_c150_64to512', (8, 150, 64, 64), (512, 512), False, iters=10,
                 plan='class-tiled thread-per-cell plan: up_gen_kernel (forward, all classes) + up_gen_bwd_tile_kernel x 5 tiles + up_combine')
    except Exception as e:
        out['C2_align_corners_true'] = {'error': repr(e)}

    # ---- "next" row f4': LovaszLoss at the config-2 label resolution (19 class segments of 4 M pixels each)
    try:
        xl = make_logits((8, 19, 512, 1024), 350, device=dev).requires_grad_(True)
        yl = make_labels((8, 512, 1024), 19, 350, 255, device=dev)
        lov = B.LovaszLoss(reduction='none')

        def lv_fwd(i):
            with torch.no_grad():
                lov(xl, yl, ignore_index=255)

        def lv_fb(i):
            xl.grad = None
            lov(xl, yl, ignore_index=255).backward()

        lv_fwd(0); lv_fb(0)
        a, b = timed_events(lv_fwd, 5), timed_events(lv_fb, 5)
        px = 8 * 512 * 1024
        out['lovasz_softmax_cityscapes_shape'] = {
            'shape': [8, 19, 512, 1024], 'dtype': 'float32', 'pixels': px,
            'fwd': dict(ms=a, mpix_s=px / a / 1e3), 'fwd_bwd': dict(ms=b, mpix_s=px / b / 1e3),
            'plan': 'ce_fwd_kernel (lse) + for all classes at once: lovasz_keys_kernel (keys + digit histograms), '
                    'hand-written segmented radix sort (lov_sort_pass_kernel x 4: ballot ranking, decoupled look-back), '
                    'lovasz_count/tilescan/grad kernels; lovasz_finalize_kernel; lovasz_bwd_kernel',
            'note': 'sort-bound (4 digit passes over 80 M key/index pairs, instruction-issue bound); no HBM roofline is '
                    'claimed for it'}
        del xl, yl
        torch.cuda.empty_cache()
    except Exception as e:   # the extra line must never cost the headline
        out['lovasz_softmax_cityscapes_shape'] = {'error': repr(e)}

    # ---- C5: mIoU sweep. (i) 500 label maps 1024x2048 int64 + float32 gt, one launch; (ii) from logits, all 500 images
    Cn, n_img = 19, 500
    g = torch.Generator(device=dev).manual_seed(555)
    gts = [make_labels((1, 1024, 2048), Cn, 500 + i, 255, device=dev)[0].float() for i in range(4)]
    gt_all = [gts[i % 4] for i in range(n_img)]
    pred_base = [torch.randint(0, Cn, (1024, 2048), generator=g, device=dev) for _ in range(8)]
    preds = [pred_base[i % 8].clone() for i in range(n_img)]      # 500 distinct int64 buffers: 8.4 GB
    gt_all = [t.clone() for t in gt_all]                          # 500 distinct fp32 buffers: 2.1 GB

    def sweep(i):
        B.areas_device(preds, gt_all, Cn, 255)

    sweep(0)
    ms = timed_events(sweep, 5)
    px = n_img * 1024 * 2048
    out['C5i_miou_label_maps'] = {'images': n_img, 'pixels': px, 'ms': ms, 'mpix_s': px / ms / 1e3,
                                  'roofline': roof(px * 12, ms), 'algorithmic_bytes': px * 12,
                                  'note': 'one b200seg_confusion_labels launch for the 500-image list, per-image areas as '
                                          'intersect_and_union returns them, image table built and uploaded inside the timed '
                                          'call; random predictions (worst case for the histogram), blocky ground truth'}
    # the same sweep as the evaluator runs it over a prepared list: table built once, (3,C) totals accumulated in the kernel
    tab5 = B.prepare_images(preds, gt_all, Cn)

    def sweep_t(i):
        B.area_totals_device(tab5, None, Cn, 255)

    sweep_t(0)
    ms_t = timed_events(sweep_t, 5)
    out['C5i_miou_label_maps']['totals_prepared'] = {'ms': ms_t, 'mpix_s': px / ms_t / 1e3, 'roofline': roof(px * 12, ms_t),
                                                      'note': 'prepare_images once + area_totals_device: the kernel alone'}
    del tab5
    del preds, pred_base
    torch.cuda.empty_cache()
    n_l = 500
    lbase = [make_logits((1, Cn, 1024, 2048), 900 + i, device=dev) for i in range(4)]
    logits = [lbase[i % 4].clone() for i in range(n_l)]           # 500 x 159 MB = 79.7 GB of distinct buffers
    gl = gt_all[:n_l]

    def sweep2(i):
        B.areas_device(logits, gl, Cn, 255, from_logits=True)

    sweep2(0)
    ms = timed_events(sweep2, 3)
    px = n_l * 1024 * 2048
    out['C5ii_miou_from_logits'] = {'images': n_l, 'pixels': px, 'ms': ms, 'mpix_s': px / ms / 1e3,
                                    'roofline': roof(px * (Cn * 4 + 4), ms), 'algorithmic_bytes': px * (Cn * 4 + 4),
                                    'note': 'all 500 images (79.7 GB of fp32 logits), fused arg-max + areas, one launch'}
    tab5 = B.prepare_images(logits, gl, Cn, from_logits=True)

    def sweep2_t(i):
        B.area_totals_device(tab5, None, Cn, 255)

    sweep2_t(0)
    ms_t = timed_events(sweep2_t, 3)
    out['C5ii_miou_from_logits']['totals_prepared'] = {'ms': ms_t, 'mpix_s': px / ms_t / 1e3,
                                                        'roofline': roof(px * (Cn * 4 + 4), ms_t),
                                                        'note': 'prepare_images once + area_totals_device: the kernel alone'}
    del logits, lbase, tab5
    torch.cuda.empty_cache()

    # ---- "next" row f2: the validation rescale fused into the arg-max — logits at 1/8 resolution (1,19,128,256) against
    # 1024x2048 ground truth, 500 images, one launch; the (1,19,1024,2048) rescaled logits are never written
    try:
        lo = [make_logits((1, Cn, 128, 256), 950 + i, device=dev) for i in range(8)]
        lows = [lo[i % 8].clone() for i in range(n_l)]
        tabr = B.prepare_images(lows, gl, Cn, from_logits=True)

        def sweep3(i):
            B.area_totals_device(tabr, None, Cn, 255)

        sweep3(0)
        ms = timed_events(sweep3, 3)
        px = n_l * 1024 * 2048
        algo = px * 4 + n_l * Cn * 128 * 256 * 4
        out['C5_resized_lowres_logits'] = {'images': n_l, 'pixels': px, 'ms': ms, 'mpix_s': px / ms / 1e3,
                                           'roofline': roof(algo, ms), 'algorithmic_bytes': algo,
                                           'note': 'fused bilinear rescale (ATen arithmetic, bit-exact) + arg-max + areas from 1/8-resolution '
                                                   'logits; issue bound (19 interpolations per output pixel), not HBM bound'}
        del lows, lo, tabr
    except Exception as e:
        out['C5_resized_lowres_logits'] = {'error': repr(e)}
    del gl, gt_all
    torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=200)
    ap.add_argument('--warmup', type=int, default=20)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--no-extras', action='store_true')
    ap.add_argument('--allreduce-every', type=int, default=None,
                    help='steps per all-reduce of the logging statistics under torchrun (default: the ring size, 8; 1 = every step)')
    args = ap.parse_args()
    if args.impl == 'reference':
        reference_main(args)
    else:
        b200_main(args)


if __name__ == '__main__':
    main()

#!/usr/bin/env python
"""Headline benchmark: supernet search-step images/sec @224^2, per-GPU bs 128 (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one SEARCH UNIT = two iterations of train_w_arch (reference train_search.py:366-426):
2 bi-sampled w-steps (4 single-path fwd+bwd, SGD) + 1 alpha-step (8-path fwd+bwd, Adam, log_softmax
renorm) on the full supernet (configs[1]): synthetic 3x224x224 fp32, bs 128 per GPU, 100 classes,
T=5, lambda_lat=0.1, target_lat=15, the reference's latency_gpu LUT.  images/sec counts train images
(2*bs*world per unit).  Prints ONE JSON line on rank 0.

  value        device-timed throughput, batches already resident in HBM
  e2e          same through the public API with pinned-host batches copied H2D inside the timed
               region every step and the losses read back (D2H)
  roofline     the kernel with the largest share of the step, timed with CUDA events on its stream
               (library event profiler) over the same K units: algorithmic bytes / time vs the
               measured HBM peak
  mixedop_roofline  SURVEY 8(d)'s micro-benchmark: alpha-mode fwd+bwd of the 18 MixedOPs (one supernet's worth, body
               executor) timed with CUDA events; ALGORITHMIC GB/s (1.694 GB) and TFLOP/s (1.27 TF) and the fraction of the
               binding bound (tf32x3 tensor floor)
  cpu_baseline the oracle port (torch CPU, all host threads) on a bounded sample (search units at bs 32, the
               reference's default batch; bs 128 does not fit host memory -- config.ref_batch states it)
  gpu_baseline the same oracle port (= the reference's own PyTorch code path: cuDNN / cuBLAS / ATen) on THIS GPU at
               bs 128, with TF32 off (true fp32, what parity requires) and on (PyTorch's cuDNN default)

--impl reference times that same CPU oracle port (the reference is pure Python/PyTorch and does
not travel to the GPU box) for the same metric/config, one search unit (at bs 32) per step.
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
sys.dont_write_bytecode = True

BS = 128
REF_BS = 32        # CPU arm: the reference's default --batch_size (train_search.py:44); bs 128 needs ~70 GB of host RAM
METRIC = 'supernet_search_step_images_per_sec'
UNIT = 'images/s'


EXTRA_WARMUP = 2      # untimed units after the requested warm-up (allocator / optimiser state settle)


def target_lat_for(world):
    """BASELINE.json: target_lat 15.0 for the single-GPU configs (2, 3), 18.0 for the 8-GPU search (config 4)."""
    return 18.0 if world >= 8 else 15.0


def base_config(world):
    tl = target_lat_for(world)
    return {'workload': 'full supernet (Network) search unit = 2 bi-sampled w-steps + 1 alpha-step, '
                        'synthetic 3x224x224 fp32, bs %d per GPU, target_lat %.1f, latency_gpu LUT' % (BS, tl),
            'per_gpu_batch': BS, 'global_batch': BS * world, 'image': '3x224x224', 'num_classes': 100,
            'T': 5.0, 'lambda_lat': 0.1, 'target_lat': tl, 'parallelism': 'dp%d' % world,
            'ref_batch': REF_BS,      # batch of the CPU arms (--impl reference, cpu_baseline): per-image throughput is reported
            'l2_policy': 'per-step working set (>20 GB of activations) far exceeds the 126 MB L2; no flush needed'}


# ------------------------------------------------------------------------------------------------
_SAMPLER_SRC = r"""
import os, sys, time
import pynvml as nv
nv.nvmlInit()
h = nv.nvmlDeviceGetHandleByIndex(int(sys.argv[1]))
print('max %f' % float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)), flush=True)
while True:
    print('%f %d' % (float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)), int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(h))), flush=True)
    time.sleep(0.2)
"""


class ClockSampler(object):
    """SM clock / throttle reasons during the timed region (B200_PROFILING.md recipe), sampled through NVML every 200 ms by a
    CHILD PROCESS: a sampling thread inside the bench process competes with the launching thread for the interpreter lock
    (measured: value leg 4 % slower than the same units without it), and nvidia-smi in a loop is heavier still."""

    def __init__(self, index):
        vis = os.environ.get('CUDA_VISIBLE_DEVICES', '')
        ids = [int(v) for v in vis.split(',') if v.strip().isdigit()]       # NVML enumerates physical devices
        self.phys = ids[index] if index < len(ids) else index
        self.proc, self.rows, self.max_mhz = None, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen([sys.executable, '-c', _SAMPLER_SRC, str(self.phys)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            first = self.proc.stdout.readline()          # wait until the child is sampling
            if first.startswith('max'):
                self.max_mhz = float(first.split()[1])
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return
        try:
            self.proc.terminate()
            out, _ = self.proc.communicate(timeout=5)
            for line in out.splitlines():
                a = line.split()
                if len(a) == 2:
                    self.rows.append((float(a[0]), int(a[1])))
        except Exception:
            pass

    def summary(self):
        sm = sorted(r[0] for r in self.rows)
        bits = {'hw_slowdown': 0x8, 'hw_thermal_slowdown': 0x40, 'sw_thermal_slowdown': 0x20, 'sw_power_cap': 0x4}
        reasons = [n for n, b in bits.items() if any(r[1] & b for r in self.rows)]
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': self.max_mhz, 'reasons': reasons,
                'samples': len(self.rows)}


def measured_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


# ------------------------------------------------------------------------------------------------
def run_reference(args, rank, world, emit=print):
    """CPU oracle port, one search unit per step on a bounded sample (bs REF_BS)."""
    if rank != 0:
        return
    import torch
    from oracle import port
    from tests import golden_inputs as gi
    from tfnas_b200 import config
    torch.set_num_threads(os.cpu_count() or 1)
    lut = gi.load_lut()
    mcs = config.get_mc_num_dddict(config.mc_mask_dddict)
    P = port.init_params(mcs, seed=2)
    g = torch.Generator().manual_seed(2)
    batches = [(torch.randn(REF_BS, 3, 224, 224, generator=g), torch.randint(0, 100, (REF_BS,), generator=g))
               for _ in range(2)]
    tl = target_lat_for(max(world, 1))
    state = {}
    for _ in range(args.warmup):
        port.search_unit_cpu(P, mcs, lut, batches, 5.0, tl, 0.1, state=state)
    t0 = time.time()
    n_img = 0
    for _ in range(args.steps):
        n_img += port.search_unit_cpu(P, mcs, lut, batches, 5.0, tl, 0.1, state=state)
    dt = time.time() - t0
    val = n_img / dt
    sample = 'search unit at bs %d on CPU (%d threads), fp32, same supernet / LUT / losses' % (REF_BS, torch.get_num_threads())
    emit(json.dumps({'impl': 'reference', 'metric': METRIC, 'value': val, 'unit': UNIT, 'n_gpus': args.gpus,
                      'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1000 * dt / args.steps,
                      'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
                      'data': 'synthetic', 'config': base_config(max(world, 1)),
                      'cpu_baseline': {'value': val, 'unit': UNIT, 'cores': torch.get_num_threads(), 'kind': 'port',
                                       'sample': sample},
                      'e2e': {'value': val, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
                      'gpu_launches': 0}))


def cpu_baseline_sample():
    import torch
    from oracle import port
    from tests import golden_inputs as gi
    from tfnas_b200 import config
    nthr = os.cpu_count() or 1
    torch.set_num_threads(nthr)
    lut = gi.load_lut()
    mcs = config.get_mc_num_dddict(config.mc_mask_dddict)
    P = port.init_params(mcs, seed=2)
    g = torch.Generator().manual_seed(2)
    batches = [(torch.randn(REF_BS, 3, 224, 224, generator=g), torch.randint(0, 100, (REF_BS,), generator=g))
               for _ in range(2)]
    state = {}
    port.search_unit_cpu(P, mcs, lut, batches, 5.0, 15.0, 0.1, state=state)       # warm-up
    t0 = time.time()
    n = 0
    for _ in range(2):
        n += port.search_unit_cpu(P, mcs, lut, batches, 5.0, 15.0, 0.1, state=state)
    dt = time.time() - t0
    return {'value': n / dt, 'unit': UNIT, 'cores': nthr, 'kind': 'port',
            'sample': '2 search units at bs %d (oracle/port.py, torch CPU fp32, %d threads), %.1f s' % (REF_BS, nthr, dt)}


def gpu_baseline_sample(dev):
    """The reference's own code path (oracle port = the same PyTorch ops: cuDNN convs, ATen BN / elementwise, torch.autograd)
    on this GPU at bs 128: search units with TF32 off (true fp32) and on (PyTorch's cuDNN default).  SURVEY 8(d) last row."""
    import torch
    from oracle import port
    from tests import golden_inputs as gi
    from tfnas_b200 import config
    lut = gi.load_lut()
    mcs = config.get_mc_num_dddict(config.mc_mask_dddict)
    g = torch.Generator().manual_seed(2)
    batches = [(torch.randn(BS, 3, 224, 224, generator=g).to(dev), torch.randint(0, 100, (BS,), generator=g).to(dev))
               for _ in range(2)]
    out = {'unit': UNIT, 'kind': 'port on cuda (stock PyTorch ops)', 'batch': BS, 'units_timed': 2}
    saved = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    try:
        for name, tf32 in (('tf32_off', False), ('tf32_on', True)):
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = tf32
            P = {k: v.to(dev) for k, v in port.init_params(mcs, seed=2).items()}
            state = {}
            port.search_unit_cpu(P, mcs, lut, batches, 5.0, 15.0, 0.1, state=state)
            torch.cuda.synchronize()
            t0 = time.time()
            n = 0
            for _ in range(2):
                n += port.search_unit_cpu(P, mcs, lut, batches, 5.0, 15.0, 0.1, state=state)
            torch.cuda.synchronize()
            out[name] = n / (time.time() - t0)
            del P, state
            torch.cuda.empty_cache()
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = saved
    return out


# SURVEY.md 8(d), N = 128, fp32, initial widths, sums over the 18 MixedOPs
ALG = dict(bytes_fwd=685.0e6, bytes_bwd=1009.3e6, gf_1x1=570.86, gf_dw=60.56, gf_se=3.52, gf_elem=47.30)


def mixedop_roofline(net, dev, reps, peak_gbs):
    """alpha-mode fwd+bwd of all 18 MixedOPs (+ the 6 sinks) at bs 128 through the body executor, CUDA events."""
    import torch
    from tfnas_b200.model_search import NoisePlan, injected
    g = torch.Generator().manual_seed(5)
    xs = [torch.randn(BS, 16, 112, 112, generator=g).to(dev).requires_grad_(True) for _ in range(2)]
    G = torch.randn(BS, 320, 7, 7, generator=g).to(dev)
    for p in net.weight_parameters():
        p.requires_grad = False
    for p in net.arch_parameters():
        p.requires_grad = True

    def once(x):
        out, lat = net._body(x, False, 'max')
        torch.autograd.backward([out, lat], [G, torch.ones_like(lat)])
        x.grad = None
    for i in range(2):
        once(xs[i % 2])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        once(xs[i % 2])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    for p in net.arch_parameters():
        p.grad = None
    nbytes = ALG['bytes_fwd'] + ALG['bytes_bwd']
    flops = 2.0 * (ALG['gf_1x1'] + ALG['gf_dw'] + ALG['gf_se']) * 1e9
    pk = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json'))) if os.path.exists(os.path.join(ROOT, 'MEASURED_PEAKS.json')) else {}
    bf16 = float(pk.get('bf16_tflops', 1617.8))
    floor_bytes = nbytes / (peak_gbs * 1e9) * 1e3
    floor_tensor = 2.0 * ALG['gf_1x1'] * 1e9 * 3.0 / (0.5 * bf16 * 1e12) * 1e3      # tf32 = 1/2 bf16 rate, 3-term split
    floor_fp32 = 2.0 * (ALG['gf_dw'] + ALG['gf_se'] + ALG['gf_elem']) * 1e9 / 74e12 * 1e3
    floor = max(floor_bytes, floor_tensor, floor_fp32)
    tp = os.path.join(ROOT, 'profiles', 'ncu_tensor_pipe.json')
    return {'what': 'alpha-mode fwd+bwd of the 18 MixedOPs at bs 128 (SURVEY 8d micro-benchmark unit, dx of the first MixedOP included)',
            'ms': ms, 'algorithmic_bytes': nbytes, 'algorithmic_flops': flops,
            'achieved_GBps': nbytes / ms / 1e6, 'hbm_frac_algorithmic': nbytes / ms / 1e6 / peak_gbs,
            'achieved_TFLOPs': flops / ms / 1e9,
            'floors_ms': {'bytes': floor_bytes, 'tensor_tf32x3': floor_tensor, 'fp32_pipe': floor_fp32},
            'binding_bound': 'tensor (tf32 x3 split the 1e-3 parity bar requires)' if floor == floor_tensor else 'other',
            'binding_bound_frac': floor / ms,
            'tensor_pipe_pct_ncu': json.load(open(tp)) if os.path.exists(tp) else None}


# ------------------------------------------------------------------------------------------------
def run_b200(args, rank, local, world, emit=print):
    import torch
    import torch.distributed as dist
    import torch.nn as nn
    from tests import golden_inputs as gi
    from tfnas_b200 import _lib, config, model_search
    from tfnas_b200.model_search import Network
    from tfnas_b200.parallel import GradSync, SearchParallel
    from tfnas_b200 import search_loop
    from tfnas_b200.search_loop import DevicePrefetcher, alpha_step, make_optimizers, w_step

    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device; the B200 path has no CPU fallback')
    _lib.load()
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    torch.manual_seed(2)
    model_search.seed_noise(2)           # identical sampling on every rank
    lut = gi.load_lut()
    mcs = config.get_mc_num_dddict(config.mc_mask_dddict)
    net = Network(100, mcs, lut)
    net.set_temperature(5.0)
    model = SearchParallel(net).to(dev).train()
    from tfnas_b200.step import FusedCrossEntropy
    criterion = FusedCrossEntropy().to(dev)
    opt_w, opt_a = make_optimizers(net)
    sync = GradSync()
    g = torch.Generator().manual_seed(2 + rank)
    npool = 3
    host = [(torch.randn(BS, 3, 224, 224, generator=g).pin_memory(), torch.randint(0, 100, (BS,), generator=g).pin_memory())
            for _ in range(npool)]
    pool = [(x.to(dev), t.to(dev)) for x, t in host]

    def host_batches(steps):
        """the pinned host batches of `steps` units in the order the units consume them"""
        for i in range(steps):
            for it in range(2):
                yield host[(2 * i + it) % npool]
                if it % 2 == 0:
                    yield host[(2 * i + it + 1) % npool]

    def unit(i, feed=None, overlap=True):
        """feed: iterator of device batches copied from pinned host memory (e2e); None: batches resident in HBM"""
        losses = []
        for it in range(2):
            x, t = next(feed) if feed is not None else pool[(2 * i + it) % npool]
            lw, _ = w_step(model, x, t, criterion, opt_w, 5.0, sync, bisample=True, overlap=overlap)
            losses.append(lw)
            if it % 2 == 0:
                xa, ta = next(feed) if feed is not None else pool[(2 * i + it + 1) % npool]
                # the attribution leg (overlap=False) runs the alpha step eagerly: a graph replay carries no profiler events
                la, ll = alpha_step(model, xa, ta, criterion, opt_a, target_lat_for(world), 0.1, 5.0, sync,
                                    graph=None if overlap else False)
                losses += [la, ll]
        if feed is not None:
            # D2H read of the unit's four losses: an async copy into pinned memory + an event; the VALUES of unit i are
            # consumed while unit i+1 runs (and the last unit's before the timed region closes), so the read-back does not
            # drain the stream between units -- how a training loop logs its meters
            slot = readback[len(pending) % 2]
            slot.copy_(torch.stack([v.detach().float() for v in losses]), non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
            pending.append((slot, ev))
            if len(pending) > 1:
                pslot, pev = pending[-2]
                pev.synchronize()
                loss_log.append([float(v) for v in pslot])
        return losses

    readback = [torch.empty(4, dtype=torch.float32).pin_memory() for _ in range(2)]
    pending, loss_log = [], []

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    host_ms = [0.0, 0.0]

    def timed(from_host, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = _lib.launch_count() + search_loop.GRAPH_LAUNCHES[0]
        e0.record()
        h0 = time.perf_counter()
        # e2e: every batch comes from pinned host memory inside the timed region (H2D on a copy stream, one batch ahead of
        # the compute: search_loop.DevicePrefetcher, what train_search.py's loops use) and the losses are read back per unit
        feed = iter(DevicePrefetcher(host_batches(steps), dev)) if from_host else None
        b0 = net.__dict__.get('host_blocked_s', 0.0)
        del pending[:]
        for i in range(steps):
            unit(i, feed)
        if from_host and pending:          # the last unit's losses, still inside the timed region
            pending[-1][1].synchronize()
            loss_log.append([float(v) for v in pending[-1][0]])
        host_ms[0] = (time.perf_counter() - h0) * 1e3 / steps     # host wall time to enqueue a unit ...
        host_ms[1] = (net.__dict__.get('host_blocked_s', 0.0) - b0) * 1e3 / steps   # ... of which blocked in the one D2H read
        # per unit (the new log_alphas the 'gumbel' sampling of the next w-step needs: it waits for the alpha update)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms, _lib.launch_count() + search_loop.GRAPH_LAUNCHES[0] - l0

    for i in range(args.warmup):
        unit(i)
    if not args.profile_only:
        # two more untimed units: every step samples new candidates, so the caching allocator (workspace sizes) and the
        # optimiser (momentum buffers of first-seen candidates) keep growing for a few units; a cudaMalloc inside the timed
        # region showed up once as a 15 % outlier of `value` next to an unaffected `e2e`
        for i in range(EXTRA_WARMUP):
            unit(args.warmup + i)
    if args.profile_only:      # under ncu: just run K more units and leave (numbers under a profiler are not bench values)
        for i in range(args.steps):
            unit(i)
        torch.cuda.synchronize()
        return
    sampler = ClockSampler(local)
    sampler.start()
    ms, launches = timed(False, args.steps)
    host_enqueue, host_blocked = host_ms[0], host_ms[1]
    sampler.stop()
    unit(0, iter(DevicePrefetcher(host_batches(1), dev)))     # warm the H2D path
    ms_e2e, _ = timed(True, args.steps)
    # pinned host -> device bandwidth of THIS box, idle GPU: the e2e leg needs 3 batches (231 MB) per unit, i.e. 2.5 GB/s at
    # 92 ms per unit -- boxes were seen where pinned copies ran at exactly that rate and e2e sat at 94.0 ms whatever the
    # kernels did, so the number is reported next to e2e
    xh = host[0][0]
    xd = torch.empty_like(pool[0][0])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    xd.copy_(xh, non_blocking=True)
    e0.record()
    for _ in range(3):
        xd.copy_(xh, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    h2d_gbps = 3 * xh.numel() * 4 / (e0.elapsed_time(e1) / 1e3) / 1e9
    del xd
    # per-kernel attribution over the same K units (CUDA events on the launching stream).  The units run with their passes
    # in sequence on ONE stream for this leg: with the two sampled passes and the weight-gradient side streams concurrent
    # (as in the timed legs) an event pair brackets a kernel that shares the SMs with its neighbours' kernels.
    prev_side = _lib.load().tfnas_config_side_stream(0)
    _lib.prof_enable(True)
    barrier()
    for i in range(args.steps):
        unit(i, overlap=False)
    torch.cuda.synchronize()
    recs = _lib.prof_collect()
    _lib.prof_enable(False)
    _lib.load().tfnas_config_side_stream(1 if prev_side != 0 else 0)

    imgs = 2 * BS * world * args.steps
    value = imgs / (ms / 1000.0)
    e2e = imgs / (ms_e2e / 1000.0)
    peak, peak_src = measured_peaks()
    tot_ms = sum(r['ms'] for r in recs) or 1.0
    recs.sort(key=lambda r: -r['ms'])
    top = recs[0]
    per_launch_ms = top['ms'] / top['launches']
    achieved = top['bytes'] / top['launches'] / per_launch_ms / 1e6          # GB/s
    traffic = None
    tp = os.path.join(ROOT, 'profiles', 'ncu_traffic.json')
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get(top['name'])
    mop = mixedop_roofline(net, dev, 6, peak) if rank == 0 else None
    for p_ in net.parameters():
        p_.grad = None
    roofline = {'kernel': top['name'], 'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
                'frac': achieved / peak, 'traffic': traffic, 'peak_source': peak_src,
                'timed': 'CUDA events per launch, passes in sequence on one stream (no concurrent neighbour kernels)',
                'bytes_model': 'materialised tensors this kernel reads + writes (DESIGN.md section 5), not SURVEY 8(d) compulsory bytes',
                'algorithmic_frac': mop['hbm_frac_algorithmic'] if mop else None,
                'share_of_step': top['ms'] / tot_ms, 'launches_per_step': top['launches'] / args.steps,
                'avg_launch_ms': per_launch_ms, 'achieved_tflops': top['flops'] / top['launches'] / per_launch_ms / 1e9,
                'kernel_ms_per_step': tot_ms / args.steps,
                'kernels': [{'name': r['name'], 'share': r['ms'] / tot_ms, 'ms_per_step': r['ms'] / args.steps,
                             'GBps': r['bytes'] / r['ms'] / 1e6, 'TFLOPs': r['flops'] / r['ms'] / 1e9,
                             'launches_per_step': r['launches'] / args.steps} for r in recs[:48]]}
    if rank != 0:
        return
    h2d = 3 * (BS * 3 * 224 * 224 * 4 + BS * 8)
    line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f32', 'data': 'synthetic', 'config': base_config(world),
            'e2e': {'value': e2e, 'unit': UNIT, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': 4 * 4,
                    'ms_per_step': ms_e2e / args.steps, 'h2d_pinned_GBps_idle': h2d_gbps},
            'gpu_launches': launches, 'launches_per_step': launches / args.steps,
            'host_enqueue_ms_per_step': host_enqueue, 'host_blocked_on_gpu_ms_per_step': host_blocked,
            'host_busy_ms_per_step': host_enqueue - host_blocked,
            'alpha_step_cuda_graph': '_alpha_graph' in net.__dict__, 'extra_untimed_warmup': EXTRA_WARMUP, 'clocks': sampler.summary(),
            'roofline': roofline}
    line['mixedop_roofline'] = mop
    if world == 1 and not args.no_cpu_baseline:
        del model, net, opt_w, opt_a, pool
        torch.cuda.empty_cache()
        line['gpu_baseline'] = gpu_baseline_sample(dev)
        line['cpu_baseline'] = cpu_baseline_sample()
    emit(json.dumps(line))


class StdoutGuard(object):
    """Route everything libraries write to fd 1 (e.g. NCCL's version banner) to stderr; the JSON line is the
    only thing that reaches the real stdout."""

    def __init__(self):
        sys.stdout.flush()
        self.real = os.dup(1)
        os.dup2(2, 1)

    def emit(self, text):
        sys.stdout.flush()
        os.write(self.real, (text + '\n').encode())


def main():
    guard = StdoutGuard()
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=8)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--no-cpu-baseline', dest='no_cpu_baseline', action='store_true')
    ap.add_argument('--profile-only', dest='profile_only', action='store_true',
                    help='warm-up + K units and exit (for ncu launch lists)')
    args = ap.parse_args()
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    if args.impl == 'reference':
        run_reference(args, rank, world, guard.emit)
        return
    from tfnas_b200.parallel import init_from_env
    rank, local, world = init_from_env()
    if world == 1 and args.gpus > 1:
        # launched without torchrun: re-exec under torch.distributed.run
        cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(args.gpus),
               '--master-addr', '127.0.0.1', '--master-port', '29531', os.path.abspath(__file__)] + sys.argv[1:]
        os.dup2(guard.real, 1)
        raise SystemExit(subprocess.call(cmd))
    run_b200(args, rank, local, world, guard.emit)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == '__main__':
    main()

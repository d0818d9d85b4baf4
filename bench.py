#!/usr/bin/env python
"""Benchmark of the IPS selection hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload traffic|mnist|mnist5000|camelyon]
                    [--precision bf16|fp32] [--impl ours|reference] [--skip a,b,...]

One "step" = one ``IPSNet.ips`` call over one batch of synthetic patches (``conf.B`` images x N patches);
metric = patches scanned per second.  Rank 0 prints ONE JSON line.

GPU arm (``--impl ours``).  ``value``: inputs resident in HBM, EXACTLY K timed steps after W warm-up steps, CUDA events,
barrier + synchronize on both sides, max over ranks; for N>1 (torch.distributed.run) every rank scans its own batch (the
images of a batch are independent: weak scaling, no data-path collective).  ``e2e``: the same call with HOST input and
the winners copied back.  Extra records in the same line:
  roofline              dominant kernel family, timed live under the SAME (native) executor through the library's
                        per-launch event hooks; peak = burst or sustained figure of MEASURED_PEAKS.json by region length
  sustained             the same step repeated for >= 2 s (power-state honest), N=1
  exact                 the index-exact precision (fp32) on the same workload, N=1
  workloads             driver-timed sub-records for the other BASELINE configs (mnist, mnist5000, camelyon), N=1
  gpu_library_baseline  the UNMODIFIED reference (baseline/_ref) on the same GPU through torch's cuDNN / cuBLAS, N=1
  cpu_baseline          the unmodified reference on the host cores (bounded sample), N=1
  train                 train images/s over the reference's track_efficiency bracket (data-parallel, gradient all-reduce)
  seq_sharded           north_star's multi-GPU mode: ONE CAMELYON bag (50k / 200k patches) or a batch of 16 slides with the
                        patch axis sharded over the ranks (strong scaling; NVLink peer-memory exchange, 'merge' and 'exact'
                        schedules, eager and as one CUDA graph), with an in-run parity self-check; at N=1 the same inputs
                        through ``IPSNet.ips`` (the 1-GPU point of the curve)

Reference arm (``--impl reference``): the reference's own CPU implementation of the path -- the unmodified
``IPSNet.ips`` from baseline/_ref (``kind: "reference"``; the oracle port when that copy is absent) -- on all host
cores, W warm-up + EXACTLY K timed steps, each step a bounded sample of the workload (stated in ``cpu_baseline.sample``).
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (preset, overrides, B, N)        -- BASELINE.json configs[1], [0], [2], [3]
    'traffic': ('traffic', {}, 16, 192),
    'mnist': ('mnist', {}, 16, 900),
    'mnist5000': ('mnist', {'N': 10000}, 2, 10000),
    'camelyon': ('camelyon', {}, 1, 50000),
}
# algorithmic work per scanned patch, SURVEY.md section 8(d)
ALG = {
    'traffic': dict(flop=883.06e6, bytes=120000),
    'mnist': dict(flop=105.17e6, bytes=10000),
    'mnist5000': dict(flop=105.17e6, bytes=10000),
    'camelyon': dict(flop=2.6225e6, bytes=8192),
}
# one step of the reference arm: (images, patches per image) -- about one second of CPU work on 16 cores
REF_SAMPLE = {'traffic': (4, 192), 'mnist': (2, 900), 'mnist5000': (1, 2000), 'camelyon': (1, 50000)}
Q_GAIN = 12.0          # conditioned logits; does not change the work


def conf_for(workload, precision):
    from ips_b200.configs import load_config
    pre, over, B, N = WORKLOADS[workload]
    conf = load_config(pre, precision=precision, pretrained=False, **over)     # random init: no network for ImageNet weights
    conf.B = B
    return conf, B, N


def patch_shape(conf, B, N):
    return (B, N, conf.n_chan_in, *conf.patch_size) if conf.is_image else (B, N, conf.n_chan_in)


def workload_config(args, world):
    """`config` of the JSON line: identical for both arms (the driver compares them)."""
    conf, B, N = conf_for(args.workload, args.precision)
    in_mib = 4 * B * N * (conf.n_chan_in * conf.patch_size[0] * conf.patch_size[1] if conf.is_image else conf.n_chan_in) / 2**20
    return {'workload': f'{args.workload}: IPSNet.ips, B={B} N={N} M={conf.M} I={conf.I} per GPU',
            'precision': args.precision,
            'parallelism': f'dp{world} (independent batches, no collective)',
            'l2': f'input {in_mib:.0f} MiB per step > 126 MB L2, no flush needed'}


def load_peaks():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            return json.load(f)
    except Exception:
        return {}


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples SM clock / throttle reasons of one GPU during the timed region (NVML)."""

    def __init__(self, index):
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {'hw_slowdown': 'nvmlClocksThrottleReasonHwSlowdown',
                 'hw_thermal_slowdown': 'nvmlClocksThrottleReasonHwThermalSlowdown',
                 'sw_thermal_slowdown': 'nvmlClocksThrottleReasonSwThermalSlowdown',
                 'sw_power_cap': 'nvmlClocksThrottleReasonSwPowerCap'}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, attr in names.items():
                    if mask & getattr(nv, attr, 0):
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(0.004)            # the timed region of a default run is ~60 ms

    def __enter__(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._loop, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thread:
            self._thread.join()

    def summary(self):
        if not self.samples:
            return None
        return {'sm_mhz': statistics.median(self.samples), 'sm_max_mhz': self.max_mhz, 'reasons': sorted(self.reasons),
                'samples': len(self.samples)}


# ----------------------------------------------------------------------------- the reference (CPU arm, GPU library bar)
def reference_step(workload, device, sample):
    """Returns (callable running ONE ips() call of the reference on `device`, patches per call, kind, description).
    kind 'reference' = the unmodified IPSNet from baseline/_ref; 'port' = oracle/ips_oracle.py when that copy is absent."""
    import torch
    pre, over, _, _ = WORKLOADS[workload]
    Bs, Ns = sample
    over = dict(over)
    if workload == 'mnist5000':
        over['N'] = Ns                                            # pos-enc table of the sample's length
    sys.path.insert(0, os.path.join(ROOT, 'baseline'))
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import ips_oracle as O                                        # deterministic synthetic inputs (both kinds)
    try:
        import ref_harness
        conf = ref_harness.reference_conf(pre, **over)
        net = ref_harness.build_reference_net(conf, device, seed=1234)
        with torch.no_grad():
            net.transf.crs_attn.q.mul_(Q_GAIN)
        x = O.make_patches(conf, Bs, Ns, 1).to(device)

        def step():
            return net.ips(x)
        return step, Bs * Ns, 'reference', net
    except Exception as e:                                        # baseline/_ref missing: the oracle port (CPU only)
        if torch.device(device).type != 'cpu':
            raise
        sys.stderr.write('reference copy unavailable (%s): timing the oracle port\n' % (e,))
        conf = O.preset(pre, **over)
        sd = O.make_state(conf, 0, q_gain=Q_GAIN)
        x = O.make_patches(conf, Bs, Ns, 1)

        def step():
            return O.ips(sd, conf, x, perm='draw')
        return step, Bs * Ns, 'port', None


def cpu_reference(workload, warmup, steps):
    """The reference on the host cores: `warmup` untimed + `steps` timed calls on the bounded sample."""
    import torch
    cores = os.cpu_count()
    torch.set_num_threads(cores)
    Bs, Ns = REF_SAMPLE[workload]
    step, patches, kind, _ = reference_step(workload, torch.device('cpu'), (Bs, Ns))
    for _ in range(warmup):
        step()
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
    mean = sum(times) / len(times)
    sample = (f'{Bs} image(s) x {Ns} patches per step, {warmup} warm-up + {steps} timed calls (mean); fp32, '
              f'{"unmodified reference IPSNet.ips (baseline/_ref)" if kind == "reference" else "oracle port"} on ATen CPU ops, {cores} threads')
    return patches / mean, 1e3 * mean, cores, kind, sample


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    t_start = time.perf_counter()
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rate, ms, cores, kind, sample = cpu_reference(args.workload, args.warmup, args.steps)
    line = {
        'impl': 'reference', 'metric': 'ips_selection_patches_per_sec', 'value': rate, 'unit': 'patches/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic',
        'config': workload_config(args, world),
        'cpu_baseline': {'value': rate, 'unit': 'patches/s', 'cores': cores, 'kind': kind, 'sample': sample},
        'e2e': {'value': rate, 'unit': 'patches/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0, 'wall_s': time.perf_counter() - t_start,
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------- GPU arm
class Bench:
    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.args = torch, dist, args
        self.world = int(os.environ.get('WORLD_SIZE', '1'))
        self.rank = int(os.environ.get('RANK', '0'))
        self.local = int(os.environ.get('LOCAL_RANK', '0'))
        torch.cuda.set_device(self.local)
        self.dev = torch.device('cuda', self.local)
        if self.world > 1:
            dist.init_process_group('nccl', device_id=self.dev)
        self.peaks = load_peaks()
        self.skip = set(filter(None, args.skip.split(',')))

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def timed(self, fn, steps, collective=True):
        """ms for `steps` calls: CUDA events, barrier + synchronize on both sides, max over ranks."""
        torch = self.torch
        if collective:
            self.barrier()
        else:
            torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        if collective:
            self.barrier()
        else:
            torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=self.dev)
        if collective and self.world > 1:
            self.dist.all_reduce(ms, op=self.dist.ReduceOp.MAX)
        return ms.item()

    def make_net(self, workload, precision, seed=None):
        from ips_b200 import IPSNet, Struct
        torch = self.torch
        conf, B, N = conf_for(workload, precision)
        torch.manual_seed(1234 + self.rank if seed is None else seed)
        net = IPSNet(self.dev, Struct(**conf.__dict__)).to(self.dev)
        net.train()
        with torch.no_grad():
            net.transf.crs_attn.q.mul_(Q_GAIN)
        return net, conf, B, N

    # ---- one workload: device-resident value (+ optional e2e) ------------------------------------------------
    def measure(self, workload, precision, steps, warmup, e2e=True, collective=True, features='fp32'):
        torch = self.torch
        from ips_b200 import ops
        net, conf, B, N = self.make_net(workload, precision)
        shape = patch_shape(conf, B, N)
        x = torch.randn(shape, device=self.dev)             # inputs resident in HBM (eager mode)
        step = lambda: net.ips(x)
        for _ in range(warmup):
            step()
        launches0 = ops.LAUNCHES
        with ClockSampler(self.local) as clk:
            ms = self.timed(step, steps, collective)
        launches = ops.LAUNCHES - launches0
        mult = self.world if collective else 1
        out = {'value': mult * B * N * steps / (ms / 1e3), 'ms_per_step': ms / steps, 'steps': steps, 'warmup': warmup,
               'clocks': clk.summary(), 'gpu_launches': launches, 'net': net, 'conf': conf, 'B': B, 'N': N, 'x': x,
               'in_bytes': x.numel() * 4}
        if e2e:
            # public API with a HOST tensor (the reference's lazy mode, conf.eager=False): ips() streams it to the device
            # chunk by chunk on a copy stream while the encoder works on the chunks that have arrived
            xh = (x.to(torch.bfloat16) if features == 'bf16' and not conf.is_image else x).cpu().pin_memory()
            res_h = torch.empty((B, conf.M, *shape[2:]), dtype=torch.float32).pin_memory()
            idx_h = torch.empty((B, conf.M), dtype=torch.int64).pin_memory()

            def e2e_step():
                mem_patch, _ = net.ips(xh)
                res_h.copy_(mem_patch, non_blocking=True)    # D2H of the step's result
                idx_h.copy_(net.last_mem_idx, non_blocking=True)

            e2e_steps = max(2, min(steps, 5))
            e2e_step()
            ms_e = self.timed(e2e_step, e2e_steps, collective)
            out['e2e'] = {'value': mult * B * N * e2e_steps / (ms_e / 1e3), 'unit': 'patches/s',
                          'h2d_bytes_per_step': xh.numel() * xh.element_size(),
                          'd2h_bytes_per_step': res_h.numel() * 4 + idx_h.numel() * 8, 'ms_per_step': ms_e / e2e_steps,
                          'steps': e2e_steps}
            del xh, res_h, idx_h
        return out

    # ---- roofline of the dominant kernel family, under the executor that was timed ---------------------------
    def roofline(self, m, workload, region_ms, reps=3):
        torch = self.torch
        from ips_b200 import ops
        net, conf, B, N, x = m['net'], m['conf'], m['B'], m['N'], m['x']
        alg = ALG[workload]
        burst = region_ms < 1000.0                           # a region this short runs at boost clocks: burst peak
        if conf.is_image and net.executor == 'native' and net.precision == 'bf16':
            net.ips(x)
            torch.cuda.synchronize()
            ops.profile_begin()
            for _ in range(reps):
                net.ips(x)
            torch.cuda.synchronize()
            recs = ops.profile_end()
            fam = [(a, b) for k, a, b in recs if k in ('stem', 'conv')]
            busy = ops.busy_ms(fam) / reps
            per_kind = {}
            for k, a, b in recs:
                per_kind[k] = per_kind.get(k, 0.0) + (b - a) / reps
            peak_key = 'bf16_tflops' if burst else 'bf16_tflops_sustained'
            peak = self.peaks.get(peak_key, 1663.8 if burst else 1403.1)
            achieved = alg['flop'] * B * N / (busy / 1e3) / 1e12
            roof = {'bound': 'tensor', 'kernel': 'conv family of the native executor (fused stem+pool + every BasicBlock convolution; tcgen05)',
                    'achieved': achieved, 'peak': peak, 'unit': 'TFLOP/s', 'frac': achieved / peak, 'traffic': None,
                    'peak_source': ('MEASURED_PEAKS.json ' + peak_key if self.peaks else 'fallback (B200_PROFILING.md)') +
                                   (' (timed region %.0f ms < 1 s: burst conditions)' % region_ms if burst else ' (timed region >= 1 s)'),
                    'launches_per_step': len(fam) // reps, 'family_busy_ms_per_step': busy,
                    'family_sum_ms_per_step': sum(b - a for a, b in fam) / reps,
                    'how': 'CUDA events around every launch on the lane it runs on (ipsb_profile_begin/end), same native executor and '
                           'lanes as the timed region; busy = union of the family\'s intervals over the concurrent lanes; achieved = '
                           'algorithmic FLOP per step / busy',
                    'algorithmic_flop_per_patch': alg['flop'],
                    'kernel_ms_all': {k: round(v, 4) for k, v in sorted(per_kind.items(), key=lambda kv: -kv[1])}}
            for name in ('r02_traffic.json', 'r01_traffic.json'):
                path = os.path.join(ROOT, 'profiles', name)
                if workload == 'traffic' and os.path.exists(path):
                    with open(path) as f:
                        tj = json.load(f)['conv_family']
                    roof['traffic'] = tj['dram_bytes'] / tj['launches']
                    roof['traffic_note'] = ('STORED CONSTANT, not measured in this run: dram read+write bytes per launch averaged over the %d '
                                            'conv-family launches of one step, from the committed ncu launch list (profiles/%s)'
                                            % (tj['launches'], name))
                    break
            return roof
        # per-call events (one library call per kernel): feature-bag path and the fp32 mode
        ops.TIMER = {}
        prev = net.executor
        net.executor = 'python'
        for _ in range(reps):
            net.ips(x)
        torch.cuda.synchronize()
        net.executor = prev
        per = {k: sum(a.elapsed_time(b) for a, b, _ in v) / reps for k, v in ops.TIMER.items()}
        counts = {k: len(v) // reps for k, v in ops.TIMER.items()}
        ops.TIMER = None
        family = ('ipsb_conv_bf16_pf', 'ipsb_conv_bf16_umma', 'ipsb_stem_pool_s2d', 'ipsb_linear_bf16_umma', 'ipsb_conv_f32',
                  'ipsb_linear_f32', 'ipsb_projector_logits', 'ipsb_projector_logits_scan')
        fam_ms = sum(v for k, v in per.items() if k in family)
        if fam_ms <= 0:
            return None
        tensor = net.precision == 'bf16'
        peak_key = ('bf16_tflops' if burst else 'bf16_tflops_sustained') if tensor else None
        peak = self.peaks.get(peak_key, 1663.8 if burst else 1403.1) if tensor else 75.0
        achieved = alg['flop'] * B * N / (fam_ms / 1e3) / 1e12
        return {'bound': 'tensor' if tensor else 'fp32-fma', 'kernel': '+'.join(k for k in family if k in per),
                'achieved': achieved, 'peak': peak, 'unit': 'TFLOP/s', 'frac': achieved / peak, 'traffic': None,
                'peak_source': ('MEASURED_PEAKS.json ' + peak_key) if tensor else 'nominal fp32 FMA rate (148 SMs x 128 lanes x 2 x ~1.9 GHz)',
                'launches_per_step': sum(v for k, v in counts.items() if k in family), 'kernel_ms_per_step': fam_ms,
                'algorithmic_flop_per_patch': alg['flop'],
                'kernel_ms_all': {k: round(v, 4) for k, v in sorted(per.items(), key=lambda kv: -kv[1])}}

    # ---- the unmodified reference on this GPU through torch's libraries --------------------------------------
    def gpu_library_baseline(self, workload, calls=5):
        torch = self.torch
        pre, over, B, N = WORKLOADS[workload]
        out = {}
        for label in ('fp32_torch_defaults', 'bf16_autocast_channels_last'):
            step, patches, kind, net = reference_step(workload, self.dev, (B, N))
            if label.startswith('bf16'):
                net.to(memory_format=torch.channels_last)
                inner = step

                def step(inner=inner):
                    with torch.autocast('cuda', dtype=torch.bfloat16):
                        return inner()
            for _ in range(2):
                step()
            ms = self.timed(step, calls, collective=False)
            out[label] = {'value': patches * calls / (ms / 1e3), 'unit': 'patches/s', 'ms_per_step': ms / calls, 'calls': calls}
            del net, step
            torch.cuda.empty_cache()
        out['what'] = ('UNMODIFIED reference IPSNet.ips (baseline/_ref) on this GPU, torch %s cuDNN/cuBLAS/ATen kernels, same workload '
                       'and synthetic input; fp32 = torch defaults (cudnn.allow_tf32=True), bf16 = torch.autocast + channels_last weights'
                       % torch.__version__)
        return out

    # ---- train images/s over the reference's track_efficiency bracket -----------------------------------------
    def train(self, m, workload, steps, sync_bn):
        torch = self.torch
        net, conf, B, N, x = m['net'], m['conf'], m['B'], m['N'], m['x']
        net.sync_bn = bool(sync_bn)
        net.sync_bn_equal_shares = True                       # every rank trains on B_train images: no row counts to exchange
        labels = {}
        for task in conf.tasks.values():
            if task['metric'] == 'multilabel_accuracy':
                labels[task['name']] = (torch.rand(B, conf.n_class, device=self.dev) < 0.3).float()
            elif task['act_fn'] == 'sigmoid':
                labels[task['name']] = torch.randint(0, 2, (B,), device=self.dev).float()
            else:
                labels[task['name']] = torch.randint(0, conf.n_class, (B,), device=self.dev)
        # the reference's bracket (training/iterative.py:124-171): loader batches of B_seq images go through ips(), their
        # winners fill a train batch of B_train images, then ONE forward/backward/optimizer step
        B_train = 16 if workload == 'camelyon' else max(B, int(getattr(conf, 'B_seq', B)))
        n_calls = B_train // B
        labels_t = {k: (v.repeat(n_calls, *([1] * (v.dim() - 1)))) for k, v in labels.items()}
        from ips_b200.train import GraphedTrainStep
        # the reference's optimizer (main.py:57: AdamW with conf.wd); `fused` = torch's multi-tensor kernel: the capturable
        # foreach form divides by 0-dim step tensors one parameter at a time (150 tiny launches per step)
        opt = torch.optim.AdamW(net.parameters(), lr=1e-4, weight_decay=conf.wd, capturable=True, fused=True)
        gstep = GraphedTrainStep(net, conf, opt, B_train, data_parallel=True if self.world > 1 else None)
        for k, v in labels_t.items():
            gstep.labels[k].copy_(v)
        for i in range(n_calls):
            net.ips(x, out=gstep.buffers, row_offset=i * B)
        try:
            gstep.capture()
        except Exception as e:                                # e.g. a collective that cannot be captured here: eager step
            sys.stderr.write('train step: CUDA-graph capture failed (%s); running eagerly\n' % (str(e)[:200],))
            net.sync_bn_equal_shares = False
            gstep = GraphedTrainStep(net, conf, opt, B_train, data_parallel=True if self.world > 1 else None)
            for k, v in labels_t.items():
                gstep.labels[k].copy_(v)
            gstep.capture()
        graphed = gstep.graph is not False and gstep.graph is not None

        def train_step():
            for i in range(n_calls):                          # ips() re-folds the updated parameters every step
                net.ips(x, out=gstep.buffers, row_offset=i * B)
            gstep()

        train_step(); train_step()
        ms_t = self.timed(train_step, steps)
        net.sync_bn = False
        return {'metric': 'train_images_per_sec', 'value': self.world * B_train * steps / (ms_t / 1e3), 'ms_per_step': ms_t / steps,
                'images_per_step': B_train, 'ips_calls_per_step': n_calls, 'graphed': graphed, 'sync_bn': bool(sync_bn), 'steps': steps,
                'note': 'ips() (plan re-folded from the updated weights every step) + forward + loss + backward + AdamW; '
                        + ('grad-mode half replayed as one CUDA graph' if graphed else 'grad-mode half eager (synchronised BatchNorm '
                           'puts collectives inside forward/backward)')
                        + ('; data parallel: ONE NCCL all-reduce of the flat gradient buffer between the [forward+backward] graph and the '
                           '[1/R scaling + AdamW] graph' if self.world > 1 and graphed else '')}

    # ---- north_star's multi-GPU mode: one sequence sharded over the ranks ----------------------------------------
    def seq_sharded(self, steps):
        torch, dist = self.torch, self.dist
        from ips_b200 import ops
        from ips_b200.distributed import ShardedIPS, shard_bounds
        from ips_b200.utils import scan_order
        R, rank = self.world, self.rank
        net, conf, _, _ = self.make_net('camelyon', 'bf16', seed=4321)           # identical weights on every rank
        ca = net.transf.crs_attn
        HT, M = ca.H * ca.n_token, conf.M
        F = conf.n_chan_in
        cases = [('bag_50k', 1, 50000), ('bag_200k', 1, 200000), ('slides_16x50k', 16, 50000)]
        records = []
        for name, B, N in cases:
            g = torch.Generator(device=self.dev).manual_seed(99)
            x = torch.randn((B, N, F), generator=g, device=self.dev)             # the same sequence on every rank
            rec = {'case': name, 'B': B, 'N': N, 'M': M, 'I': conf.I, 'n_gpus': R, 'scaling': 'strong'}
            if R == 1:
                step = lambda: net.ips(x)
                for rng in ('reference', 'device'):
                    net.scan_order_rng = rng
                    for _ in range(3):
                        step()
                    ms = self.timed(step, steps) / steps
                    r1 = dict(rec)
                    r1.update(mode='single GPU IPSNet.ips', scan_order_rng=rng, ms=ms, patches_per_s=B * N / (ms / 1e3))
                    records.append(r1)
                net.scan_order_rng = 'reference'
                del x
                continue
            lo, hi = shard_bounds(N, R)[rank]
            local = x[:, lo:hi].contiguous()
            variants = [('merge', 'replicated'), ('exact', 'replicated')]
            if B >= R:
                variants.append(('merge', 'batch_split'))      # slide b delivered only to the rank whose train step consumes it
            for rng in ('reference', 'device'):          # scan order: the reference's host RNG calls / drawn on the device
                net.scan_order_rng = rng
                for mode, output in variants:
                    if rng == 'device' and mode == 'exact':
                        continue
                    sh = ShardedIPS(net, B, N, (F,), mode=mode, output=output)
                    spr = sh.spr
                    my_b = list(range(rank * spr, min(B, (rank + 1) * spr))) if spr else list(range(B))
                    for _ in range(3):
                        sh(local)
                    ms_eager = self.timed(lambda: sh(local), steps) / steps
                    # parity self-check (eager call, known seeds)
                    torch.manual_seed(100 + rank if mode == 'merge' else 7)
                    mem_patch, _ = sh(local)
                    got_idx = net.last_mem_idx.clone()
                    got_sum = mem_patch.double().sum().item()
                    rows_ok = all(bool(torch.equal(mem_patch[j], x[b, got_idx[j]])) for j, b in enumerate(my_b))
                    if mode == 'exact':
                        torch.manual_seed(7)
                        ref_patch, _ = net.ips(x)                                     # the same call on ONE GPU
                        same = bool(torch.equal(net.last_mem_idx[my_b], got_idx)) and bool(torch.equal(ref_patch[my_b], mem_patch))
                        parity = {'exact_equals_single_gpu_bit_for_bit': same}
                    else:
                        # the same schedule composed in ONE process from the product kernels: per-slice loop, candidates in rank
                        # order, one global re-score
                        z = net.patch_logits(x)
                        cz, ci = [], []
                        for r, (a, b_) in enumerate(shard_bounds(N, R)):
                            torch.manual_seed(100 + r)
                            perm, per_inst = scan_order(net.shuffle, net.shuffle_style, B, b_ - a, torch.device('cpu'), rng, self.dev)
                            perm = None if perm is None else perm.to(self.dev)
                            zl = z[:, a:b_].contiguous()
                            cand = ops.select_loop(zl, perm, per_inst, ca.H, ca.n_token, M, conf.I)[1]
                            cz.append(torch.gather(zl, 1, cand.unsqueeze(-1).expand(-1, -1, HT)))
                            ci.append(cand + a)
                        cz, ci = torch.cat(cz, 1).contiguous(), torch.cat(ci, 1)
                        pos = ops.merge_candidates(cz, ca.H, ca.n_token, M)
                        ref_idx = torch.gather(ci, 1, pos)
                        parity = {'merge_equals_single_process_schedule': bool(torch.equal(ref_idx[my_b], got_idx))}
                    parity['rows_are_the_selected_patches'] = rows_ok
                    flag = torch.tensor([int(all(parity.values()))], device=self.dev)
                    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
                    parity['all_ranks'] = bool(flag.item())
                    parity['exchange_status'] = sh.ex.status()
                    all_idx = sh.mem_src                                              # (B, M) on every rank
                    owned = ((all_idx >= lo) & (all_idx < hi))
                    if spr:                                                         # rows this rank sends to OTHER ranks
                        dest = torch.arange(B, device=self.dev).unsqueeze(1) // spr
                        sent_rows = int((owned & (dest != rank)).sum().item())
                    else:
                        sent_rows = int(owned.sum().item()) * (R - 1)
                    pushed = (R - 1) * (B * M * (HT * 4 + 8) if mode == 'merge' else B * (hi - lo) * HT * 4) + sent_rows * F * 4
                    # the whole call as ONE CUDA graph
                    ms_graph, graph_err = None, None
                    try:
                        sh.capture(local)
                        gin = sh.static_input                                        # the slice lives in the graph's input buffer
                        for _ in range(3):
                            sh(gin)
                        ms_graph = self.timed(lambda: sh(gin), steps) / steps
                        torch.manual_seed(100 + rank if mode == 'merge' else 7)
                        sh(gin)
                        parity['graph_replay_equals_eager'] = bool(torch.equal(net.last_mem_idx, got_idx))
                    except Exception as e:                                           # keep the eager record
                        graph_err = str(e)[:200]
                    best = min(v for v in (ms_eager, ms_graph) if v is not None)
                    r2 = dict(rec)
                    r2.update(mode=mode, output=output, scan_order_rng=rng, transport='nvlink peer memory (CUDA IPC exchange buffers, push + flag kernels; no NCCL on the data path)',
                              ms_eager=ms_eager, ms_graph=ms_graph, ms=best, patches_per_s=B * N / (best / 1e3),
                              bytes_pushed_to_peers_per_rank=pushed, parity_check=parity)
                    if graph_err:
                        r2['graph_error'] = graph_err
                    records.append(r2)
                    del sh
            net.scan_order_rng = 'reference'
            # the collective baseline transport (one packed all-gather + all-reduce of the winners), eager
            from ips_b200.distributed import ips_sharded
            if name != 'slides_16x50k':
                for _ in range(2):
                    ips_sharded(net, local, N, mode='merge', transport='nccl')
                ms_nccl = self.timed(lambda: ips_sharded(net, local, N, mode='merge', transport='nccl'), steps) / steps
                r3 = dict(rec)
                r3.update(mode='merge', transport='nccl baseline (packed all-gather of candidates, all-reduce of winners)', ms=ms_nccl,
                          patches_per_s=B * N / (ms_nccl / 1e3),
                          nccl_bytes_per_rank=B * M * (HT + 2) * 4 * (R - 1) + 2 * B * M * F * 4 * (R - 1) // R)
                records.append(r3)
            del x, local
            net.__dict__.pop('_sharded', None)
            torch.cuda.empty_cache()
        out = {'workload': 'camelyon features (2048-d), M=I=5000, bf16 tensor cores; patch axis sharded over the ranks (strong scaling)',
               'records': records}
        if R > 1 and 'oracle' not in self.skip:
            out['merge_fp32_vs_oracle'] = self.sharded_vs_oracle()
        return out

    def sharded_vs_oracle(self, N=20000):
        """fp32-mode 'merge' schedule on R GPUs against oracle.ips_sharded (P5) on a conditioned bag; set comparison + P3."""
        torch, dist = self.torch, self.dist
        sys.path.insert(0, os.path.join(ROOT, 'oracle'))
        import ips_oracle as O
        from ips_b200 import IPSNet, Struct
        from ips_b200.distributed import ips_sharded, shard_bounds
        R, rank = self.world, self.rank
        conf = O.preset('camelyon', precision='fp32', M=1000, I=1000, attn_dropout=0.0, dropout=0.0)
        sd = O.make_state(conf, 3, q_gain=Q_GAIN)
        x = O.make_patches(conf, 1, N, 4)
        net = IPSNet(self.dev, Struct(**conf.__dict__)).to(self.dev)
        net.load_state_dict(sd)
        lo, hi = shard_bounds(N, R)[rank]
        torch.manual_seed(100 + rank)
        ips_sharded(net, x[:, lo:hi].contiguous().to(self.dev), N, mode='merge')
        got = net.last_mem_idx.cpu()
        res = None
        if rank == 0:
            blocks = []
            for r, (a, b) in enumerate(shard_bounds(N, R)):
                torch.manual_seed(100 + r)
                blocks.append(O.draw_permutation(conf, 1, b - a) + a)
            _, _, o_src = O.ips_sharded(sd, conf, x, R=R, perm=torch.cat(blocks, dim=1), tie='stable')
            a, g = set(o_src[0].tolist()), set(got[0].tolist())
            res = {'N': N, 'M': conf.M, 'identical_order': bool(torch.equal(got, o_src)), 'same_set': a == g, 'differing_picks': len(a ^ g) // 2}
        self.barrier()
        return res


def run_ours(args):
    bench = Bench(args)
    torch, dist = bench.torch, bench.dist
    world, rank = bench.world, bench.rank
    warmup = max(args.warmup, 3)
    skip = bench.skip
    extras_ok = world == 1

    def guarded(name, fn):
        if name in skip:
            return None
        try:
            return fn()
        except Exception as e:                               # an auxiliary record must not cost the headline line
            import traceback
            sys.stderr.write('bench: %s failed on rank %d: %s\n' % (name, rank, traceback.format_exc()))
            return {'error': str(e)[:300]}

    m = bench.measure(args.workload, args.precision, args.steps, warmup, e2e=True, features=args.features)
    region_ms = m['ms_per_step'] * args.steps
    roof = guarded('roofline', lambda: bench.roofline(m, args.workload, region_ms)) if rank == 0 else None

    sustained = None
    if extras_ok and 'sustained' not in skip:
        net, x = m['net'], m['x']
        n_sus = max(args.steps, int(2200.0 / m['ms_per_step']))
        with ClockSampler(bench.local) as clk_s:
            ms_s = bench.timed(lambda: net.ips(x), n_sus)
        v = m['B'] * m['N'] * n_sus / (ms_s / 1e3)
        sustained = {'seconds': ms_s / 1e3, 'steps': n_sus, 'value': v, 'ms_per_step': ms_s / n_sus, 'clocks': clk_s.summary()}
        if roof and roof.get('bound') == 'tensor' and 'family_busy_ms_per_step' in roof:
            share = roof['family_busy_ms_per_step'] / m['ms_per_step']
            ach = ALG[args.workload]['flop'] * m['B'] * m['N'] / (share * ms_s / n_sus / 1e3) / 1e12
            pk = bench.peaks.get('bf16_tflops_sustained', 1403.1)
            sustained['roofline'] = {'achieved': ach, 'peak': pk, 'frac': ach / pk, 'unit': 'TFLOP/s',
                                     'how': 'family share of the step (from the roofline pass) x sustained step time; peak = bf16_tflops_sustained'}

    # the staging kernel alone (TMA-staged unfold): patches -> space-to-depth frame for half a step's patches,
    # input + output larger than L2, CUDA events around 20 launches on the current stream
    staging = None
    if extras_ok and rank == 0 and m['conf'].is_image and 'staging' not in skip:
        def run_staging():
            from ips_b200 import ops
            xs = m['x'].reshape(-1, *m['x'].shape[2:])
            rows = xs.shape[0] // 2
            C, H, W = xs.shape[1:]
            if H % 2 or W % 2 or C > 4:
                return None
            for _ in range(3):
                ops.stage_patches_s2d(xs, rows, C, H, W)
            reps = 20
            ms_st = bench.timed(lambda: ops.stage_patches_s2d(xs, rows, C, H, W, first_row=rows), reps, collective=False) / reps
            by = rows * (C * H * W * 4 + (H // 2 + 3) * (W // 2 + 3) * 32)
            pk = bench.peaks.get('hbm_gbs', 6545.6)
            return {'kernel': 'stage_s2d_tma_kernel' if (W % 4 == 0 and not os.environ.get('IPSB_STAGE_NO_TMA')) else 'stage_s2d_kernel', 'rows': rows, 'ms': ms_st,
                    'bytes': by, 'achieved': by / (ms_st / 1e3) / 1e9, 'peak': pk, 'unit': 'GB/s', 'frac': by / (ms_st / 1e3) / 1e9 / pk,
                    'how': 'algorithmic bytes (fp32 patch read once + bf16 frame written once) / mean launch time, kernel alone'}
        staging = guarded('staging', run_staging)

    train = guarded('train', lambda: bench.train(m, args.workload, max(2, min(args.steps, 5)), sync_bn=world > 1))
    train_local = None
    if world > 1 and 'train' not in skip:
        train_local = guarded('train', lambda: bench.train(m, args.workload, max(2, min(args.steps, 5)), sync_bn=False))

    exact = None
    if extras_ok and args.precision == 'bf16' and 'exact' not in skip:
        def run_exact():
            me = bench.measure(args.workload, 'fp32', 3, 1, e2e=False)
            return {'precision': 'fp32 (CUDA-core SIMT kernels, one fp32 accumulation chain per output: selected indices equal the '
                                 'reference\'s, tests/test_gpu_ips.py)', 'value': me['value'], 'ms_per_step': me['ms_per_step'], 'steps': 3}
        exact = guarded('exact', run_exact)

        def run_x3():
            me = bench.measure(args.workload, 'bf16x3', 5, 2, e2e=False)
            return {'precision': "bf16x3 (near-fp32 on the tensor cores: operands as hi + lo bf16 pairs, three tcgen05 products per "
                                 "multiply, fp32 accumulation and fp32 between layers; same winners as the fp32 reference on the parity "
                                 "fixtures and at the BASELINE sizes, tests/test_gpu_ips.py)",
                    'value': me['value'], 'ms_per_step': me['ms_per_step'], 'steps': 5}
        exact_tc = guarded('exact', run_x3)
        if isinstance(exact, dict) and exact_tc is not None:
            exact['tensor_core_near_fp32'] = exact_tc

    workloads = None
    if extras_ok and 'workloads' not in skip:
        workloads = {}
        for w in WORKLOADS:
            if w == args.workload:
                continue

            def run_w(w=w):
                k = 20 if w != 'mnist5000' else 10
                mw = bench.measure(w, args.precision, k, 3, e2e=True)
                r = bench.roofline(mw, w, mw['ms_per_step'] * k)
                rec = {'value': mw['value'], 'unit': 'patches/s', 'ms_per_step': mw['ms_per_step'], 'steps': k, 'warmup': 3,
                       'e2e': mw['e2e'], 'gpu_launches': mw['gpu_launches'], 'clocks': mw['clocks'],
                       'config': f"{w}: IPSNet.ips, B={mw['B']} N={mw['N']} M={mw['conf'].M} I={mw['conf'].I}", 'roofline': r}
                if mw['N'] >= 5000:
                    # long sequences: the reference's host-side randperm (0.4 ms for 50 000 patches) is longer than the
                    # GPU work of the call, so the default configuration above is HOST bound; the same call with the scan
                    # order drawn on the device (conf.scan_order_rng='device': a shuffle, but not the reference's random stream)
                    net_w, x_w = mw['net'], mw['x']
                    sub = {}
                    for label, shuffle, rng in (('device_scan_order', True, 'device'), ('no_shuffle', False, 'reference')):
                        net_w.shuffle, net_w.scan_order_rng = shuffle, rng
                        for _ in range(3):
                            net_w.ips(x_w)
                        ms_d = bench.timed(lambda: net_w.ips(x_w), k) / k
                        sub[label] = {'value': bench.world * mw['B'] * mw['N'] / (ms_d / 1e3), 'ms_per_step': ms_d}
                    net_w.shuffle, net_w.scan_order_rng = True, 'reference'
                    sub['note'] = ('value / ms_per_step above: default configuration = scan order drawn with the reference\'s host RNG calls '
                                   '(torch.randperm on the CPU, then copied), which bounds the call; device_scan_order: conf.scan_order_rng='
                                   '"device" (keyed bijection drawn on the GPU); no_shuffle: conf.shuffle=False')
                    rec['scan_order_variants'] = sub
                del mw
                torch.cuda.empty_cache()
                if 'library' not in skip:
                    try:
                        rec['gpu_library_baseline'] = bench.gpu_library_baseline(w, calls=3)
                    except Exception as e:
                        rec['gpu_library_baseline'] = {'error': str(e)[:200]}
                return rec
            workloads[w] = guarded('workloads', run_w)
            torch.cuda.empty_cache()

    lib = None
    if extras_ok and 'library' not in skip:
        lib = guarded('library', lambda: bench.gpu_library_baseline(args.workload))

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu and 'cpu' not in skip:
        def run_cpu():
            rate, ms, cores, kind, sample = cpu_reference(args.workload, 2, 10)
            return {'value': rate, 'unit': 'patches/s', 'cores': cores, 'kind': kind, 'sample': sample}
        cpu = guarded('cpu', run_cpu)

    del m['x']
    m.pop('net')
    torch.cuda.empty_cache()
    seq = guarded('seq', lambda: bench.seq_sharded(max(3, min(args.steps, 10)))) if 'seq' not in skip else None

    if rank == 0:
        line = {
            'metric': 'ips_selection_patches_per_sec', 'value': m['value'], 'unit': 'patches/s',
            'n_gpus': world, 'steps': args.steps, 'warmup': warmup, 'ms_per_step': m['ms_per_step'],
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'bf16' if args.precision == 'bf16' else 'f32', 'data': 'synthetic',
            'config': workload_config(args, world),
            'clocks': m['clocks'], 'e2e': m['e2e'], 'gpu_launches': m['gpu_launches'],
            'roofline': roof, 'staging': staging, 'cpu_baseline': cpu, 'sustained': sustained, 'exact': exact, 'workloads': workloads,
            'gpu_library_baseline': lib, 'train': train, 'seq_sharded': seq,
        }
        if train_local is not None:
            line['train_local_bn_stats'] = train_local
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='traffic', choices=sorted(WORKLOADS))
    ap.add_argument('--precision', default='bf16', choices=['bf16', 'fp32', 'bf16x3'])
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline leg')
    ap.add_argument('--skip', default='', help='comma list of auxiliary records to skip: roofline,sustained,train,exact,workloads,'
                                               'library,cpu,seq,oracle')
    ap.add_argument('--features', default='fp32', choices=['fp32', 'bf16'],
                    help='feature-bag workloads: dtype of the HOST features of the e2e leg (bf16 = flat bf16 bags, SURVEY 8f N4)')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()

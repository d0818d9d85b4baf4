#!/usr/bin/env python
"""Benchmark of the IPS selection hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload traffic|mnist|mnist5000|camelyon]
                    [--precision bf16|fp32] [--impl ours|reference]

One "step" = one ``IPSNet.ips`` call over one batch of synthetic patches
(``conf.B`` images x N patches); metric = patches scanned per second.  For N>1 GPUs
(launched with torch.distributed.run) every rank scans its own batch (weak scaling, no
data-path collective); the time is the max over ranks.  Rank 0 prints ONE JSON line.

``--impl reference`` times the reference algorithm on the host CPU (the oracle port:
the same ATen CPU operators the reference's modules dispatch to) on a bounded sample of
the same workload.
"""
import argparse
import json
import os
import statistics
import subprocess
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


def conf_for(workload, precision):
    from ips_b200.configs import load_config
    pre, over, B, N = WORKLOADS[workload]
    conf = load_config(pre, precision=precision, **over)
    conf.B = B
    return conf, B, N


def patch_shape(conf, B, N):
    return (B, N, conf.n_chan_in, *conf.patch_size) if conf.is_image else (B, N, conf.n_chan_in)


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
            self._stop.wait(0.004)            # the timed region of a default run is ~30 ms

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


# ----------------------------------------------------------------------------- CPU legs (oracle)
def cpu_reference_rate(workload, sample_images, repeats):
    """Reference algorithm on the host cores; returns (patches/s, cores, sample description)."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import ips_oracle as O
    cores = os.cpu_count()
    torch.set_num_threads(cores)
    pre, over, B, N = WORKLOADS[workload]
    conf = O.preset(pre, **over)
    if workload in ('mnist5000', 'camelyon'):
        sample_images = 1
        N = min(N, 20000 if workload == 'camelyon' else 2000)
        if conf.use_pos:
            conf.N = N
    Bs = min(B, sample_images)
    sd = O.make_state(conf, 0, q_gain=12.0)
    x = O.make_patches(conf, Bs, N, 1)
    times = []
    O.ips(sd, conf, x[:, :max(conf.M + conf.I, N // 4)], perm='draw')            # warm-up on a slice
    for _ in range(repeats):
        t0 = time.perf_counter()
        O.ips(sd, conf, x, perm='draw')
        times.append(time.perf_counter() - t0)
    t = statistics.median(times)
    return Bs * N / t, cores, f'{Bs} image(s) x {N} patches, {repeats} calls, median; fp32 ATen CPU ops, {cores} threads'


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    t_start = time.perf_counter()
    rate, cores, sample = cpu_reference_rate(args.workload, sample_images=16, repeats=max(1, min(args.steps, 5)))
    conf, B, N = conf_for(args.workload, 'fp32')
    line = {
        'impl': 'reference', 'metric': 'ips_selection_patches_per_sec', 'value': rate, 'unit': 'patches/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': 1e3 * B * N / rate, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': f'{args.workload}: IPSNet.ips, B={B} N={N} M={conf.M} I={conf.I} per GPU', 'l2': 'n/a (CPU)'},
        'cpu_baseline': {'value': rate, 'unit': 'patches/s', 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': rate, 'unit': 'patches/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0, 'wall_s': time.perf_counter() - t_start,
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------- GPU arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    from ips_b200 import IPSNet, Struct, ops

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)

    conf, B, N = conf_for(args.workload, args.precision)
    torch.manual_seed(1234 + rank)
    net = IPSNet(dev, Struct(**conf.__dict__)).to(dev)
    net.train()
    with torch.no_grad():
        net.transf.crs_attn.q.mul_(12.0)           # conditioned logits; does not change the work
    shape = patch_shape(conf, B, N)
    x = torch.randn(shape, device=dev)             # inputs resident in HBM (eager mode)
    in_bytes = x.numel() * 4

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    # ---- device-resident throughput ------------------------------------------------
    seq = args.shard == 'sequence' and world > 1
    if seq:                                        # one batch, patch axis sharded: same data on every rank's slice
        from ips_b200.distributed import ips_sharded, shard_bounds
        torch.manual_seed(1234)
        lo_s, hi_s = shard_bounds(N, world)[rank]
        x_local = x[:, lo_s:hi_s].contiguous()
        step = lambda: ips_sharded(net, x_local, N, mode=args.shard_mode)
    else:
        step = lambda: net.ips(x)
    for _ in range(max(args.warmup, 3)):
        step()
    launches0 = ops.LAUNCHES
    with ClockSampler(local) as clk:
        ms = timed(step, args.steps)
    launches = ops.LAUNCHES - launches0
    value = (1 if seq else world) * B * N * args.steps / (ms / 1e3)

    # ---- end to end through the public API with HOST buffers ---------------------
    xh = (x.to(torch.bfloat16) if args.features == 'bf16' and not conf.is_image else x).cpu().pin_memory()
    res_h = torch.empty((B, conf.M, *shape[2:]), dtype=torch.float32).pin_memory()
    idx_h = torch.empty((B, conf.M), dtype=torch.int64).pin_memory()

    def e2e_step():
        # public API with a HOST tensor (the reference's lazy mode, conf.eager=False): ips() streams it to the
        # device chunk by chunk on a copy stream while the encoder works on the chunks that have arrived
        mem_patch, _ = net.ips(xh)
        res_h.copy_(mem_patch, non_blocking=True)  # D2H of the step's result
        idx_h.copy_(net.last_mem_idx, non_blocking=True)

    e2e_steps = max(2, min(args.steps, 5))
    e2e_step()
    ms_e2e = timed(e2e_step, e2e_steps)
    e2e_value = world * B * N * e2e_steps / (ms_e2e / 1e3)

    # ---- secondary metric: train images/s over the reference's track_efficiency bracket
    #      (ips + forward + loss + backward + AdamW step; the grad-mode half runs on PyTorch autograd in round 1)
    train = None
    if not args.no_train:
        import torch.nn.functional as Fn
        opt = torch.optim.AdamW(net.parameters(), lr=1e-4, weight_decay=conf.wd)
        labels = {}
        for task in conf.tasks.values():
            if task['metric'] == 'multilabel_accuracy':
                labels[task['name']] = (torch.rand(B, conf.n_class, device=dev) < 0.3).float()
            elif task['act_fn'] == 'sigmoid':
                labels[task['name']] = torch.randint(0, 2, (B,), device=dev).float()
            else:
                labels[task['name']] = torch.randint(0, conf.n_class, (B,), device=dev)

        # the reference's bracket (training/iterative.py:124-171): loader batches of B_seq images go through ips(),
        # their winners fill a train batch of B_train images, then ONE forward/backward/optimizer step
        B_train = max(B, int(getattr(conf, 'B_seq', B)))
        if args.workload == 'camelyon':
            B_train = 16                                  # config/camelyon_config.yml: B = 16, B_seq = 1
        n_calls = B_train // B
        labels_t = {k: (v.repeat(n_calls, *([1] * (v.dim() - 1)))) for k, v in labels.items()}

        # forward + loss + backward + AdamW of the B_train x M winners as ONE CUDA graph over static buffers; every
        # ips() call writes its winners straight into them (ips_b200/train.py)
        from ips_b200.train import GraphedTrainStep
        opt = torch.optim.AdamW(net.parameters(), lr=1e-4, weight_decay=conf.wd, capturable=True)
        hook = None
        if world > 1:
            from ips_b200.distributed import allreduce_gradients
            hook = allreduce_gradients
        gstep = GraphedTrainStep(net, conf, opt, B_train, grad_hook=hook)
        for k, v in labels_t.items():
            gstep.labels[k].copy_(v)
        for i in range(n_calls):
            net.ips(x, out=gstep.buffers, row_offset=i * B)
        graphed = True
        try:
            gstep.capture()
        except Exception as e:                      # e.g. a collective that cannot be captured: eager step
            graphed = False
            sys.stderr.write('train step: CUDA-graph capture failed (%s); running eagerly\n' % (e,))

        def train_step():
            for i in range(n_calls):
                net.ips(x, out=gstep.buffers, row_offset=i * B)
            if graphed:
                gstep()
            else:
                gstep._step()

        tsteps = max(2, min(args.steps, 5))
        train_step(); train_step()
        ms_t = timed(train_step, tsteps)
        train = {'metric': 'train_images_per_sec', 'value': world * B_train * tsteps / (ms_t / 1e3), 'ms_per_step': ms_t / tsteps,
                 'images_per_step': B_train, 'ips_calls_per_step': n_calls,
                 'graphed': graphed,
                 'note': 'ips(), every nn.Linear, LayerNorm, BatchNorm1d and the attention core run forward AND backward on the '
                         'library kernels (conv encoder: tcgen05 forward, dgrad and wgrad), replayed as one CUDA graph; pooling / residual / ReLU glue and AdamW are PyTorch ops inside the graph'}

    # ---- roofline of the dominant kernel family (per-launch CUDA events, same work) ---
    roof = None
    if rank == 0:
        ops.TIMER = {}
        net.executor = 'python'                    # one library call per layer so each launch can be bracketed
        net.ips(x)
        net.ips(x)
        torch.cuda.synchronize()
        net.executor = 'native'
        per = {k: sum(a.elapsed_time(b) for a, b, _ in v) / 2 for k, v in ops.TIMER.items()}   # ms per step
        counts = {k: len(v) // 2 for k, v in ops.TIMER.items()}
        ops.TIMER = None
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        except Exception:
            pass
        # the tensor-core convolution / GEMM entry points form one kernel family
        family = ('ipsb_conv_bf16_pf', 'ipsb_conv_bf16_umma', 'ipsb_stem_pool_s2d', 'ipsb_linear_bf16_umma', 'ipsb_conv_f32',
                  'ipsb_linear_f32')     # (the fused stem kernel carries the stem's FLOPs and the max-pool)
        fam_ms = sum(v for k, v in per.items() if k in family)
        fam_n = sum(v for k, v in counts.items() if k in family)
        dom = max(per, key=per.get)
        if fam_ms >= per[dom]:
            dom = 'conv/gemm family (' + '+'.join(k for k in family if k in per) + ')'
            per[dom], counts[dom] = fam_ms, fam_n
        alg = ALG[args.workload]
        if dom.startswith('conv/gemm'):
            peak = peaks.get('bf16_tflops_sustained', 1400.0)
            enc_flop = alg['flop'] * B * N
            achieved = enc_flop / (per[dom] / 1e3) / 1e12
            roof = {'bound': 'tensor', 'kernel': dom, 'achieved': achieved, 'peak': peak, 'unit': 'TFLOP/s',
                    'frac': achieved / peak, 'traffic': None,
                    'peak_source': 'MEASURED_PEAKS.json bf16_tflops_sustained' if peaks else 'fallback 1.4 PFLOP/s',
                    'launches_per_step': counts[dom], 'kernel_ms_per_step': per[dom],
                    'algorithmic_flop_per_patch': alg['flop']}
        else:
            peak = peaks.get('hbm_gbs', 6650.0)
            achieved = alg['bytes'] * B * N / (per[dom] / 1e3) / 1e9
            roof = {'bound': 'hbm', 'kernel': dom, 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
                    'frac': achieved / peak, 'traffic': None,
                    'peak_source': 'MEASURED_PEAKS.json hbm_gbs' if peaks else 'fallback 6650 GB/s',
                    'launches_per_step': counts[dom], 'kernel_ms_per_step': per[dom]}
        roof['kernel_ms_all'] = {k: round(v, 4) for k, v in sorted(per.items(), key=lambda kv: -kv[1])}
        # DRAM bytes per launch of the same kernel family, from the committed ncu launch list of this command
        # (tools/launch_traffic.py -> profiles/r01_traffic.json; cold-cache, serialised replays)
        try:
            if args.workload == 'traffic' and args.precision == 'bf16' and dom.startswith('conv/gemm'):
                with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'profiles', 'r01_traffic.json')) as f:
                    tj = json.load(f)['conv_family']
                roof['traffic'] = tj['dram_bytes'] / tj['launches']
                roof['traffic_note'] = ('dram read+write bytes per launch averaged over the %d conv-family launches of one step '
                                        '(ncu, profiles/r01_traffic.json); algorithmic = FLOP-bound family' % tj['launches'])
        except Exception:
            pass

    # ---- CPU baseline on rank 0 at N=1 -------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        rate, cores, sample = cpu_reference_rate(args.workload, sample_images=16, repeats=3)
        cpu = {'value': rate, 'unit': 'patches/s', 'cores': cores, 'kind': 'port', 'sample': sample}

    if rank == 0:
        line = {
            'metric': 'ips_selection_patches_per_sec', 'value': value, 'unit': 'patches/s',
            'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3), 'ms_per_step': ms / args.steps,
            'higher_is_better': True, 'scaling': 'strong' if seq else 'weak', 'vs_baseline': None,
            'dtype': 'bf16' if args.precision == 'bf16' else 'f32', 'data': 'synthetic',
            'config': {'workload': f'{args.workload}: IPSNet.ips, B={B} N={N} M={conf.M} I={conf.I} per GPU',
                       'precision': args.precision, 'parallelism': ((f'sp{world} (patch axis sharded; all-gather of logits, replicated loop, all-reduce of winners)' if args.shard_mode == 'exact' else f'sp{world} (patch axis sharded; local top-M per rank, all-gather of M candidates, global re-score, all-reduce of winners)') if seq
                                       else f'dp{world} (independent batches, no collective)'),
                       'l2': f'input {in_bytes / 2**20:.0f} MiB per step > 126 MB L2, no flush needed'},
            'clocks': clk.summary(),
            'e2e': {'value': e2e_value, 'unit': 'patches/s', 'h2d_bytes_per_step': xh.numel() * xh.element_size(),
                    'd2h_bytes_per_step': res_h.numel() * 4 + idx_h.numel() * 8, 'ms_per_step': ms_e2e / e2e_steps},
            'gpu_launches': launches,
            'roofline': roof, 'cpu_baseline': cpu, 'train': train,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='traffic', choices=sorted(WORKLOADS))
    ap.add_argument('--precision', default='bf16', choices=['bf16', 'fp32'])
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline leg')
    ap.add_argument('--no-train', action='store_true', help='skip the secondary train images/s measurement')
    ap.add_argument('--features', default='fp32', choices=['fp32', 'bf16'],
                    help='feature-bag workloads: dtype of the HOST features of the e2e leg (bf16 = flat bf16 bags, SURVEY 8f N4)')
    ap.add_argument('--shard-mode', default='merge', choices=['exact', 'merge'],
                    help="--shard sequence: 'merge' = local top-M per rank + candidate merge (north_star; the loop shards), "
                         "'exact' = logit table all-gathered, loop replicated (bit-identical to one GPU)")
    ap.add_argument('--shard', default='batch', choices=['batch', 'sequence'],
                    help="N>1: 'batch' = every rank scans its own batch (weak scaling); 'sequence' = ONE batch whose patch axis "
                         "is sharded over the ranks (strong scaling, NCCL all-gather of logits + all-reduce of winners)")
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()

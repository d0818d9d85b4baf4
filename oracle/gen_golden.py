"""Generate tests/golden/*.npz from the UNMODIFIED reference.  TEST INFRASTRUCTURE.

Runs only in the build container, where /root/reference exists:

    python oracle/gen_golden.py            # rewrites tests/golden/

For every case in ``CASES`` it
  1. builds the reference ``IPSNet`` (architecture/ips_net.py:85) from the
     case's config and loads ``ips_oracle.make_state`` weights (strict),
  2. runs ``net.ips`` (ips_net.py:169) exactly as training/iterative.py:135
     does, in train() mode, with the global CPU RNG seeded so the 'batch'
     shuffle permutation is reproducible,
  3. runs ``net.transf.get_scores`` / ``net.encoder`` / ``net.score_and_select``
     on seeded tensors (unit-level pins),
  4. runs the reference's own ``compute_loss`` (training/iterative.py:65) in
     train() mode with dropout 0 and back-propagates,
and stores the outputs.  Inputs and weights are NOT stored: they are
regenerated from seeds by ``ips_oracle.make_state/make_patches`` (CPU
generator, machine independent for a fixed torch build).
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ips_oracle as O  # noqa: E402

REF = '/root/reference'

# name -> (preset, overrides, B, N, q_gain, sparse)
CASES = {
    'mnist_small': ('mnist', dict(N=36, M=8, I=8, attn_dropout=0.0, dropout=0.0), 2, 36, 12.0, 0.0),
    'mnist_ties': ('mnist', dict(N=36, M=8, I=8, attn_dropout=0.0, dropout=0.0), 2, 36, 1.0, 0.9),
    'mnist_noshuffle': ('mnist', dict(N=30, M=6, I=10, shuffle=False, attn_dropout=0.0, dropout=0.0), 3, 30, 12.0, 0.0),
    'mnist_instance': ('mnist', dict(N=25, M=5, I=7, shuffle_style='instance', attn_dropout=0.0, dropout=0.0), 2, 25, 12.0, 0.0),
    'traffic_small': ('traffic', dict(N=12, M=4, I=5, attn_dropout=0.0, dropout=0.0), 2, 12, 12.0, 0.0),
    'camelyon_small': ('camelyon', dict(M=128, I=128, attn_dropout=0.0, dropout=0.0), 1, 700, 12.0, 0.0),
    'camelyon_batch': ('camelyon', dict(M=32, I=48, attn_dropout=0.0, dropout=0.0), 3, 333, 12.0, 0.0),
    'camelyon_shortcut': ('camelyon', dict(M=64, I=64, attn_dropout=0.0, dropout=0.0), 2, 40, 12.0, 0.0),
}
STATE_SEED, DATA_SEED, RNG_SEED, LABEL_SEED = 11, 23, 5, 7


def ref_modules():
    sys.path.insert(0, REF)
    from architecture.ips_net import IPSNet            # noqa
    from utils.utils import Struct                      # noqa
    from training import iterative                      # noqa
    return IPSNet, Struct, iterative


def recover_src(patches, mem_patch):
    """Match every selected patch back to its position in the original axis."""
    B, N = patches.shape[:2]
    flat = patches.reshape(B, N, -1)
    out = torch.zeros(mem_patch.shape[:2], dtype=torch.int64)
    for b in range(B):
        sel = mem_patch[b].reshape(mem_patch.shape[1], -1)
        eq = (sel.unsqueeze(1) == flat[b].unsqueeze(0)).all(-1)       # (M,N)
        assert (eq.sum(1) >= 1).all()
        out[b] = eq.float().argmax(1)
    return out


def run_case(name, IPSNet, Struct, iterative):
    pre, over, B, N, gain, sparse = CASES[name]
    conf = O.preset(pre, **over)
    conf.B = B
    sd = O.make_state(conf, STATE_SEED, q_gain=gain)
    patches = O.make_patches(conf, B, N, DATA_SEED, sparse=sparse)

    net = IPSNet(torch.device('cpu'), Struct(**conf.__dict__))
    net.load_state_dict(sd, strict=True)
    net.train()

    out = {}
    # ---- unit pins -------------------------------------------------------
    g = torch.Generator().manual_seed(101)
    L = min(conf.M + conf.I, 64)
    emb = torch.randn(B, L, conf.D, generator=g)
    with torch.no_grad():
        net.eval()
        out['unit_emb_scores'] = net.transf.get_scores(emb).numpy()
        x_small = patches[:, :3].reshape(-1, *patches.shape[2:])
        out['unit_encode'] = net.encoder(x_small).reshape(x_small.shape[0], -1).numpy()
        idx = torch.arange(L).unsqueeze(0).expand(B, -1)
        m_small = max(1, L // 3)
        se, si = net.score_and_select(emb, None, m_small, idx)
        out['unit_select_idx'] = si.numpy()
        net.train()

    # ---- ips end to end --------------------------------------------------
    torch.manual_seed(RNG_SEED)
    mem_patch, mem_pos = net.ips(patches)
    assert net.training and net.encoder.training and net.transf.training
    out['mem_patch_sum'] = np.array(mem_patch.double().sum().item())
    out['mem_patch_shape'] = np.array(mem_patch.shape)
    if conf.M < N:
        out['mem_src'] = recover_src(patches, mem_patch).numpy() if sparse == 0 else np.zeros(0)
    if mem_pos is not None:
        out['mem_pos'] = mem_pos.numpy()
    # full tensors only when small
    if mem_patch.numel() <= 200_000:
        out['mem_patch'] = mem_patch.numpy()

    # diagnostics: boundary gap of the final iteration via the oracle trace
    trace = []
    torch.manual_seed(RNG_SEED)
    o_patch, o_pos, o_src = O.ips(sd, conf, patches, perm='draw', tie='topk', trace=trace)
    same = torch.equal(o_patch, mem_patch)
    gap = None
    if trace:
        s = trace[-1][0].sort(-1, descending=True)[0]
        gap = float(((s[:, conf.M - 1] - s[:, conf.M]) / s[:, conf.M - 1]).min())
        lg = O.attn_logits(sd, conf, emb)
        print(f'  {name}: oracle==reference {same}; last-iter rel boundary gap {gap:.3e}; '
              f'unit logit std {float(lg.std()):.3f}')
    else:
        print(f'  {name}: oracle==reference {same} (shortcut)')

    # ---- train step (reference compute_loss, dropout 0, train-mode BN) ---
    labels = O.make_labels(conf, B, LABEL_SEED)
    crit = {}
    for task in conf.tasks.values():
        crit[task['name']] = torch.nn.NLLLoss() if task['act_fn'] == 'softmax' else torch.nn.BCELoss()
    loss, info = iterative.compute_loss(net, mem_patch, mem_pos, crit, labels, Struct(**conf.__dict__))
    loss.backward()
    out['loss'] = np.array(loss.item())
    for t, p in info[1].items():
        out['pred_' + t] = np.asarray(p)
    gsd = {k: v.grad for k, v in net.named_parameters()}
    for key in ['transf.crs_attn.q', 'transf.crs_attn.k_w.weight', 'transf.mlp.w_2.bias',
                'encoder.0.weight', 'encoder.1.weight']:
        if key in gsd and gsd[key] is not None:
            out['grad_' + key] = gsd[key].reshape(-1)[:256].numpy().copy()
            out['gradnorm_' + key] = np.array(gsd[key].double().norm().item())
    if not conf.is_image:
        out['bn_running_mean_after'] = net.encoder[2].running_mean.numpy().copy()
    else:
        out['bn_running_mean_after'] = net.encoder[1].running_mean.numpy().copy()

    meta = dict(preset=pre, over=over, B=B, N=N, q_gain=gain, sparse=sparse,
                state_seed=STATE_SEED, data_seed=DATA_SEED, rng_seed=RNG_SEED, label_seed=LABEL_SEED,
                boundary_gap=gap, torch=torch.__version__)
    out['meta'] = np.array(json.dumps(meta))
    return out


def main():
    if not os.path.isdir(REF):
        raise SystemExit('needs /root/reference (build container only)')
    torch.set_num_threads(os.cpu_count())
    IPSNet, Struct, iterative = ref_modules()
    dst = os.path.join(os.path.dirname(HERE), 'tests', 'golden')
    os.makedirs(dst, exist_ok=True)
    for name in CASES:
        out = run_case(name, IPSNet, Struct, iterative)
        np.savez_compressed(os.path.join(dst, name + '.npz'), **out)
    print('wrote', dst)


if __name__ == '__main__':
    main()

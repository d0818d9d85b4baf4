"""CPU oracle for the IPS hot path.  TEST INFRASTRUCTURE ONLY.

This file is a functional restatement (torch CPU ops on a plain ``dict`` of
tensors, no ``nn.Module``) of the reference's iterative patch selection path.
It is the *checker* for the CUDA product in ``ips_b200/``; only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl
reference`` legs may import it.  Nothing under ``ips_b200/`` imports it.

Pinning: the reference ships no tests or golden vectors, so
the pin is the live reference itself: ``oracle/gen_golden.py`` imports the
unmodified reference from ``/root/reference`` in the build container, loads
the weights produced by :func:`make_state` through ``load_state_dict`` and
stores its outputs under ``tests/golden/``; ``tests/test_oracle_golden.py``
asserts this restatement reproduces them bit-for-bit.

The arithmetic of the reference lives in a third-party dependency that is not
vendored in its tree: PyTorch (pinned ``pytorch=1.13.0``,
``requirements.txt:156``; this image has 2.11.0) and torchvision's ResNet
definition (``torchvision=0.14.0``, ``requirements.txt:181``).  The oracle
calls the same ATen CPU operators (``conv2d``, ``batch_norm``, ``linear``,
``softmax``, ``topk``) that the reference's modules dispatch to.

Reference map (all paths relative to /root/reference):
  pos_table            architecture/transformer.py:6-18
  attn_scores          architecture/transformer.py:29-34,71-83,143-148
  aggregate            architecture/transformer.py:85-109,122-132,150-152
  encode (image)       architecture/ips_net.py:17-52 (torchvision resnet18 trunk)
  encode (features)    architecture/ips_net.py:54-60
  score_and_select     architecture/ips_net.py:136-155
  ips                  architecture/ips_net.py:169-262, utils/utils.py:33-58
  forward / heads      architecture/ips_net.py:157-166,264-283
  loss                 training/iterative.py:65-100, main.py:53-61
"""
import math

import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------
# configuration bags
# --------------------------------------------------------------------------


class Conf:
    """Attribute bag with the reference's YAML keys (utils/utils.py:10-12)."""

    def __init__(self, **kw):
        self.__dict__.update(kw)

    def replace(self, **kw):
        d = dict(self.__dict__)
        d.update(kw)
        return Conf(**d)


_MNIST_TASKS = {
    'task0': {'id': 0, 'name': 'majority', 'act_fn': 'softmax', 'metric': 'accuracy'},
    'task1': {'id': 1, 'name': 'max', 'act_fn': 'softmax', 'metric': 'accuracy'},
    'task2': {'id': 2, 'name': 'top', 'act_fn': 'softmax', 'metric': 'accuracy'},
    'task3': {'id': 3, 'name': 'multi', 'act_fn': 'sigmoid', 'metric': 'multilabel_accuracy'},
}


def preset(name, **over):
    """The three shipped configs (config/*.yml), as attribute bags.

    ``pretrained`` is forced to False for traffic (no network; north_star asks
    for random init).  Any key can be overridden, e.g. to shrink a fixture.
    """
    common = dict(eps=1e-6, seed=0, eager=True, shuffle=True, shuffle_style='batch',
                  attn_dropout=0.1, dropout=0.1, H=8, pretrained=False,
                  track_efficiency=False, track_epoch=0, wd=0.1)
    if name == 'mnist':
        c = dict(B=16, B_seq=16, n_class=10, is_image=True, enc_type='resnet18', n_chan_in=1,
                 n_res_blocks=2, n_token=4, N=900, M=100, I=100, patch_size=[50, 50],
                 patch_stride=[50, 50], use_pos=True, D=128, D_k=16, D_v=16, D_inner=512,
                 lr=1e-3, tasks=_MNIST_TASKS)
    elif name == 'traffic':
        c = dict(B=16, B_seq=16, n_class=4, is_image=True, enc_type='resnet18', n_chan_in=3,
                 n_res_blocks=4, n_token=1, N=192, M=10, I=32, patch_size=[100, 100],
                 patch_stride=[100, 100], use_pos=False, D=512, D_k=64, D_v=64, D_inner=2048,
                 lr=3e-4,
                 tasks={'task0': {'id': 0, 'name': 'sign', 'act_fn': 'softmax', 'metric': 'accuracy'}})
    elif name == 'camelyon':
        # camelyon_config.yml has no N / n_res_blocks / patch_size keys
        c = dict(B=16, B_seq=1, n_class=1, is_image=False, enc_type='resnet50', n_chan_in=2048,
                 n_token=1, M=5000, I=5000, use_pos=False, D=512, D_k=64, D_v=64, D_inner=2048,
                 lr=3e-4,
                 tasks={'task0': {'id': 0, 'name': 'metastases', 'act_fn': 'sigmoid', 'metric': 'auc'}})
    else:
        raise KeyError(name)
    common.update(c)
    common.update(over)
    return Conf(**common)


# --------------------------------------------------------------------------
# parameter inventory + deterministic fixture weights
# --------------------------------------------------------------------------

def _resnet18_trunk_shapes(n_chan_in, n_res_blocks):
    """(key, shape) for the truncated resnet18 (ips_net.py:34-50 keeps children
    0,1,4,5[,6,7] of the Sequential)."""
    out = [('encoder.0.weight', (64, n_chan_in, 7, 7))]
    out += _bn_shapes('encoder.1', 64)
    widths = [64, 128, 256, 512][:n_res_blocks]
    cin = 64
    for li, w in enumerate(widths):
        for blk in range(2):
            p = f'encoder.{4 + li}.{blk}.'
            stride = 2 if (li > 0 and blk == 0) else 1
            out.append((p + 'conv1.weight', (w, cin, 3, 3)))
            out += _bn_shapes(p + 'bn1', w)
            out.append((p + 'conv2.weight', (w, w, 3, 3)))
            out += _bn_shapes(p + 'bn2', w)
            if stride != 1 or cin != w:
                out.append((p + 'downsample.0.weight', (w, cin, 1, 1)))
                out += _bn_shapes(p + 'downsample.1', w)
            cin = w
    return out


def _bn_shapes(prefix, c):
    return [(prefix + '.weight', (c,)), (prefix + '.bias', (c,)),
            (prefix + '.running_mean', (c,)), (prefix + '.running_var', (c,)),
            (prefix + '.num_batches_tracked', ())]


def param_shapes(conf):
    """Ordered (key, shape) list of the reference IPSNet state_dict."""
    D = conf.D
    if conf.is_image:
        out = _resnet18_trunk_shapes(conf.n_chan_in, conf.n_res_blocks)
    else:
        out = [('encoder.1.weight', (D, conf.n_chan_in)), ('encoder.1.bias', (D,))]
        out += _bn_shapes('encoder.2', D)
    H, Dk, Dv = conf.H, conf.D_k, conf.D_v
    out += [('transf.crs_attn.q', (1, conf.n_token, D)),
            ('transf.crs_attn.q_w.weight', (H * Dk, D)),
            ('transf.crs_attn.k_w.weight', (H * Dk, D)),
            ('transf.crs_attn.v_w.weight', (H * Dv, D)),
            ('transf.crs_attn.fc.weight', (D, H * Dv)),
            ('transf.crs_attn.layer_norm.weight', (D,)),
            ('transf.crs_attn.layer_norm.bias', (D,)),
            ('transf.mlp.w_1.weight', (conf.D_inner, D)), ('transf.mlp.w_1.bias', (conf.D_inner,)),
            ('transf.mlp.w_2.weight', (D, conf.D_inner)), ('transf.mlp.w_2.bias', (D,)),
            ('transf.mlp.layer_norm.weight', (D,)), ('transf.mlp.layer_norm.bias', (D,))]
    for task in conf.tasks.values():
        out += [(f"output_layers.{task['name']}.0.weight", (conf.n_class, D)),
                (f"output_layers.{task['name']}.0.bias", (conf.n_class,))]
    return out


def make_state(conf, seed, q_gain=1.0):
    """Deterministic fixture weights from a CPU generator (machine independent
    for a fixed torch version).  He-scaled weights, non-trivial BN statistics so
    the eval-mode BN folding is exercised; ``q_gain`` conditions the logits
    (default init makes selection rounding noise)."""
    g = torch.Generator().manual_seed(int(seed))
    sd = {}
    for key, shape in param_shapes(conf):
        leaf = key.rsplit('.', 1)[-1]
        if leaf == 'num_batches_tracked':
            t = torch.zeros((), dtype=torch.int64)
        elif leaf == 'running_var':
            t = torch.rand(shape, generator=g) + 0.5
        elif leaf == 'running_mean':
            t = torch.randn(shape, generator=g) * 0.1
        elif key.endswith('crs_attn.q'):
            lim = math.sqrt(1.0 / conf.D_k)
            t = (torch.rand(shape, generator=g) * 2 - 1) * lim * q_gain
        elif len(shape) == 1 and leaf == 'weight':          # BN / LN scale
            t = torch.rand(shape, generator=g) + 0.5
        elif len(shape) == 1:                               # biases
            t = torch.randn(shape, generator=g) * 0.1
        else:                                               # conv / linear weights
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            t = torch.randn(shape, generator=g) * math.sqrt(2.0 / fan_in)
        sd[key] = t
    return sd


def make_patches(conf, B, N, seed, sparse=0.0):
    """Synthetic input of the layout the reference datasets emit
    (data/megapixel_mnist/mnist_dataset.py:47-53): (B,N,C,ph,pw) or (B,N,F)."""
    g = torch.Generator().manual_seed(int(seed))
    if conf.is_image:
        x = torch.randn(B, N, conf.n_chan_in, *conf.patch_size, generator=g)
    else:
        x = torch.randn(B, N, conf.n_chan_in, generator=g)
    if sparse > 0:
        keep = torch.rand(B, N, generator=g) >= sparse
        x = x * keep.view(B, N, *([1] * (x.dim() - 2))).to(x.dtype)
    return x


# --------------------------------------------------------------------------
# building blocks
# --------------------------------------------------------------------------

def pos_table(D, n):
    """Sin/cos table (n, D); transformer.py:6-18."""
    if D % 2:
        raise ValueError('odd D')
    tab = torch.zeros(n, D)
    pos = torch.arange(0, n).unsqueeze(1)
    freq = torch.exp(torch.arange(0, D, 2, dtype=torch.float) * -(math.log(10000.0) / D))
    tab[:, 0::2] = torch.sin(pos.float() * freq)
    tab[:, 1::2] = torch.cos(pos.float() * freq)
    return tab


def _bn2d(x, sd, p, train, momentum=0.1):
    return F.batch_norm(x, sd[p + '.running_mean'], sd[p + '.running_var'], sd[p + '.weight'],
                        sd[p + '.bias'], training=train, momentum=momentum, eps=1e-5)


def _block(x, sd, p, stride, train):
    y = F.relu(_bn2d(F.conv2d(x, sd[p + 'conv1.weight'], stride=stride, padding=1), sd, p + 'bn1', train))
    y = _bn2d(F.conv2d(y, sd[p + 'conv2.weight'], padding=1), sd, p + 'bn2', train)
    if (p + 'downsample.0.weight') in sd:
        x = _bn2d(F.conv2d(x, sd[p + 'downsample.0.weight'], stride=stride), sd, p + 'downsample.1', train)
    return F.relu(y + x)


def encode(sd, conf, x, train=False):
    """Patch encoder.  x: (P,C,ph,pw) or (P,F) -> (P,D).  ips_net.py:17-60.
    ``train`` selects batch statistics (and updates running stats in ``sd``
    in place, like nn.BatchNorm)."""
    if conf.is_image:
        y = F.conv2d(x, sd['encoder.0.weight'], stride=2, padding=3)
        y = F.relu(_bn2d(y, sd, 'encoder.1', train))
        y = F.max_pool2d(y, kernel_size=3, stride=2, padding=1)
        for li in range(conf.n_res_blocks):
            for blk in range(2):
                y = _block(y, sd, f'encoder.{4 + li}.{blk}.', 2 if (li > 0 and blk == 0) else 1, train)
        return F.adaptive_avg_pool2d(y, 1).flatten(1)
    y = F.layer_norm(x, (x.shape[-1],), eps=1e-5)
    y = F.linear(y, sd['encoder.1.weight'], sd['encoder.1.bias'])
    y = F.batch_norm(y, sd['encoder.2.running_mean'], sd['encoder.2.running_var'],
                     sd['encoder.2.weight'], sd['encoder.2.bias'], training=train, momentum=0.1, eps=1e-5)
    return F.relu(y)


def attn_logits(sd, conf, x):
    """(B,L,D) -> (B,H,T,L) pre-softmax logits; transformer.py:29-31,71-83."""
    H, Dk, T = conf.H, conf.D_k, conf.n_token
    B, L = x.shape[:2]
    q = F.linear(sd['transf.crs_attn.q'], sd['transf.crs_attn.q_w.weight']).view(1, T, H, Dk).transpose(1, 2)
    k = F.linear(x, sd['transf.crs_attn.k_w.weight']).view(B, L, H, Dk).transpose(1, 2)
    return torch.matmul(q / (Dk ** 0.5), k.transpose(2, 3))


def attn_scores(sd, conf, x):
    """(B,L,D) -> (B,L): softmax over L per (head, token), mean over heads then
    tokens; transformer.py:143-148 (dropout is identity in eval)."""
    a = torch.softmax(attn_logits(sd, conf, x), dim=-1)
    return a.mean(dim=1).transpose(1, 2).mean(-1)


def select_topm(scores, M, tie='topk'):
    """Positions of the M best scores per row, best first.
    tie='topk'   -> torch.topk exactly as ips_net.py:148 (tie order implementation defined)
    tie='stable' -> stable descending sort, lowest position wins ties: the
                    contract of the CUDA kernel."""
    if tie == 'topk':
        return torch.topk(scores, M, dim=-1)[1]
    return torch.sort(scores, dim=-1, descending=True, stable=True)[1][:, :M]


def score_and_select(sd, conf, emb, emb_pos, M, idx, tie='topk'):
    """ips_net.py:136-155."""
    s = attn_scores(sd, conf, emb_pos if torch.is_tensor(emb_pos) else emb)
    top = select_topm(s, M, tie)
    keep_emb = torch.gather(emb, 1, top.unsqueeze(-1).expand(-1, -1, emb.shape[2]))
    return keep_emb, torch.gather(idx, 1, top), s


def draw_permutation(conf, B, N):
    """RNG contract of the shuffle (utils/utils.py:33-58): 'batch' draws ONE
    randperm(N) from the global CPU generator, 'instance' draws rand(B,N) on the
    data's device and argsorts each row.  Returns (B,N) int64 or None."""
    if not conf.shuffle:
        return None
    if conf.shuffle_style == 'batch':
        return torch.randperm(N).unsqueeze(0).expand(B, -1)
    if conf.shuffle_style == 'instance':
        return torch.rand(B, N).argsort(1)
    return None  # the reference silently ignores unknown styles (ips_net.py:124-133)


_M64 = (1 << 64) - 1


def _splitmix64(x):
    x = (x + 0x9E3779B97F4A7C15) & _M64
    z = x
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _M64
    return z ^ (z >> 31)


def keyed_scan_order(key0, key1, rows, N):
    """CPU restatement of the library's DEVICE-side scan order (conf.scan_order_rng = 'device'; no counterpart in the
    reference, whose shuffle is the host randperm of utils/utils.py:33-58): row r of the (rows, N) result is a keyed
    bijection of [0, N): eight rounds of (xor c_r, multiply by odd m_r mod 2^k, xor-shift right by (k+1)//2 or k//3+1) on
    k = ceil(log2 N) bits, re-applied until the value is < N (cycle walking); (m_r, c_r) from splitmix64 seeded with
    key0 ^ splitmix64(key1 + r).  key0 / key1 are the two int64 the library draws from the CUDA generator."""
    import numpy as np
    out = np.zeros((rows, N), np.int64)
    if N == 1:
        return torch.from_numpy(out)
    k = max(1, int(N - 1).bit_length())
    mask = (1 << k) - 1
    h1, h2 = (k + 1) // 2, k // 3 + 1
    for row in range(rows):
        t = ((key0 & _M64) ^ _splitmix64(((key1 & _M64) + row) & _M64)) & _M64
        mul, add = [], []
        for _ in range(8):
            t = _splitmix64(t)
            mul.append(((t >> 32) | 1) & 0xffffffff)
            add.append(t & 0xffffffff)
        x = np.arange(N, dtype=np.uint64)
        todo = np.ones(N, bool)
        while todo.any():
            y = x[todo]
            for r in range(8):
                y ^= np.uint64(add[r] & mask)
                y = (y * np.uint64(mul[r])) & np.uint64(mask)
                y ^= y >> np.uint64(h1 if r % 2 == 0 else h2)
            x[todo] = y
            todo = x >= N
        out[row] = x.astype(np.int64)
    return torch.from_numpy(out)


def ips(sd, conf, patches, perm='draw', tie='topk', trace=None):
    """The no-grad selection loop, ips_net.py:169-262.

    Returns (mem_patch, mem_pos, mem_src) where mem_src (B,M) int64 are the
    winners' positions in the ORIGINAL (unshuffled) patch axis, best first.
    ``perm``: 'draw' consumes the RNG like the reference; a (B,N) tensor fixes
    the scan order; None disables shuffling.  ``trace`` (list) receives the
    (scores, mem_idx) of every iteration."""
    with torch.no_grad():
        M, I, D = conf.M, conf.I, conf.D
        B, N = patches.shape[:2]
        tail = patches.shape[2:]
        pos = pos_table(D, conf.N).unsqueeze(0).expand(B, -1, -1) if conf.use_pos else None
        if M >= N:                                               # :185-188
            src = torch.arange(N).unsqueeze(0).expand(B, -1)
            return patches, pos, src
        if isinstance(perm, str):
            perm = draw_permutation(conf, B, N)
        if perm is None:
            perm = torch.arange(N).unsqueeze(0).expand(B, -1)
        # shuffled views (the reference materialises these copies, utils.py:39,56)
        px = torch.gather(patches, 1, perm.reshape(B, N, *([1] * len(tail))).expand(-1, -1, *tail))
        ppos = torch.gather(pos, 1, perm.unsqueeze(-1).expand(-1, -1, D)) if conf.use_pos else None

        mem_emb = encode(sd, conf, px[:, :M].reshape(-1, *tail)).view(B, M, -1)
        idx = torch.arange(N, dtype=torch.int64).unsqueeze(0).expand(B, -1)
        mem_idx = idx[:, :M]
        for it in range(math.ceil((N - M) / I)):
            lo = it * I + M
            hi = min(lo + I, N)
            new_emb = encode(sd, conf, px[:, lo:hi].reshape(-1, *tail)).view(B, -1, D)
            all_emb = torch.cat((mem_emb, new_emb), dim=1)
            all_idx = torch.cat((mem_idx, idx[:, lo:hi]), dim=1)
            all_pos = None
            if conf.use_pos:
                all_pos = all_emb + torch.gather(ppos, 1, all_idx.view(B, -1, 1).expand(-1, -1, D))
            mem_emb, mem_idx, s = score_and_select(sd, conf, all_emb, all_pos, M, all_idx, tie)
            if trace is not None:
                trace.append((s.clone(), mem_idx.clone()))
        mem_patch = torch.gather(px, 1, mem_idx.view(B, -1, *([1] * len(tail))).expand(-1, -1, *tail))
        mem_pos = torch.gather(ppos, 1, mem_idx.unsqueeze(-1).expand(-1, -1, D)) if conf.use_pos else None
        return mem_patch, mem_pos, torch.gather(perm, 1, mem_idx)


def _layer_norm(x, sd, p, eps):
    return F.layer_norm(x, (x.shape[-1],), sd[p + '.weight'], sd[p + '.bias'], eps)


def aggregate(sd, conf, x):
    """Cross-attention pooling + MLP (dropout off): transformer.py:85-109,122-132."""
    H, Dk, Dv, T = conf.H, conf.D_k, conf.D_v, conf.n_token
    B, L = x.shape[:2]
    qtok = sd['transf.crs_attn.q']
    q = F.linear(qtok, sd['transf.crs_attn.q_w.weight']).view(1, T, H, Dk).transpose(1, 2)
    k = F.linear(x, sd['transf.crs_attn.k_w.weight']).view(B, L, H, Dk).transpose(1, 2)
    v = F.linear(x, sd['transf.crs_attn.v_w.weight']).view(B, L, H, Dv).transpose(1, 2)
    a = torch.softmax(torch.matmul(q / (Dk ** 0.5), k.transpose(2, 3)), dim=-1)
    o = torch.matmul(a, v).transpose(1, 2).contiguous().view(B, T, H * Dv)
    o = F.linear(o, sd['transf.crs_attn.fc.weight']) + qtok
    o = _layer_norm(o, sd, 'transf.crs_attn.layer_norm', 1e-6)
    h = F.linear(torch.relu(F.linear(o, sd['transf.mlp.w_1.weight'], sd['transf.mlp.w_1.bias'])),
                 sd['transf.mlp.w_2.weight'], sd['transf.mlp.w_2.bias'])
    return _layer_norm(h + o, sd, 'transf.mlp.layer_norm', 1e-6)


def forward(sd, conf, mem_patch, mem_pos=None, train=True):
    """ips_net.py:264-283 with dropout disabled.  ``train`` = BN in
    batch-statistics mode as under net.train()."""
    B, M = mem_patch.shape[:2]
    emb = encode(sd, conf, mem_patch.reshape(-1, *mem_patch.shape[2:]), train=train).view(B, M, -1)
    if torch.is_tensor(mem_pos):
        emb = emb + mem_pos
    tok = aggregate(sd, conf, emb)
    preds = {}
    for task in conf.tasks.values():
        z = F.linear(tok[:, task['id']], sd[f"output_layers.{task['name']}.0.weight"],
                     sd[f"output_layers.{task['name']}.0.bias"])
        preds[task['name']] = torch.softmax(z, dim=-1) if task['act_fn'] == 'softmax' else torch.sigmoid(z)
    return preds


def loss_fn(conf, preds, labels):
    """training/iterative.py:75-98 with main.py:53-61 criteria: NLL on
    log(p + eps) for softmax heads, BCE for sigmoid heads, mean over tasks."""
    total = 0
    for task in conf.tasks.values():
        p = preds[task['name']].squeeze(-1)
        y = labels[task['name']]
        if task['act_fn'] == 'softmax':
            total = total + F.nll_loss(torch.log(p + conf.eps), y)
        else:
            total = total + F.binary_cross_entropy(p.view(-1), y.view(-1).float())
    return total / len(conf.tasks)


def make_labels(conf, B, seed):
    g = torch.Generator().manual_seed(int(seed))
    out = {}
    for task in conf.tasks.values():
        if task['metric'] == 'multilabel_accuracy':
            out[task['name']] = (torch.rand(B, conf.n_class, generator=g) < 0.3).float()
        elif task['act_fn'] == 'sigmoid':
            out[task['name']] = torch.randint(0, 2, (B,), generator=g)
        else:
            out[task['name']] = torch.randint(0, conf.n_class, (B,), generator=g)
    return out


# --------------------------------------------------------------------------
# sequence-sharded schedule: the reference loop per slice,
# then one score_and_select over the concatenated candidates.
# --------------------------------------------------------------------------

def shard_bounds(N, R):
    """Contiguous slices of the (shuffled) patch axis, sizes differ by <= 1."""
    base, rem = divmod(N, R)
    out, lo = [], 0
    for r in range(R):
        hi = lo + base + (1 if r < rem else 0)
        out.append((lo, hi))
        lo = hi
    return out


def ips_sharded(sd, conf, patches, R, perm=None, tie='stable'):
    """Each of R ranks scans its slice keeping a local top-M (a slice no longer
    than M keeps everything, the M>=N shortcut), then every rank re-scores the
    gathered candidates in ONE buffer and keeps the global top-M."""
    with torch.no_grad():
        M, D = conf.M, conf.D
        B, N = patches.shape[:2]
        tail = patches.shape[2:]
        if perm is None:
            perm = torch.arange(N).unsqueeze(0).expand(B, -1)
        table = pos_table(D, conf.N) if conf.use_pos else None
        cand = []
        for lo, hi in shard_bounds(N, R):
            sub_perm = perm[:, lo:hi]
            sub = torch.gather(patches, 1, sub_perm.reshape(B, -1, *([1] * len(tail))).expand(-1, -1, *tail))
            n = hi - lo
            if n <= M:
                cand.append(sub_perm)
                continue
            sub_conf = conf
            if conf.use_pos:
                # local scan must see the patches' true positions
                _, _, src = _ips_with_pos(sd, conf, sub, table[sub_perm], tie)
            else:
                _, _, src = ips(sd, sub_conf, sub, perm=None, tie=tie)
            cand.append(torch.gather(sub_perm, 1, src))
        cand = torch.cat(cand, dim=1)                              # (B, <=R*M) original indices
        cp = torch.gather(patches, 1, cand.reshape(B, -1, *([1] * len(tail))).expand(-1, -1, *tail))
        emb = encode(sd, conf, cp.reshape(-1, *tail)).view(B, cand.shape[1], -1)
        emb_pos = emb + table[cand] if conf.use_pos else None
        if cand.shape[1] <= M:
            return cp, (table[cand] if conf.use_pos else None), cand
        _, win, _ = score_and_select(sd, conf, emb, emb_pos, M, cand, tie)
        mem_patch = torch.gather(patches, 1, win.reshape(B, -1, *([1] * len(tail))).expand(-1, -1, *tail))
        return mem_patch, (table[win] if conf.use_pos else None), win


def _ips_with_pos(sd, conf, px, ppos, tie):
    """ips() body for an explicit per-patch position table (B,n,D)."""
    M, I, D = conf.M, conf.I, conf.D
    B, N = px.shape[:2]
    tail = px.shape[2:]
    mem_emb = encode(sd, conf, px[:, :M].reshape(-1, *tail)).view(B, M, -1)
    idx = torch.arange(N, dtype=torch.int64).unsqueeze(0).expand(B, -1)
    mem_idx = idx[:, :M]
    for it in range(math.ceil((N - M) / I)):
        lo = it * I + M
        hi = min(lo + I, N)
        new_emb = encode(sd, conf, px[:, lo:hi].reshape(-1, *tail)).view(B, -1, D)
        all_emb = torch.cat((mem_emb, new_emb), dim=1)
        all_idx = torch.cat((mem_idx, idx[:, lo:hi]), dim=1)
        all_pos = all_emb + torch.gather(ppos, 1, all_idx.view(B, -1, 1).expand(-1, -1, D))
        mem_emb, mem_idx, _ = score_and_select(sd, conf, all_emb, all_pos, M, all_idx, tie)
    return None, None, mem_idx

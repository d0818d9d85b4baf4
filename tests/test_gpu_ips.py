"""End-to-end parity of IPSNet.ips on the B200 against the CPU oracle and the golden
fixtures produced by the unmodified reference (tests/golden, oracle/gen_golden.py)."""
import numpy as np
import pytest
import torch

import ips_oracle as O
from golden_util import CASE_NAMES, load_case

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def _net(conf, sd, precision):
    from ips_b200 import IPSNet, Struct
    c = conf.replace(precision=precision)
    net = IPSNet(torch.device(DEV), Struct(**c.__dict__)).to(DEV)
    net.load_state_dict(sd, strict=True)
    net.train()                                   # ips() is called in train mode by training/iterative.py:109,135
    return net


def _boundary_report(sd, conf, patches, rng_seed, got_src):
    """P3: every index the GPU picked that the oracle did not must have an oracle
    score within `tol` of the rank-M boundary of the final iteration."""
    trace = []
    torch.manual_seed(rng_seed)
    _, _, o_src = O.ips(sd, conf, patches, perm='draw', tie='stable', trace=trace)
    worst = 0.0
    for b in range(o_src.shape[0]):
        a, g = set(o_src[b].tolist()), set(got_src[b].tolist())
        if a != g:
            s = trace[-1][0][b].sort(descending=True)[0]
            worst = max(worst, float((s[conf.M - 1] - s[conf.M]).abs() / s[conf.M - 1]))
    return o_src, worst


@pytest.mark.parametrize('name', CASE_NAMES)
def test_ips_fp32_matches_golden(name):
    """fp32 scoring mode: selected indices identical to the reference's (golden) selection."""
    z, meta, conf, sd, patches = load_case(name)
    net = _net(conf, sd, 'fp32')
    torch.manual_seed(meta['rng_seed'])
    # 'instance' shuffling draws torch.rand on the DATA's device (utils/utils.py:48): the golden was produced with the
    # patches on the host, so the same RNG stream is reached through the lazy (host-resident) input path
    mem_patch, mem_pos = net.ips(patches if name == 'mnist_instance' else patches.to(DEV))
    assert net.training and net.encoder.training and net.transf.training
    assert list(mem_patch.shape) == list(z['mem_patch_shape'])
    if conf.M >= meta['N']:                                       # shortcut path
        assert torch.equal(mem_patch.cpu(), patches)
        return
    got = net.last_mem_idx.cpu()
    # the output really is the selected patches, in the returned order
    ref_rows = torch.stack([patches[b, got[b]] for b in range(got.shape[0])])
    assert torch.equal(mem_patch.cpu(), ref_rows)
    if conf.use_pos:
        tab = O.pos_table(conf.D, conf.N)
        assert torch.equal(mem_pos.cpu(), tab[got])
    else:
        assert mem_pos is None
    if 'mem_src' in z and z['mem_src'].size:
        gold = torch.from_numpy(z['mem_src'])
        same_set = all(set(gold[b].tolist()) == set(got[b].tolist()) for b in range(gold.shape[0]))
        assert same_set, f'selection differs from the reference; oracle boundary gap {meta["boundary_gap"]}'
        if 'ties' not in name:
            assert torch.equal(got, gold), 'order differs although no ties are expected'
        assert mem_patch.double().sum().item() == pytest.approx(float(z['mem_patch_sum']), rel=1e-12)


@pytest.mark.parametrize('name', [n for n in CASE_NAMES if 'ties' not in n and 'instance' not in n])
def test_ips_bf16x3_matches_golden(name):
    """'bf16x3' (operands as hi + lo bf16 pairs, three tensor-core products per multiply, fp32 sums): the selection equals the
    reference's on the conditioned fixtures, and the logits agree with the fp32 oracle to ~1e-4 of their scale."""
    z, meta, conf, sd, patches = load_case(name)
    net = _net(conf, sd, 'bf16x3')
    B, N = patches.shape[:2]
    if conf.M < N:
        zt = net.patch_logits(patches.to(DEV)).cpu()
        emb = O.encode(sd, conf, patches.reshape(B * N, *patches.shape[2:])).view(B, N, -1)
        if conf.use_pos:
            emb = emb + O.pos_table(conf.D, conf.N)
        ref = O.attn_logits(sd, conf, emb).permute(0, 3, 1, 2).reshape(B, N, -1)
        err = (zt - ref).abs().max().item() / ref.abs().max().item()
        print(f'{name}: bf16x3 logit max err / max |logit| = {err:.3e}')
        assert err < 3e-4
    torch.manual_seed(meta['rng_seed'])
    mem_patch, mem_pos = net.ips(patches.to(DEV))
    if conf.M >= N:
        assert torch.equal(mem_patch.cpu(), patches)
        return
    got = net.last_mem_idx.cpu()
    gold = torch.from_numpy(z['mem_src'])
    assert torch.equal(got, gold), f'selection differs from the reference; oracle boundary gap {meta["boundary_gap"]}'


@pytest.mark.parametrize('name', ['mnist_ties'])
def test_ips_fp32_tie_fixture(name):
    """Unconditioned, 90 % all-zero patches: scores differ only through the pos-enc; report P3."""
    z, meta, conf, sd, patches = load_case(name)
    net = _net(conf, sd, 'fp32')
    torch.manual_seed(meta['rng_seed'])
    net.ips(patches.to(DEV))
    got = net.last_mem_idx.cpu()
    o_src, worst = _boundary_report(sd, conf, patches, meta['rng_seed'], got)
    assert worst < 1e-5, f'picks differ where the oracle boundary gap is {worst}'


def _oracle_final_buffer(conf, trace, perm, B):
    """(original index, oracle score) of every entry of the LAST iteration's buffer, and the rank-M boundary scores."""
    M, I = conf.M, conf.I
    n_iter = len(trace)
    N = perm.shape[1]
    lo = M + (n_iter - 1) * I
    hi = min(lo + I, N)
    prev_mem = trace[-2][1] if n_iter > 1 else torch.arange(M).unsqueeze(0).expand(B, -1)
    buf_pos = torch.cat((prev_mem, torch.arange(lo, hi).unsqueeze(0).expand(B, -1)), dim=1)
    return torch.gather(perm, 1, buf_pos), trace[-1][0]


@pytest.mark.parametrize('pre,over,B,N', [('camelyon', {}, 1, 50000),          # C4, full size: M = I = 5000, 9 iterations
                                          ('mnist', {}, 2, 900),               # C1: 8 iterations, pos-enc
                                          ('traffic', {}, 2, 192),             # C2: 6 iterations, ragged last chunk (22 of 32)
                                          ('mnist', {'N': 10000}, 1, 10000)])  # C3: 99 iterations
@pytest.mark.parametrize('precision', ['fp32', 'bf16x3'])
def test_ips_fp32_matches_oracle_at_baseline_size(pre, over, B, N, precision):
    """BASELINE.json sizes (per-image shapes exactly as the configs; B reduced so the CPU oracle finishes in seconds):
    fp32-mode `ips()` against `O.ips` (ips_net.py:169-262) on the same weights, inputs and scan order.  Indices must be
    identical; where they are not, P3 applies -- a differing pick must sit within fp32 rounding of the
    oracle's rank-M boundary score -- and the report is printed."""
    conf = O.preset(pre, attn_dropout=0.0, dropout=0.0, **over)
    sd = O.make_state(conf, 91, q_gain=12.0)
    patches = O.make_patches(conf, B, N, 92)
    net = _net(conf, sd, precision)               # 'bf16x3': the same comparison for the near-fp32 tensor-core mode
    torch.manual_seed(93)
    mem_patch, mem_pos = net.ips(patches.to(DEV))
    got = net.last_mem_idx.cpu()
    torch.manual_seed(93)
    perm = torch.randperm(N).unsqueeze(0).expand(B, -1)           # the 'batch' shuffle's draw (utils/utils.py:38)
    trace = []
    o_patch, o_pos, o_src = O.ips(sd, conf, patches, perm=perm, tie='stable', trace=trace)
    assert len(trace) == -(-(N - conf.M) // conf.I)
    for b in range(B):                                            # the output really is the selected patches, in order
        assert torch.equal(mem_patch[b].cpu(), patches[b, got[b]])
    if conf.use_pos:
        assert torch.equal(mem_pos.cpu(), O.pos_table(conf.D, conf.N)[got])
    n_diff, worst, outside = 0, 0.0, 0
    buf_idx, buf_score = _oracle_final_buffer(conf, trace, perm, B)
    for b in range(B):
        a, g = set(o_src[b].tolist()), set(got[b].tolist())
        if a == g:
            continue
        srt = buf_score[b].sort(descending=True)[0]
        boundary = 0.5 * float(srt[conf.M - 1] + srt[conf.M])
        score_of = dict(zip(buf_idx[b].tolist(), buf_score[b].tolist()))
        for i in a ^ g:
            n_diff += 1
            if i in score_of:
                worst = max(worst, abs(score_of[i] - boundary) / boundary)
            else:
                outside += 1                                      # evicted earlier in the oracle's run (an earlier boundary flip)
    order_equal = torch.equal(got, o_src)
    print(f'{pre} B={B} N={N} {precision}: identical={order_equal} differing picks={n_diff} (of {B * conf.M}), '
          f'max |score - boundary| / boundary = {worst:.2e}, not in the oracle\'s last buffer: {outside}')
    tol = 2e-5 if precision == 'fp32' else 5e-4                   # bf16x3 carries ~17 significant bits per operand
    assert n_diff <= (2 if precision == 'fp32' else 10) * max(1, B * conf.M // 1000), f'{n_diff} differing picks'
    assert worst < tol, f'a differing pick sits {worst:.2e} (relative) away from the oracle boundary'
    if n_diff == 0 and not order_equal:                           # same set, different order: only exact-score neighbours may swap
        sc = {b: dict(zip(buf_idx[b].tolist(), buf_score[b].tolist())) for b in range(B)}
        for b in range(B):
            for m in (got[b] != o_src[b]).nonzero().flatten().tolist():
                x, y = sc[b][int(got[b, m])], sc[b][int(o_src[b, m])]
                assert abs(x - y) <= tol * abs(y), (b, m, x, y)     # measured on C4: 3e-6 (GPU expf vs CPU exp)


@pytest.mark.parametrize('name', ['mnist_small', 'traffic_small', 'camelyon_small', 'camelyon_batch'])
def test_ips_bf16_close(name):
    """bf16 tensor-core mode: logits within tolerance of the fp32 oracle, selection overlaps."""
    z, meta, conf, sd, patches = load_case(name)
    net = _net(conf, sd, 'bf16')
    B, N = patches.shape[:2]
    zt = net.patch_logits(patches.to(DEV)).cpu()                    # (B,N,HT)
    emb = O.encode(sd, conf, patches.reshape(B * N, *patches.shape[2:])).view(B, N, -1)
    if conf.use_pos:
        emb = emb + O.pos_table(conf.D, conf.N)
    ref = O.attn_logits(sd, conf, emb).permute(0, 3, 1, 2).reshape(B, N, -1)
    err = (zt - ref).abs().max().item() / ref.abs().max().item()
    big = ref.abs() >= 0.05 * ref.abs().max()                       # element-wise relative error where the logit is not ~0
    rel = ((zt - ref).abs()[big] / ref.abs()[big])
    print(f'{name}: bf16 logit max err / max |logit| = {err:.3e}; element-wise relative error (|logit| >= 5% of max): '
          f'median {rel.median().item():.3e}, p99 {rel.quantile(0.99).item():.3e}, max {rel.max().item():.3e}')
    assert err < 3e-2
    assert rel.median().item() < 2e-2
    torch.manual_seed(meta['rng_seed'])
    net.ips(patches.to(DEV))
    got = net.last_mem_idx.cpu()
    torch.manual_seed(meta['rng_seed'])
    _, _, o_src = O.ips(sd, conf, patches, perm='draw', tie='stable')
    overlap = np.mean([len(set(got[b].tolist()) & set(o_src[b].tolist())) / conf.M for b in range(B)])
    print(f'{name}: bf16 selection overlap with fp32 oracle = {overlap:.3f}')
    assert overlap >= 0.85


@pytest.mark.parametrize('name,precision,over', [('mnist_small', 'fp32', {}),
                                                 # ring of three chunk buffers: no whole-tensor device copy (O(chunk) memory)
                                                 ('mnist_small', 'fp32', dict(lazy_resident_bytes=0, chunk_patches=5)),
                                                 ('traffic_small', 'bf16', dict(lazy_resident_bytes=0, chunk_patches=3)),
                                                 ('camelyon_batch', 'bf16', dict(lazy_resident_bytes=0, chunk_patches=100))])
def test_ips_lazy_host_input_equals_eager(name, precision, over):
    z, meta, conf, sd, patches = load_case(name)
    net = _net(conf.replace(eager=False, **over), sd, precision)
    torch.manual_seed(1)
    a_patch, a_pos = net.ips(patches)                              # host tensor: streamed
    a_idx = net.last_mem_idx.clone()
    torch.manual_seed(1)
    b_patch, b_pos = net.ips(patches.to(DEV))
    assert a_patch.is_cuda and torch.equal(a_idx, net.last_mem_idx)
    assert torch.equal(a_patch, b_patch) and ((a_pos is None and b_pos is None) or torch.equal(a_pos, b_pos))


def test_ips_does_not_touch_bn_stats_or_mode():
    z, meta, conf, sd, patches = load_case('mnist_small')
    net = _net(conf, sd, 'fp32')
    before = {k: v.clone() for k, v in net.state_dict().items()}
    net.ips(patches.to(DEV))
    for k, v in net.state_dict().items():
        assert torch.equal(v, before[k]), k
    net.eval()
    net.ips(patches.to(DEV))
    assert not net.training


def test_score_and_select_matches_reference_unit():
    z, meta, conf, sd, patches = load_case('camelyon_small')
    net = _net(conf, sd, 'fp32')
    g = torch.Generator().manual_seed(101)
    L = min(conf.M + conf.I, 64)
    emb = torch.randn(meta['B'], L, conf.D, generator=g)
    idx = torch.arange(L).unsqueeze(0).expand(meta['B'], -1).contiguous()
    _, si = net.score_and_select(emb.to(DEV), None, max(1, L // 3), idx.to(DEV))
    assert np.array_equal(si.cpu().numpy(), z['unit_select_idx'])
    torch.testing.assert_close(net.transf.get_scores(emb.to(DEV)).cpu(), torch.from_numpy(z['unit_emb_scores']),
                               rtol=2e-4, atol=1e-8)


@pytest.mark.parametrize('pre,B,N', [('traffic', 16, 192), ('camelyon', 1, 50000), ('mnist', 4, 900)])
def test_full_size_properties(pre, B, N):
    """BASELINE.json sizes: size-independent properties (the oracle would take minutes)."""
    conf = O.preset(pre, attn_dropout=0.0, dropout=0.0)
    sd = O.make_state(conf, 31, q_gain=12.0)
    net = _net(conf, sd, 'bf16')
    g = torch.Generator(device=DEV).manual_seed(32)
    shape = (B, N, conf.n_chan_in, *conf.patch_size) if conf.is_image else (B, N, conf.n_chan_in)
    x = torch.randn(shape, generator=g, device=DEV)
    torch.manual_seed(3)
    mem_patch, mem_pos = net.ips(x)
    idx = net.last_mem_idx
    assert mem_patch.shape[:2] == (B, conf.M)
    for b in range(B):
        assert idx[b].unique().numel() == conf.M                   # no patch selected twice
        assert torch.equal(mem_patch[b], x[b, idx[b]])            # rows are the selected patches, bit exact
    # idempotence: selecting again from the winners (N' = M) is the shortcut and returns them
    again, _ = net.ips(mem_patch)
    assert torch.equal(again, mem_patch)
    # the winners survive a rescan that presents them first, without shuffling
    net.shuffle = False
    rest = torch.stack([x[b, torch.randperm(N, device=DEV)[: conf.I]] for b in range(B)])
    net.ips(torch.cat([mem_patch, rest], dim=1))
    assert net.last_mem_idx.shape == (B, conf.M)
    # scan-order invariance of the selected SET when H = n_token = 1
    conf1 = conf.replace(H=1, n_token=1, tasks={'t': conf.tasks['task0']}) if pre == 'camelyon' else None
    if conf1 is not None:
        sd1 = O.make_state(conf1, 33, q_gain=12.0)
        n1 = _net(conf1, sd1, 'bf16')
        torch.manual_seed(1)
        n1.ips(x)
        a = n1.last_mem_idx.sort(-1)[0]
        torch.manual_seed(2)
        n1.ips(x)
        assert torch.equal(a, n1.last_mem_idx.sort(-1)[0])


@pytest.mark.parametrize('name,precision', [('mnist_small', 'fp32'), ('traffic_small', 'bf16'), ('mnist_small', 'bf16')])
def test_native_executor_equals_per_layer_calls(name, precision):
    """The C++ executor issues the same kernels as the per-layer Python path: identical logits."""
    z, meta, conf, sd, patches = load_case(name)
    net = _net(conf.replace(chunk_patches=5), sd, precision)       # several ragged chunks
    x = patches.to(DEV)
    net.executor = 'native'
    a = net.patch_logits(x)
    net.executor = 'python'
    b = net.patch_logits(x)
    assert torch.equal(a, b)


@pytest.mark.parametrize('pre,rows', [('traffic', 700), ('mnist', 1900)])
def test_lanes_and_chunking_do_not_change_logits(pre, rows):
    """A patch's logits are a property of the patch: the same bits whatever the chunk size, the number of chunks in
    flight (internal streams), the tile a row lands in, single CTA or CTA pair."""
    conf = O.preset(pre, attn_dropout=0.0, dropout=0.0)
    sd = O.make_state(conf, 41, q_gain=12.0)
    g = torch.Generator(device=DEV).manual_seed(42)
    x = torch.randn((1, rows, conf.n_chan_in, *conf.patch_size), generator=g, device=DEV)
    ref = None
    for lanes, chunk in ((1, 0), (2, 0), (3, 100), (2, 257), (4, 64)):
        net = _net(conf.replace(lanes=lanes, chunk_patches=chunk, N=rows), sd, 'bf16')
        z = net.patch_logits(x)
        z2 = net.patch_logits(x)                   # second call: reused workspace, no re-zeroing
        assert torch.equal(z, z2)
        if ref is None:
            ref = z
        assert torch.equal(z, ref), (lanes, chunk)


@pytest.mark.parametrize('name', [n for n in CASE_NAMES if 'shortcut' not in n and 'instance' not in n])
@pytest.mark.parametrize('precision', ['fp32', 'bf16'])
def test_inference_forward_matches_oracle(name, precision):
    """`net.eval(); net(mem_patch, mem_pos)` under no_grad (the reference's evaluate()) runs on the library's
    kernels end to end and reproduces the oracle's probabilities (eval-mode BatchNorm, dropout off)."""
    z, meta, conf, sd, patches = load_case(name)
    net = _net(conf, sd, precision)
    torch.manual_seed(meta['rng_seed'])
    mem_patch, mem_pos, _ = O.ips(sd, conf, patches, perm='draw', tie='topk')
    ref = O.forward(sd, conf, mem_patch, mem_pos, train=False)
    net.eval()
    with torch.no_grad():
        got = net(mem_patch.to(DEV), None if mem_pos is None else mem_pos.to(DEV))
    tol = dict(rtol=2e-4, atol=2e-6) if precision == 'fp32' else dict(rtol=5e-2, atol=5e-3)
    for t in ref:
        torch.testing.assert_close(got[t].cpu(), ref[t], **tol)
    assert not net.training


@pytest.mark.parametrize('name', ['camelyon_small', 'camelyon_batch', 'mnist_small', 'traffic_small'])
@pytest.mark.parametrize('precision', ['fp32', 'bf16'])
def test_train_step_gradients_match_oracle(name, precision):
    """Grad-mode forward + loss + backward (train-mode BatchNorm, dropout 0): every nn.Linear runs forward AND
    backward on the library's GEMM kernels (NT / NN / TN); loss and gradients against the CPU oracle."""
    z, meta, conf, sd, patches = load_case(name)
    net = _net(conf, sd, precision)
    torch.backends.cudnn.allow_tf32 = False          # the conv encoder's grad-mode half is still cuDNN: keep it fp32
    torch.manual_seed(meta['rng_seed'])
    mem_patch, mem_pos, _ = O.ips(sd, conf, patches, perm='draw', tie='topk')
    labels = O.make_labels(conf, meta['B'], meta['label_seed'])
    preds = net(mem_patch.to(DEV), None if mem_pos is None else mem_pos.to(DEV))
    loss = O.loss_fn(conf, preds, {k: v.to(DEV) for k, v in labels.items()})
    loss.backward()
    ref_loss = float(z['loss'])
    tol = 1e-4 if precision == 'fp32' else 2e-2
    assert abs(loss.item() - ref_loss) <= tol * max(1.0, abs(ref_loss))
    grads = dict(net.named_parameters())
    for key in z.files:
        if not key.startswith('gradnorm_'):
            continue
        g = grads[key[9:]].grad
        got, want = g.double().norm().item(), float(z[key])
        assert abs(got - want) <= (1e-3 if precision == 'fp32' else 6e-2) * max(want, 1e-6), (key, got, want)
        if precision == 'fp32':
            want_v = z['grad_' + key[9:]]
            # element-wise, relative to the slice's scale (train-mode BatchNorm over a handful of patches amplifies
            # the cuDNN-vs-CPU rounding of the conv encoder in individual small entries)
            np.testing.assert_allclose(g.reshape(-1)[:256].cpu().numpy(), want_v, rtol=2e-3, atol=2e-2 * float(np.abs(want_v).max()))


def _unfold(images, ph, pw, sh, sw):
    """The reference's CPU patchify (mnist_dataset.py:47-53), batched."""
    p = images.unfold(2, ph, sh).unfold(3, pw, sw)              # (B, C, nr, nc, ph, pw)
    B, C, nr, nc = p.shape[:4]
    return p.permute(0, 2, 3, 1, 4, 5).reshape(B, nr * nc, C, ph, pw).contiguous()


@pytest.mark.parametrize('pre,img,stride,precision', [('traffic', (300, 500), None, 'bf16'), ('mnist', (250, 300), (25, 25), 'bf16'),
                                                      ('mnist', (250, 300), (25, 25), 'fp32'), ('traffic', (301, 507), (67, 101), 'bf16')])
def test_ips_image_equals_ips_of_unfolded_patches(pre, img, stride, precision):
    """On-device patchify: reading the patch grid straight out of the images gives the same winners,
    bit for bit, as ips() on the reference's unfolded patch tensor -- also for overlapping and odd strides."""
    conf = O.preset(pre, attn_dropout=0.0, dropout=0.0)
    ph, pw = conf.patch_size
    sh, sw = stride or (ph, pw)
    B = 3
    g = torch.Generator(device=DEV).manual_seed(52)
    images = torch.randn((B, conf.n_chan_in, *img), generator=g, device=DEV)
    patches = _unfold(images, ph, pw, sh, sw)
    N = patches.shape[1]
    conf = conf.replace(N=N, M=min(conf.M, 20), I=min(conf.I, 20))
    sd = O.make_state(conf, 51, q_gain=12.0)
    net = _net(conf, sd, precision)
    from ips_b200 import ops
    geo, n = ops.image_geo(images, (ph, pw), (sh, sw))
    assert n == N and torch.equal(ops.gather_patches_image(images, geo, None, (ph, pw)), patches)
    torch.manual_seed(5)
    a_patch, a_pos = net.ips(patches)
    a_idx = net.last_mem_idx.clone()
    torch.manual_seed(5)
    b_patch, b_pos = net.ips_image(images, (ph, pw), (sh, sw))
    assert torch.equal(a_idx, net.last_mem_idx)
    assert torch.equal(a_patch, b_patch)
    assert (a_pos is None and b_pos is None) or torch.equal(a_pos, b_pos)


def test_ips_writes_into_train_buffers():
    """Batch assembly: ips(out=..., row_offset=n_prep) fills the init_batch buffers in place
    exactly as fill_batch (training/iterative.py:31-50) would."""
    z, meta, conf, sd, patches = load_case('mnist_small')
    net = _net(conf, sd, 'fp32')
    B = patches.shape[0]
    x = patches.to(DEV)
    torch.manual_seed(2)
    ref_patch, ref_pos = net.ips(x)
    buf_patch = torch.zeros((2 * B + 1, conf.M, *x.shape[2:]), device=DEV)
    buf_pos = torch.zeros((2 * B + 1, conf.M, conf.D), device=DEV)
    torch.manual_seed(2)
    got_patch, got_pos = net.ips(x, out=(buf_patch, buf_pos), row_offset=B)
    assert got_patch.data_ptr() == buf_patch[B].data_ptr()
    assert torch.equal(buf_patch[B:2 * B], ref_patch) and torch.equal(buf_pos[B:2 * B], ref_pos)
    assert float(buf_patch[:B].abs().max()) == 0.0 and float(buf_patch[2 * B:].abs().max()) == 0.0


@pytest.mark.parametrize('name', ['traffic_small', 'mnist_small', 'camelyon_batch'])
def test_graphed_train_step_equals_eager(name):
    """The CUDA-graph train step (forward + loss + backward + AdamW) replays to the same parameters as the eager
    step from the same state, and keeps doing so on new buffer contents."""
    from ips_b200.train import GraphedTrainStep, compute_loss
    z, meta, conf, sd, patches = load_case(name)
    conf = conf.replace(attn_dropout=0.0, dropout=0.0)
    B = meta['B']

    def fresh():
        net = _net(conf, sd, 'bf16')
        opt = torch.optim.AdamW(net.parameters(), lr=1e-3, weight_decay=conf.wd, capturable=True)
        return net, opt

    batches = []
    g = torch.Generator().manual_seed(77)
    for i in range(2):
        torch.manual_seed(meta['rng_seed'] + i)
        mp, mpos, _ = O.ips(sd, conf, patches + 0.05 * i * torch.randn(patches.shape, generator=g), perm='draw', tie='topk')
        batches.append((mp.to(DEV), None if mpos is None else mpos.to(DEV),
                        {k: v.to(DEV) for k, v in O.make_labels(conf, B, meta['label_seed'] + i).items()}))

    net_e, opt_e = fresh()
    losses_e = []
    for mp, mpos, lab in batches:
        opt_e.zero_grad(set_to_none=False)
        loss = compute_loss(conf, net_e(mp, mpos), lab)
        loss.backward()
        opt_e.step()
        losses_e.append(loss.item())

    net_g, opt_g = fresh()
    step = GraphedTrainStep(net_g, conf, opt_g, B)
    step.mem_patch.copy_(batches[0][0])
    step.capture()
    losses_g = []
    for mp, mpos, lab in batches:
        step.mem_patch.copy_(mp)
        if mpos is not None:
            step.mem_pos.copy_(mpos)
        for k, v in lab.items():
            step.labels[k].copy_(v)
        losses_g.append(step().item())
    assert np.allclose(losses_g, losses_e, rtol=1e-4, atol=1e-5), (losses_g, losses_e)
    for (k, a), b in zip(net_e.state_dict().items(), net_g.state_dict().values()):
        torch.testing.assert_close(b.float(), a.float(), rtol=1e-3, atol=1e-5, msg=k)
    # ips() after graph replays selects with the TRAINED weights: a replay moves parameters and BatchNorm statistics
    # behind the tensors' version counters, so the step invalidates the folded-parameter plan
    x = patches.to(DEV)
    z_init = _net(conf, sd, 'bf16').patch_logits(x)
    z_g, z_e = net_g.patch_logits(x), net_e.patch_logits(x)
    scale = float(z_e.abs().max())
    assert float((z_g - z_e).abs().max()) <= 5e-2 * scale
    assert float((z_g - z_init).abs().max()) > 1e-3 * scale, 'ips() still uses the weights from before training'


@pytest.mark.parametrize('patch', [(256, 320), (30, 18), (240, 64)])
def test_ips_unusual_patch_sizes(patch):
    """Patch sizes beyond the shipped configs: wide patches fall back from the shifted-window stem to the gather
    stem; the bf16 selection still agrees with the fp32 CPU oracle on a conditioned fixture."""
    conf = O.preset('traffic', attn_dropout=0.0, dropout=0.0, patch_size=list(patch), N=24, M=4, I=6)
    sd = O.make_state(conf, 61, q_gain=12.0)
    patches = O.make_patches(conf, 2, 24, 62)
    net = _net(conf, sd, 'bf16')
    torch.manual_seed(9)
    mp, _ = net.ips(patches.to(DEV))
    got = net.last_mem_idx.cpu()
    torch.manual_seed(9)
    _, _, o_src = O.ips(sd, conf, patches, perm='draw', tie='stable')
    overlap = np.mean([len(set(got[b].tolist()) & set(o_src[b].tolist())) / conf.M for b in range(2)])
    assert overlap >= 0.75, overlap
    for b in range(2):
        assert torch.equal(mp[b].cpu(), patches[b, got[b]])


@pytest.mark.parametrize('name', ['mnist_small', 'traffic_small', 'camelyon_batch'])
def test_fused_head_loss_matches_compute_loss(name):
    """IPSNet.loss (heads + NLL/BCE + gradient in one kernel per task) == compute_loss on the
    probabilities of forward(): value and every parameter gradient."""
    from ips_b200.train import compute_loss
    z, meta, conf, sd, patches = load_case(name)
    conf = conf.replace(attn_dropout=0.0, dropout=0.0)
    torch.manual_seed(meta['rng_seed'])
    mem_patch, mem_pos, _ = O.ips(sd, conf, patches, perm='draw', tie='topk')
    labels = {k: v.to(DEV) for k, v in O.make_labels(conf, meta['B'], meta['label_seed']).items()}
    mp, mpos = mem_patch.to(DEV), None if mem_pos is None else mem_pos.to(DEV)
    res = []
    for fused in (False, True):
        net = _net(conf, sd, 'fp32')
        loss = net.loss(mp, mpos, labels, conf.eps) if fused else compute_loss(conf, net(mp, mpos), labels)
        loss.backward()
        res.append((loss.item(), {k: p.grad.clone() for k, p in net.named_parameters() if p.grad is not None}))
    assert abs(res[0][0] - res[1][0]) <= 1e-5 * max(1.0, abs(res[0][0]))
    assert res[0][1].keys() == res[1][1].keys()
    scale = max(float(g.abs().max()) for g in res[0][1].values())      # (some gradients are pure rounding noise: bias before BN)
    for k in res[0][1]:
        a, b = res[0][1][k], res[1][1][k]
        assert (a - b).abs().max() <= 1e-4 * float(a.abs().max()) + 1e-6 * scale, k


def test_bf16_feature_bag_equals_upcast_fp32():
    """bf16-stored feature bags: ips() on the bf16 tensor -- device-resident or streamed from the host --
    selects exactly what it selects on the same values upcast to fp32, and returns the winners as fp32."""
    conf = O.preset('camelyon', attn_dropout=0.0, dropout=0.0, M=50, I=70)
    sd = O.make_state(conf, 71, q_gain=12.0)
    xb = O.make_patches(conf, 2, 1000, 72).to(torch.bfloat16)
    net = _net(conf, sd, 'bf16')
    torch.manual_seed(4)
    ref_patch, _ = net.ips(xb.float().to(DEV))
    ref_idx = net.last_mem_idx.clone()
    for src in (xb.to(DEV), xb.pin_memory()):
        torch.manual_seed(4)
        got_patch, _ = net.ips(src)
        assert torch.equal(net.last_mem_idx, ref_idx)
        assert got_patch.dtype == torch.float32 and torch.equal(got_patch, ref_patch)


def test_native_executor_dense_group_equals_per_layer_calls():
    """With an even number of full 128-pixel tiles per chunk the native executor runs layer 3 (7x7 maps, 256 channels) on dense
    activations through the im2col-mode kernel; the logits equal the per-layer padded-flat path bit for bit."""
    conf = O.preset('traffic', attn_dropout=0.0, dropout=0.0)
    sd = O.make_state(conf, 81, q_gain=12.0)
    g = torch.Generator(device=DEV).manual_seed(82)
    x = torch.randn((2, 256, conf.n_chan_in, *conf.patch_size), generator=g, device=DEV)
    net = _net(conf.replace(chunk_patches=256, lanes=2, N=256), sd, 'bf16')
    net.executor = 'native'
    a = net.patch_logits(x)
    net.executor = 'python'
    b = net.patch_logits(x)
    assert torch.equal(a, b)

"""Unit-level parity of every CUDA kernel against the CPU oracle / ATen CPU operators.
All calls go through the C ABI (ips_b200.ops -> ctypes -> libips_b200.so)."""
import math

import pytest
import torch
import torch.nn.functional as F

import ips_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def dev():
    from ips_b200 import _lib
    lib = _lib.load()
    _lib.check(lib.ipsb_device_ok())
    return torch.device('cuda:0')


def _rand(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g) * scale


# ------------------------------------------------------------------ data movement

@pytest.mark.parametrize('C,H,W', [(1, 50, 50), (3, 100, 100), (3, 7, 9)])
@pytest.mark.parametrize('dt', ['f32', 'bf16'])
def test_stage_patches(dev, C, H, W, dt):
    from ips_b200 import ops
    x = _rand(11, C, H, W, seed=1)
    code = ops.F32 if dt == 'f32' else ops.BF16
    idx = torch.tensor([5, 0, 10, 3, 3], dtype=torch.int64)
    ref = torch.zeros(11, H, W, 4)
    ref[..., :C] = x.permute(0, 2, 3, 1)
    if dt == 'bf16':
        ref = ref.to(torch.bfloat16)
    out = ops.stage_patches(x.to(dev), 5, C, H, W, code, row_idx=idx.to(dev))
    assert torch.equal(out.cpu(), ref[idx])                       # gather + layout + cast are exact
    out = ops.stage_patches(x.to(dev), 4, C, H, W, code, first_row=6)
    assert torch.equal(out.cpu(), ref[6:10])


@pytest.mark.parametrize('C,H,W,P', [(3, 100, 100, 37), (1, 52, 48, 9), (4, 20, 36, 5), (3, 100, 100, 700), (2, 6, 4, 3), (3, 160, 160, 2)])
def test_stage_s2d_tma_equals_definition(dev, C, H, W, P, monkeypatch):
    """Space-to-depth staging through the copy engine (stage_tma.cu: 4-D tensor-map boxes with out-of-bounds zero fill,
    bulk stores) == the frame's definition frame(Y', X')[(dy*2+dx)*4 + c] = in(c, 2Y'+dy-4, 2X'+dx-4), bit for bit, for
    contiguous and gathered rows; and == the load/store kernel it replaces."""
    from ips_b200 import ops, _lib
    x = _rand(P + 3, C, H, W, seed=7)
    xd = x.to(dev)
    assert _lib.load().ipsb_stage_tma_ok(xd.data_ptr(), C, H, W) == 1
    Ys, Wp = H // 2 + 3, W // 2 + 3
    xp = torch.zeros(P + 3, 4, 2 * Ys, 2 * Wp)
    xp[:, :C, 4:4 + H, 4:4 + W] = x                                  # in(c, y, x) at padded (y + 4, x + 4)
    # frame channel (dy*2+dx)*4 + c at (Y', X') = padded (2Y'+dy, 2X'+dx)
    ref = xp.view(P + 3, 4, Ys, 2, Wp, 2).permute(0, 2, 4, 3, 5, 1).reshape(P + 3, Ys * Wp, 16).to(torch.bfloat16)
    got = ops.stage_patches_s2d(xd, P, C, H, W, first_row=2).view(P, Ys * Wp, 16).cpu()
    assert torch.equal(got, ref[2:2 + P])
    idx = torch.randperm(P + 3, generator=torch.Generator().manual_seed(3))[:P]
    got = ops.stage_patches_s2d(xd, P, C, H, W, row_idx=idx.to(dev)).view(P, Ys * Wp, 16).cpu()
    assert torch.equal(got, ref[idx])
    out = torch.empty((P * Ys * Wp, 16), dtype=torch.bfloat16, device=dev)
    ops._call('ipsb_stage_patches_s2d_tma', xd.data_ptr(), 0, 1, P, C, H, W, out.data_ptr(), ops._stream())
    assert torch.equal(out.view(P, Ys * Wp, 16).cpu(), ref[1:1 + P])


@pytest.mark.parametrize('C,H,W', [(1, 50, 50), (3, 100, 100), (3, 8, 10)])
def test_stage_patches_padded(dev, C, H, W):
    from ips_b200 import ops
    x = _rand(7, C, H, W, seed=1)
    idx = torch.tensor([5, 0, 6, 3], dtype=torch.int64)
    ref = torch.zeros(7, H + 6, W + 6, 4)
    ref[:, 3:3 + H, 4:4 + W, :C] = x.permute(0, 2, 3, 1)
    ref = ref.to(torch.bfloat16)
    assert torch.equal(ops.stage_patches_padded(x.to(dev), 4, C, H, W, row_idx=idx.to(dev)).cpu(), ref[idx])
    assert torch.equal(ops.stage_patches_padded(x.to(dev), 3, C, H, W, first_row=2).cpu(), ref[2:5])


@pytest.mark.parametrize('H,W,P', [(50, 50, 37), (100, 100, 9), (20, 36, 5)])
def test_stem_tma(dev, H, W, P):
    """7x7/2 stem through the 5-D overlapping-window tensor map, against conv2d on the same bf16 data."""
    from ips_b200 import ops
    x = _rand(P, 3, H, W, seed=30)
    w = _rand(64, 3, 7, 7, seed=31, scale=math.sqrt(2.0 / 147)).to(torch.bfloat16)
    scale, shift = torch.rand(64) + 0.5, _rand(64, seed=32, scale=0.1)
    frame = ops.stage_patches_padded(x.to(dev), P, 3, H, W)
    wp = torch.zeros(64, 8, 8, 4)
    wp[:, :7, 1:8, :3] = w.float().permute(0, 2, 3, 1)
    got = ops.conv_bf16(frame, wp.reshape(64, 256).to(torch.bfloat16).to(dev), scale.to(dev), shift.to(dev), None,
                        64, 7, 7, 2, 3, True, 3).cpu().float()
    xb = x.to(torch.bfloat16).float()
    ref = torch.relu(F.conv2d(xb, w.float(), stride=2, padding=3) * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1))
    ref = ref.permute(0, 2, 3, 1)
    assert got.shape == ref.shape
    torch.testing.assert_close(got, ref, rtol=1e-2, atol=1e-2)


@pytest.mark.parametrize('row', [(1, 50, 50), (3, 100, 100), (2048,), (5,), (128,)])
def test_gather_rows(dev, row):
    from ips_b200 import ops
    B, N, M = 3, 37, 9
    src = _rand(B, N, *row, seed=2)
    g = torch.Generator().manual_seed(3)
    idx = torch.stack([torch.randperm(N, generator=g)[:M] for _ in range(B)])
    out = ops.gather_rows(src.to(dev), idx.to(dev), N)
    ref = torch.stack([src[b, idx[b]] for b in range(B)])
    assert torch.equal(out.cpu(), ref)
    tab = _rand(N, *row, seed=4)                                   # shared table (pos-enc gather)
    out = ops.gather_rows(tab.to(dev), idx.to(dev), 0)
    assert torch.equal(out.cpu(), tab[idx])


@pytest.mark.parametrize('H,W,C', [(25, 25, 64), (50, 50, 64), (7, 6, 8)])
def test_pools(dev, H, W, C):
    from ips_b200 import ops
    x = _rand(5, C, H, W, seed=5)
    nhwc = x.permute(0, 2, 3, 1).contiguous()
    ref = F.max_pool2d(x, 3, 2, 1).permute(0, 2, 3, 1)
    assert torch.equal(ops.maxpool3x3s2(nhwc.to(dev), ops.F32).cpu(), ref)
    xb = nhwc.to(torch.bfloat16)
    refb = F.max_pool2d(xb.float().permute(0, 3, 1, 2), 3, 2, 1).permute(0, 2, 3, 1).to(torch.bfloat16)
    assert torch.equal(ops.maxpool3x3s2(xb.to(dev), ops.BF16).cpu(), refb)
    ref = F.adaptive_avg_pool2d(x, 1).flatten(1)
    torch.testing.assert_close(ops.avgpool(nhwc.to(dev), ops.F32).cpu(), ref, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(ops.avgpool(xb.to(dev), ops.BF16).cpu(),
                               xb.float().mean(dim=(1, 2)), rtol=1e-5, atol=1e-6)


def test_layernorm_rows(dev):
    from ips_b200 import ops
    x = _rand(33, 2048, seed=6) * 3 + 0.7
    ref = F.layer_norm(x, (2048,), eps=1e-5)
    torch.testing.assert_close(ops.layernorm_rows(x.to(dev), 1e-5).cpu(), ref, rtol=1e-5, atol=1e-5)
    got = ops.rows_to_bf16(x.to(dev), True, 1e-5).cpu().float()
    torch.testing.assert_close(got, ref, rtol=8e-3, atol=8e-3)
    assert torch.equal(ops.rows_to_bf16(x.to(dev), False).cpu(), x.to(torch.bfloat16))


# ------------------------------------------------------------------ fp32 encoder layers

CONV_CASES = [  # Cin, Cout, k, stride, pad, H, W
    (4, 64, 7, 2, 3, 50, 50),       # stem on channel-padded input
    (64, 64, 3, 1, 1, 13, 13),
    (64, 128, 3, 2, 1, 13, 13),
    (64, 128, 1, 2, 0, 13, 13),
    (128, 128, 3, 1, 1, 7, 7),
    (256, 512, 3, 2, 1, 7, 7),
]


def _conv_ref(x_nhwc, w, scale, shift, res, stride, pad, relu):
    y = F.conv2d(x_nhwc.permute(0, 3, 1, 2), w, stride=stride, padding=pad)
    y = y * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)
    y = y.permute(0, 2, 3, 1)
    if res is not None:
        y = y + res
    return torch.relu(y) if relu else y


@pytest.mark.parametrize('case', CONV_CASES)
def test_conv_f32(dev, case):
    from ips_b200 import ops
    Cin, Cout, k, s, p, H, W = case
    P = 5
    x = _rand(P, H, W, Cin, seed=7)
    w = _rand(Cout, Cin, k, k, seed=8, scale=math.sqrt(2.0 / (Cin * k * k)))
    scale = torch.rand(Cout) + 0.5
    shift = _rand(Cout, seed=9, scale=0.1)
    Ho, Wo = (H + 2 * p - k) // s + 1, (W + 2 * p - k) // s + 1
    res = _rand(P, Ho, Wo, Cout, seed=10)
    w_kc = w.permute(2, 3, 1, 0).reshape(k * k * Cin, Cout).contiguous()
    for use_res, relu in ((False, True), (True, True), (False, False)):
        ref = _conv_ref(x, w, scale, shift, res if use_res else None, s, p, relu)
        got = ops.conv_f32(x.to(dev), w_kc.to(dev), scale.to(dev), shift.to(dev), res.to(dev) if use_res else None,
                           Cout, k, k, s, p, relu).cpu()
        torch.testing.assert_close(got, ref, rtol=2e-5, atol=2e-5)


@pytest.mark.parametrize('M,N,K', [(70, 512, 2048), (5, 10, 128), (300, 64, 512), (129, 130, 132)])
def test_linear_f32(dev, M, N, K):
    from ips_b200 import ops
    a, w = _rand(M, K, seed=11), _rand(N, K, seed=12, scale=1 / math.sqrt(K))
    scale, shift = torch.rand(N) + 0.5, _rand(N, seed=13)
    ref = torch.relu((a @ w.t()) * scale + shift)
    got = ops.linear_f32(a.to(dev), w.to(dev), scale.to(dev), shift.to(dev), relu=True).cpu()
    torch.testing.assert_close(got, ref, rtol=2e-5, atol=2e-5)
    got = ops.linear_f32(a.to(dev), w.to(dev)).cpu()
    torch.testing.assert_close(got, a @ w.t(), rtol=2e-5, atol=2e-5)


# ------------------------------------------------------------------ tcgen05 layers (bf16 in, fp32 accumulate)

@pytest.mark.parametrize('case', CONV_CASES + [(64, 64, 3, 1, 1, 25, 25), (512, 512, 3, 1, 1, 4, 4), (128, 256, 3, 2, 1, 13, 13),
                                               (64, 64, 3, 1, 1, 9, 14)])
@pytest.mark.parametrize('path', ['tma', 'gather'])
def test_conv_bf16_umma(dev, case, path):
    from ips_b200 import ops
    Cin, Cout, k, s, p, H, W = case
    if Cin == 4 and path == 'gather':
        pytest.skip('the stem has a single path')
    P = 37                                                       # several M tiles + a ragged tail
    x = _rand(P, H, W, Cin, seed=14).to(torch.bfloat16)
    w = _rand(Cout, Cin, k, k, seed=15, scale=math.sqrt(2.0 / (Cin * k * k))).to(torch.bfloat16)
    scale = torch.rand(Cout) + 0.5
    shift = _rand(Cout, seed=16, scale=0.1)
    Ho, Wo = (H + 2 * p - k) // s + 1, (W + 2 * p - k) // s + 1
    res = _rand(P, Ho, Wo, Cout, seed=17).to(torch.bfloat16)
    if Cin == 4:
        wp = torch.zeros(Cout, 8, 8, 4)
        wp[:, :k, 1:k + 1] = w.float().permute(0, 2, 3, 1)       # taps shifted by one column
        w_nk, mode = wp.reshape(Cout, 256).to(torch.bfloat16), 1
    else:
        w_nk, mode = w.permute(0, 2, 3, 1).reshape(Cout, k * k * Cin).contiguous(), (0 if path == 'tma' else 2)
    for use_res, relu in ((False, True), (True, True), (False, False)):
        ref = _conv_ref(x.float(), w.float(), scale, shift, res.float() if use_res else None, s, p, relu)
        got = ops.conv_bf16(x.to(dev), w_nk.to(dev), scale.to(dev), shift.to(dev), res.to(dev) if use_res else None,
                            Cout, k, k, s, p, relu, mode).cpu().float()
        # exact products, fp32 accumulation; only the bf16 output rounding (2^-8 relative) differs
        torch.testing.assert_close(got, ref, rtol=1e-2, atol=1e-2)
        assert (got - ref).abs().mean() < 4e-3 * ref.abs().mean() + 1e-4


PF_CASES = [  # Cin, Cout, k, stride, pad, H, W, in_pf, out_pf
    (64, 64, 3, 1, 1, 25, 25, True, True),       # shifted-window kernel, resident weights
    (64, 64, 3, 1, 1, 13, 13, True, True),
    (128, 128, 3, 1, 1, 13, 13, True, True),     # shifted-window kernel, streamed weights
    (128, 128, 3, 1, 1, 7, 7, True, True),
    (64, 64, 3, 1, 1, 6, 11, True, True),
    (64, 128, 3, 2, 1, 25, 25, True, True),      # TMA box kernel reading / writing PF
    (64, 128, 1, 2, 0, 25, 25, True, True),
    (128, 256, 3, 2, 1, 13, 13, True, True),
    (256, 256, 3, 1, 1, 7, 7, True, True),
    (512, 512, 3, 1, 1, 4, 4, True, False),
    (64, 64, 3, 1, 1, 13, 13, False, True),
]


@pytest.mark.parametrize('case', PF_CASES)
def test_conv_bf16_pf(dev, case):
    """Padded-flat activations: shifted-window 3x3 kernel and the PF in/out modes of the box kernel."""
    from ips_b200 import ops
    Cin, Cout, k, s, p, H, W, in_pf, out_pf = case
    P = 41
    x = _rand(P, H, W, Cin, seed=40).to(torch.bfloat16)
    w = _rand(Cout, Cin, k, k, seed=41, scale=math.sqrt(2.0 / (Cin * k * k))).to(torch.bfloat16)
    scale, shift = torch.rand(Cout) + 0.5, _rand(Cout, seed=42, scale=0.1)
    Ho, Wo = (H + 2 * p - k) // s + 1, (W + 2 * p - k) // s + 1
    res = _rand(P, Ho, Wo, Cout, seed=43).to(torch.bfloat16)
    w_nk = w.permute(0, 2, 3, 1).reshape(Cout, k * k * Cin).contiguous().to(dev)
    xd = x.to(dev)
    xin = ops.to_pf(xd) if in_pf else xd
    for use_res, relu in ((False, True), (True, True), (False, False)):
        ref = _conv_ref(x.float(), w.float(), scale, shift, res.float() if use_res else None, s, p, relu)
        r = None
        if use_res:
            r = ops.to_pf(res.to(dev)) if out_pf else res.to(dev)
        y = ops.conv_bf16_pf(xin, w_nk, scale.to(dev), shift.to(dev), r, P, H, W, Cout, k, k, s, p, relu, in_pf, out_pf)
        if out_pf:
            got = ops.from_pf(y, P, Ho, Wo).cpu().float()
            # pad rows must still be zero: the next layer relies on it
            chk = y.clone()
            rows, G0, Wp, Sp = ops.pf_geo(P, Ho, Wo)
            chk[G0:G0 + P * Sp].view(P, Ho + 1, Wp, Cout)[:, :Ho, :Wo] = 0
            assert float(chk.abs().max()) == 0.0
        else:
            got = y.cpu().float()
        torch.testing.assert_close(got, ref, rtol=1e-2, atol=1e-2)


def test_pools_pf(dev):
    from ips_b200 import ops
    x = _rand(9, 64, 50, 50, seed=44).to(torch.bfloat16)
    nhwc = x.permute(0, 2, 3, 1).contiguous().to(dev)
    y = ops.maxpool3x3s2_pf(nhwc)
    ref = F.max_pool2d(x.float(), 3, 2, 1).permute(0, 2, 3, 1).to(torch.bfloat16)
    assert torch.equal(ops.from_pf(y, 9, 25, 25).cpu(), ref)
    assert torch.equal(ops.to_pf(ref.to(dev)), y)                  # pads zero, same layout as the helper
    a = ops.avgpool_pf(y, 9, 25, 25).cpu()
    torch.testing.assert_close(a, ref.float().mean(dim=(1, 2)), rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize('M,N,K', [(128, 64, 64), (70, 512, 2048), (1000, 128, 512), (257, 192, 128)])
def test_linear_bf16_umma(dev, M, N, K):
    from ips_b200 import ops
    a = _rand(M, K, seed=18).to(torch.bfloat16)
    w = _rand(N, K, seed=19, scale=1 / math.sqrt(K)).to(torch.bfloat16)
    scale, shift = torch.rand(N) + 0.5, _rand(N, seed=20)
    ref = torch.relu((a.float() @ w.float().t()) * scale + shift)
    got = ops.linear_bf16(a.to(dev), w.to(dev), scale.to(dev), shift.to(dev), relu=True).cpu()
    torch.testing.assert_close(got, ref, rtol=1e-4, atol=1e-4)    # fp32 output: accumulation order only


# ------------------------------------------------------------------ scoring

@pytest.mark.parametrize('pre', ['mnist', 'traffic', 'camelyon'])
def test_logits_and_scores(dev, pre):
    """P2: scores vs Transformer.get_scores of the reference (through the oracle)."""
    from ips_b200 import IPSNet, Struct
    conf = O.preset(pre, N=64) if pre != 'camelyon' else O.preset(pre)
    sd = O.make_state(conf, 21, q_gain=12.0)
    net = IPSNet(dev, Struct(**conf.__dict__)).to(dev)
    net.load_state_dict(sd)
    B, L = 3, 77
    emb = _rand(B, L, conf.D, seed=22)
    ref_logits = O.attn_logits(sd, conf, emb)                     # (B,H,T,L)
    got_logits = net.transf.crs_attn.get_logits(emb.to(dev)).cpu().view(B, L, conf.H, conf.n_token).permute(0, 2, 3, 1)
    torch.testing.assert_close(got_logits, ref_logits, rtol=1e-4, atol=2e-5)
    ref = O.attn_scores(sd, conf, emb)
    got = net.transf.get_scores(emb.to(dev)).cpu()
    torch.testing.assert_close(got, ref, rtol=2e-4, atol=1e-7)
    torch.testing.assert_close(got.sum(-1), torch.ones(B), rtol=0, atol=1e-5)


@pytest.mark.parametrize('H,T,L', [(8, 1, 42), (8, 4, 200), (8, 1, 10000), (3, 2, 130), (1, 1, 33)])
def test_scores_from_logits(dev, H, T, L):
    from ips_b200 import ops
    z = _rand(2, L, H * T, seed=23, scale=2.0)
    a = torch.softmax(z.view(2, L, H, T).permute(0, 2, 3, 1), dim=-1)      # (B,H,T,L)
    ref = a.mean(dim=1).transpose(1, 2).mean(-1)
    got = ops.scores_from_logits(z.to(dev), H, T).cpu()
    torch.testing.assert_close(got, ref, rtol=1e-5, atol=1e-9)


@pytest.mark.parametrize('L,M', [(42, 10), (200, 100), (10000, 5000), (16384, 5000), (7, 7), (1000, 1)])
def test_topm_stable(dev, L, M):
    """P1: exactly the first M of a stable descending sort, including heavy ties."""
    from ips_b200 import ops
    g = torch.Generator().manual_seed(24)
    s = torch.rand(3, L, generator=g)
    s[1] = (s[1] * 8).floor() / 8                                # 8 distinct values -> many ties
    s[2] = 0.25                                                  # all tied -> positions 0..M-1
    val, idx = ops.topm_stable(s.to(dev), M)
    rv, ri = torch.sort(s, dim=-1, descending=True, stable=True)
    assert torch.equal(idx.cpu(), ri[:, :M])
    assert torch.equal(val.cpu(), rv[:, :M])


def _loop_oracle(sd, conf, emb, perm, M, I):
    """The reference loop (ips_net.py:213-241) composed from the oracle's own `score_and_select` on embeddings, exactly
    as `O.ips` iterates it: cat(memory, next I) -> score -> top-M.  Returns (positions in scan order, original indices),
    best first, plus the final iteration's sorted scores (for the boundary report)."""
    B, N, D = emb.shape
    order = perm if perm is not None else torch.arange(N).unsqueeze(0).expand(B, -1)
    px = torch.gather(emb, 1, order.unsqueeze(-1).expand(-1, -1, D))
    idx = torch.arange(N).unsqueeze(0).expand(B, -1)
    mem_emb, mem_idx = px[:, :M], idx[:, :M]
    s = None
    for it in range(math.ceil((N - M) / I)):
        lo = M + it * I
        hi = min(lo + I, N)
        all_emb = torch.cat((mem_emb, px[:, lo:hi]), dim=1)
        all_idx = torch.cat((mem_idx, idx[:, lo:hi]), dim=1)
        mem_emb, mem_idx, s = O.score_and_select(sd, conf, all_emb, None, M, all_idx, tie='stable')
    return mem_idx, torch.gather(order, 1, mem_idx), s.sort(-1, descending=True)[0]


@pytest.mark.parametrize('N,M,I,H,T', [(192, 10, 32, 8, 1), (900, 100, 100, 8, 4), (3000, 500, 500, 8, 1),
                                       (50, 49, 7, 2, 2), (23000, 5000, 5000, 8, 1), (9000, 1500, 1500, 8, 1),
                                       (12345, 4000, 777, 8, 1),
                                       (45000, 5000, 35000, 8, 1),       # 40 000-entry buffer: cluster of 16 CTAs
                                       (30000, 5000, 5000, 8, 4)])       # H*T = 32 on a long buffer: cluster of 16 CTAs
@pytest.mark.parametrize('shuffle', ['none', 'batch', 'instance'])
def test_select_loop(dev, N, M, I, H, T, shuffle):
    """The in-kernel loop on the logit table == the oracle's `score_and_select` iterated on the embeddings
    (ips_net.py:136-155,218-241): same winners in the same order; a differing pick is accepted only when the oracle's
    own boundary gap is within fp32 rounding (P3)."""
    from ips_b200 import ops
    B, D, Dk = 2, 64, 8
    conf = O.preset('camelyon', M=M, I=I, H=H, n_token=T, D=D, D_k=Dk, D_v=Dk, D_inner=64, n_chan_in=16, shuffle=False)
    if T > 1:
        conf.tasks = {f'task{t}': {'id': t, 'name': f't{t}', 'act_fn': 'softmax', 'metric': 'accuracy'} for t in range(T)}
    sd = O.make_state(conf, 27, q_gain=6.0)
    emb = _rand(B, N, D, seed=25, scale=1.0)
    z = O.attn_logits(sd, conf, emb).permute(0, 3, 1, 2).reshape(B, N, H * T).contiguous()       # (B,N,HT): h-major, t-minor
    g = torch.Generator().manual_seed(26)
    perm, per_inst = None, False
    if shuffle == 'batch':
        perm = torch.randperm(N, generator=g).unsqueeze(0)
    elif shuffle == 'instance':
        perm, per_inst = torch.stack([torch.randperm(N, generator=g) for _ in range(B)]), True
    pos, src, score = ops.select_loop(z.to(dev), None if perm is None else perm.to(dev), per_inst, H, T, M, I)
    ref_pos, ref_src, ref_sorted = _loop_oracle(sd, conf, emb, None if perm is None else perm.expand(B, -1), M, I)
    mism = (src.cpu() != ref_src).sum().item()
    if mism:
        for b in range(B):
            a, c = set(src[b].tolist()), set(ref_src[b].tolist())
            if a != c:                                          # a different SET: only at an fp32-rounding-sized boundary gap
                gap = float((ref_sorted[b, M - 1] - ref_sorted[b, M]) / ref_sorted[b, M - 1]) if ref_sorted.shape[1] > M else 1.0
                assert gap < 1e-5, f'image {b}: {len(a ^ c) // 2} different picks at boundary gap {gap:.2e}'
    else:
        assert torch.equal(pos.cpu(), ref_pos)
    sc = score.cpu()
    assert (sc[:, :-1] >= sc[:, 1:]).all()                        # best first


@pytest.mark.parametrize('B,L,M', [(2, 40000, 5000), (1, 10000, 5000), (3, 20000, 5000), (2, 300, 100)])
def test_merge_candidates(dev, B, L, M):
    """The candidate merge of the sequence-sharded schedule (global re-score + stable top-M of R*M candidates): one
    iteration of the cluster loop on long lists == scores_from_logits + topm_stable (order swaps only between
    candidates whose scores agree to fp32 rounding)."""
    from ips_b200 import ops
    H, T = 8, 1
    cz = _rand(B, L, H * T, seed=29, scale=1.5).to(dev)
    got = ops.merge_candidates(cz, H, T, M)
    sc = ops.scores_from_logits(cz, H, T)
    ref = ops.topm_stable(sc, M)[1]
    assert torch.equal(got.sort(-1)[0], ref.sort(-1)[0])
    a, b = torch.gather(sc, 1, got), torch.gather(sc, 1, ref)
    assert float(((a - b).abs() / b).max()) <= 2e-6
    assert bool((a[:, :-1] >= a[:, 1:] * (1 - 2e-6)).all())


@pytest.mark.parametrize('L,M', [(20000, 5000), (40000, 5000), (16385, 1), (33000, 8192)])
def test_topm_stable_long_rows(dev, L, M):
    """Rows longer than the single-CTA sort (the candidate merge of the sequence-sharded schedule, L = ranks * M):
    radix select + compaction + sort of the winners == the first M of a stable descending sort, heavy ties included."""
    from ips_b200 import ops
    g = torch.Generator().manual_seed(28)
    s = torch.rand(3, L, generator=g)
    s[1] = (s[1] * 8).floor() / 8
    s[2] = 0.25
    val, idx = ops.topm_stable(s.to(dev), M)
    rv, ri = torch.sort(s, dim=-1, descending=True, stable=True)
    assert torch.equal(idx.cpu(), ri[:, :M])
    assert torch.equal(val.cpu(), rv[:, :M])


@pytest.mark.parametrize('rows,K,N,HT,in_bf16', [(333, 2048, 512, 8, False), (128, 256, 512, 8, False), (5000, 2048, 512, 8, True),
                                                  (1000, 512, 256, 4, False), (40000, 2048, 512, 8, False), (129, 1024, 512, 1, True)])
def test_projector_logits_fused(dev, rows, K, N, HT, in_bf16):
    """One kernel from features to logits (csrc/umma_projector.cu) against the fp32 definition of the same chain
    (ips_net.py:54-60 + transformer.py:71-83) evaluated on the bf16-rounded operands: LayerNorm applied algebraically in
    the epilogue, BatchNorm + ReLU, score projection; ragged last tile, fp32 and bf16 features, 256- and 512-wide."""
    from ips_b200 import ops
    x = _rand(rows, K, seed=60) * 1.5 + 0.3
    w = _rand(N, K, seed=61) * (1.0 / math.sqrt(K))
    scale, shift = torch.rand(N, generator=torch.Generator().manual_seed(62)) + 0.5, _rand(N, seed=63) * 0.2
    U = _rand(N, HT, seed=64) * 0.1
    wb = w.to(torch.bfloat16)
    xin = x.to(torch.bfloat16) if in_bf16 else x
    tab = ops.projector_table(scale.to(dev), shift.to(dev), wb.to(dev), U.to(dev))
    z = ops.projector_logits(xin.to(dev), wb.to(dev), tab, HT, 1e-5).cpu()
    xr = x.to(torch.bfloat16).double()                                  # what the tensor cores see
    mean, var = xr.mean(1, keepdim=True), xr.var(1, unbiased=False, keepdim=True)
    y = ((xr - mean) / torch.sqrt(var + 1e-5)) @ wb.double().t()
    emb = torch.relu(y * scale.double() + shift.double())
    ref = (emb @ U.double()).float()
    err = (z - ref).abs().max().item() / ref.abs().max().item()
    assert err < 2e-5, err                                              # fp32 accumulation of exact bf16 products
    # and against the unfused chain of round 1 (LayerNorm -> bf16 -> GEMM -> logits): same result within bf16 rounding of LN(x)
    a = ops.rows_to_bf16(xin.to(dev), layernorm=True, eps=1e-5)
    emb2 = ops.linear_bf16(a, wb.to(dev), scale.to(dev), shift.to(dev), relu=True)
    z2 = ops.logits(emb2, U.to(dev).contiguous()).cpu()
    assert (z - z2).abs().max().item() <= 2e-2 * ref.abs().max().item()


@pytest.mark.parametrize('B,N,M,I,shuffle,overlap', [(1, 50000, 5000, 5000, 'batch', True), (1, 50000, 5000, 5000, None, True),
                                                      (2, 9000, 1500, 1000, 'instance', True), (3, 7000, 1024, 1024, 'batch', None),
                                                      (1, 20011, 2000, 3000, 'batch', False), (1, 2500, 2000, 3000, 'batch', True)])
def test_projector_select_streamed(dev, B, N, M, I, shuffle, overlap):
    """Streamed selection (ops.projector_select: projector kernel reading the rows through the scan order + the cluster
    loop consuming its tiles through flags, on two streams) == the sequential composition (projector in original order ->
    ipsb_select_loop with the same perm): same logits bit for bit (scan order), same winners in the same order; shared and
    per-instance scan orders, ragged last tile / last chunk, one chunk only, back-to-back fallback for more than two bags."""
    from ips_b200 import ops
    K, Nw, H, T = 512, 512, 8, 1
    HT = H * T
    g = torch.Generator().manual_seed(70 + N)
    x = (torch.randn(B * N, K, generator=g) * 1.5 + 0.3).to(dev)
    wb = (torch.randn(Nw, K, generator=g) / math.sqrt(K)).to(torch.bfloat16).to(dev)
    scale, shift = (torch.rand(Nw, generator=g) + 0.5).to(dev), (torch.randn(Nw, generator=g) * 0.2).to(dev)
    U = (torch.randn(Nw, HT, generator=g) * 0.5).to(dev)
    tab = ops.projector_table(scale, shift, wb, U)
    perm, per_inst = None, False
    if shuffle == 'batch':
        perm = torch.randperm(N, generator=g).unsqueeze(0).to(dev)
    elif shuffle == 'instance':
        perm, per_inst = torch.rand((B, N), generator=g).argsort(1).to(dev).contiguous(), True
    assert ops.streamed_select_ok(B, N, HT, M, I)
    z = ops.projector_logits(x, wb, tab, HT, 1e-5).view(B, N, HT)
    pos0, src0, sc0 = ops.select_loop(z, perm, per_inst, H, T, M, I)
    for rep in range(3):                                                # repeated: flags / sync words are per call
        pos1, src1, sc1, zs, status = ops.projector_select(x, wb, tab, B, N, perm, per_inst, H, T, M, I, overlap=overlap)
        torch.cuda.synchronize()
        if status is not None:
            assert status.tolist() == [1, 0]
        idx = (torch.arange(N, device=dev).expand(B, N) if perm is None else perm.expand(B, N))
        assert torch.equal(zs, torch.gather(z, 1, idx.unsqueeze(-1).expand(B, N, HT)))
        assert torch.equal(pos0, pos1) and torch.equal(src0, src1) and torch.equal(sc0, sc1)


@pytest.mark.parametrize('rows,N', [(1, 1), (1, 2), (3, 3), (2, 1000), (1, 50000), (4, 65536), (1, 200001)])
def test_keyed_scan_order_matches_oracle(dev, rows, N):
    """Device-side scan order (ipsb_keyed_scan_order) == its CPU restatement, element for element, and every row is a
    permutation of [0, N)."""
    from ips_b200 import ops
    key = torch.tensor([-0x1234567890abcdef + N, 0x0fedcba987654321 - rows], dtype=torch.int64)
    got = ops.keyed_scan_order(key.to(dev), rows, N).cpu()
    ref = O.keyed_scan_order(int(key[0]), int(key[1]), rows, N)
    assert torch.equal(got, ref)
    assert torch.equal(got.sort(1).values, torch.arange(N).expand(rows, N))
    if rows > 1 and N > 100:
        assert not torch.equal(got[0], got[1])


def test_streamed_select_in_ips_and_graph(dev):
    """IPSNet.ips on a feature bag takes the streamed path and returns what the sequential path returns; the two-stream
    call is capturable as one CUDA graph (fork / join) and the replay gives the same winners."""
    from ips_b200 import IPSNet, Struct
    conf = O.preset('camelyon', M=2000, I=2000, shuffle=True)
    net = IPSNet(dev, Struct(**conf.__dict__)).to(dev)
    net.load_state_dict(O.make_state(conf, 9, q_gain=8.0))
    net.eval()
    x = O.make_patches(conf, 1, 12000, 31).to(dev)
    torch.manual_seed(5)
    a, _ = net.ips(x)
    idx_a = net.last_mem_idx.clone()
    assert net.last_stream_status is not None and net.last_stream_status.tolist() == [1, 0]
    net.streamed_select = False
    torch.manual_seed(5)
    b, _ = net.ips(x)
    assert torch.equal(idx_a, net.last_mem_idx) and torch.equal(a, b)
    net.streamed_select = True
    net.shuffle = False                                                  # (the host-side randperm cannot be captured)
    ref, _ = net.ips(x)
    idx_ref = net.last_mem_idx.clone()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            out, _ = net.ips(x)
            idx_g = net.last_mem_idx
    torch.cuda.current_stream().wait_stream(side)
    for _ in range(3):
        out.zero_()
        g.replay()
        torch.cuda.synchronize()
        assert torch.equal(out, ref) and torch.equal(idx_g, idx_ref)
    # scan order drawn on the device: part of the graph; a replay after manual_seed == the eager call on the same seed
    net.shuffle, net.scan_order_rng = True, 'device'
    torch.manual_seed(77)
    ref, _ = net.ips(x)
    idx_ref = net.last_mem_idx.clone()
    torch.cuda.synchronize()
    g2 = torch.cuda.CUDAGraph()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        with torch.cuda.graph(g2, stream=side):
            out2, _ = net.ips(x)
            idx_g2 = net.last_mem_idx
    torch.cuda.current_stream().wait_stream(side)
    torch.manual_seed(77)
    g2.replay()
    torch.cuda.synchronize()
    assert torch.equal(out2, ref) and torch.equal(idx_g2, idx_ref)
    g2.replay()                                                           # next draw: another order, still a valid selection
    torch.cuda.synchronize()
    assert torch.equal(out2, x[0, idx_g2[0]].unsqueeze(0))


@pytest.mark.parametrize('pre,precision,stem_env', [('traffic', 'bf16', None), ('mnist', 'bf16', None), ('traffic', 'fp32', None),
                                                    ('mnist', 'bf16', 'tma')])
def test_fold_plan_one_launch(dev, pre, precision, stem_env, monkeypatch):
    """ipsb_fold_plan (every conv weight re-laid out + BatchNorm(eval) folded, ONE launch) against the per-layer torch
    formulas it replaces: K-major / (K, N) layouts, the space-to-depth and 8x8 stem packings, scale / shift."""
    from ips_b200 import IPSNet, Struct
    if stem_env:
        monkeypatch.setenv('IPS_B200_STEM', stem_env)
    conf = O.preset(pre, precision=precision)
    net = IPSNet(dev, Struct(**conf.__dict__)).to(dev)
    net.load_state_dict(O.make_state(conf, 5, q_gain=3.0))
    plan = net._build_plan()

    def expect(conv, bn, e, stem):
        w = conv.weight.detach().float()
        cout, cin, kh, kw = w.shape
        scale = bn.weight.detach() * torch.rsqrt(bn.running_var.detach() + bn.eps)
        shift = bn.bias.detach() - bn.running_mean.detach() * scale
        if stem:
            w4 = torch.zeros((cout, 4, kh, kw), device=w.device)
            w4[:, :cin] = w
            if precision == 'bf16' and e['mode'] == 4:
                wp = torch.zeros((cout, 8, 8, 4), device=w.device)
                wp[:, 1:kh + 1, 1:kw + 1] = w4.permute(0, 2, 3, 1)
                ww = wp.view(cout, 4, 2, 4, 2, 4).permute(0, 1, 3, 2, 4, 5).reshape(cout, 256).to(torch.bfloat16)
            elif precision == 'bf16':
                wp = torch.zeros((cout, 8, 8, 4), device=w.device)
                wp[:, :kh, 1:kw + 1] = w4.permute(0, 2, 3, 1)
                ww = wp.reshape(cout, 256).to(torch.bfloat16)
            else:
                ww = w4.permute(2, 3, 1, 0).reshape(kh * kw * 4, cout)
        elif precision == 'bf16':
            ww = w.permute(0, 2, 3, 1).reshape(cout, kh * kw * cin).to(torch.bfloat16)
        else:
            ww = w.permute(2, 3, 1, 0).reshape(kh * kw * cin, cout)
        assert e['w'].shape == ww.shape and e['w'].dtype == ww.dtype
        assert torch.equal(e['w'], ww.contiguous())
        torch.testing.assert_close(e['scale'], scale, rtol=2e-6, atol=0)
        torch.testing.assert_close(e['shift'], shift, rtol=1e-5, atol=1e-7)

    enc = net.encoder
    expect(enc[0], enc[1], plan['stem'], True)
    blocks = [blk for child in list(enc.children())[4:-1] for blk in child]
    assert len(blocks) == len(plan['blocks'])
    for blk, b in zip(blocks, plan['blocks']):
        expect(blk.conv1, blk.bn1, b['c1'], False)
        expect(blk.conv2, blk.bn2, b['c2'], False)
        assert (blk.downsample is None) == (b['ds'] is None)
        if blk.downsample is not None:
            expect(blk.downsample[0], blk.downsample[1], b['ds'], False)
    # a second build after an in-place update refills the SAME buffers
    ptr = plan['stem']['w'].data_ptr()
    with torch.no_grad():
        enc[0].weight.mul_(2.0)
    plan2 = net._get_plan()
    assert plan2['stem']['w'].data_ptr() == ptr
    expect(enc[0], enc[1], plan2['stem'], True)


# ------------------------------------------------------------------ aggregator + heads (no-grad forward)

@pytest.mark.parametrize('B,M,H,Dk,Dv,T', [(2, 100, 8, 16, 16, 4), (3, 10, 8, 64, 64, 1), (1, 5000, 8, 64, 64, 1), (2, 7, 3, 8, 24, 2)])
def test_cross_attention(dev, B, M, H, Dk, Dv, T):
    from ips_b200 import ops
    q = _rand(T, H * Dk, seed=50) / math.sqrt(Dk)
    k, v = _rand(B, M, H * Dk, seed=51), _rand(B, M, H * Dv, seed=52)
    qq = q.view(1, T, H, Dk).transpose(1, 2)
    kk = k.view(B, M, H, Dk).transpose(1, 2)
    vv = v.view(B, M, H, Dv).transpose(1, 2)
    ref = torch.matmul(torch.softmax(torch.matmul(qq, kk.transpose(2, 3)), -1), vv).transpose(1, 2).reshape(B, T, H * Dv)
    got = ops.cross_attention(q.to(dev), k.to(dev), v.to(dev), H, Dk, Dv).cpu()
    torch.testing.assert_close(got, ref, rtol=1e-4, atol=1e-5)


def test_residual_layernorm_and_heads(dev):
    from ips_b200 import ops
    x, r = _rand(12, 128, seed=53), _rand(4, 128, seed=54)
    g, b = torch.rand(128) + 0.5, _rand(128, seed=55)
    ref = F.layer_norm(x + r.repeat(3, 1), (128,), g, b, 1e-6)
    torch.testing.assert_close(ops.residual_layernorm(x.to(dev), r.to(dev), g.to(dev), b.to(dev), 1e-6).cpu(), ref, rtol=1e-5, atol=1e-5)
    ref = F.layer_norm(x, (128,), g, b, 1e-6)
    torch.testing.assert_close(ops.residual_layernorm(x.to(dev), None, g.to(dev), b.to(dev), 1e-6).cpu(), ref, rtol=1e-5, atol=1e-5)
    z = _rand(9, 10, seed=56) * 3
    torch.testing.assert_close(ops.head_activation(z.to(dev), 'softmax').cpu(), torch.softmax(z, -1), rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(ops.head_activation(z.to(dev), 'sigmoid').cpu(), torch.sigmoid(z), rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(ops.add(x.to(dev), x.to(dev)).cpu(), x + x)


# ------------------------------------------------------------------ general tensor-core GEMM (train step)

@pytest.mark.parametrize('mode', ['nt', 'nn', 'tn'])
@pytest.mark.parametrize('M,N,K', [(128, 64, 64), (512, 2048, 8000), (1000, 512, 2048), (256, 128, 130000), (384, 192, 72)])
def test_gemm_bf16_modes(dev, mode, M, N, K):
    """NT / NN / TN products against fp32 matmul of the same bf16 operands (MN-major operands, split-K)."""
    from ips_b200 import ops
    if mode == 'tn' and M % 128:
        pytest.skip('TN split-K needs M % 128 == 0')
    if mode != 'tn' and K > 10000:
        K = 4096
    g = torch.Generator().manual_seed(60)
    x = (torch.randn(M, K, generator=g) / math.sqrt(K)).to(torch.bfloat16)
    w = torch.randn(N, K, generator=g).to(torch.bfloat16)
    ref = x.float() @ w.float().t()
    if mode == 'nt':
        got = ops.gemm_bf16('nt', x.to(dev), w.to(dev))
    elif mode == 'nn':
        got = ops.gemm_bf16('nn', x.to(dev), w.t().contiguous().to(dev))
    else:
        got = ops.gemm_bf16('tn', x.t().contiguous().to(dev), w.t().contiguous().to(dev))
    torch.testing.assert_close(got.cpu(), ref, rtol=2e-3, atol=2e-3)
    if mode == 'nt':
        shift = torch.randn(N, generator=g)
        got = ops.gemm_bf16('nt', x.to(dev), w.to(dev), shift=shift.to(dev), relu=True, out_dtype=torch.bfloat16)
        torch.testing.assert_close(got.cpu().float(), torch.relu(ref + shift), rtol=2e-2, atol=2e-2)


# ------------------------------------------------------------------ grad-mode operators vs torch autograd

def test_batchnorm_train_fn(dev):
    from ips_b200.autograd import BatchNormTrainFn
    x = _rand(300, 70, seed=70) * 2 + 0.5
    g, b = torch.rand(70) + 0.5, _rand(70, seed=71)
    dy = _rand(300, 70, seed=72)
    bn = torch.nn.BatchNorm1d(70)
    with torch.no_grad():
        bn.weight.copy_(g); bn.bias.copy_(b)
    xr = x.clone().requires_grad_(True)
    yr = torch.relu(bn(xr)); yr.backward(dy)
    xg = x.to(dev).requires_grad_(True)
    gg, bg = g.to(dev).requires_grad_(True), b.to(dev).requires_grad_(True)
    rm, rv = torch.zeros(70, device=dev), torch.ones(70, device=dev)
    y = BatchNormTrainFn.apply(xg, gg, bg, rm, rv, 0.1, 1e-5, True)
    y.backward(dy.to(dev))
    torch.testing.assert_close(y.detach().cpu(), yr.detach(), rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(xg.grad.cpu(), xr.grad, rtol=1e-3, atol=1e-5)
    torch.testing.assert_close(gg.grad.cpu(), bn.weight.grad, rtol=1e-3, atol=1e-4)
    torch.testing.assert_close(bg.grad.cpu(), bn.bias.grad, rtol=1e-3, atol=1e-4)
    torch.testing.assert_close(rm.cpu(), bn.running_mean, rtol=1e-4, atol=1e-6)
    torch.testing.assert_close(rv.cpu(), bn.running_var, rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize('affine', [True, False])
def test_layernorm_fn(dev, affine):
    from ips_b200.autograd import LayerNormFn
    x, dy = _rand(2, 9, 128, seed=73) * 3, _rand(2, 9, 128, seed=74)
    g, b = (torch.rand(128) + 0.5, _rand(128, seed=75)) if affine else (None, None)
    xr = x.clone().requires_grad_(True)
    gr = g.clone().requires_grad_(True) if affine else None
    br = b.clone().requires_grad_(True) if affine else None
    F.layer_norm(xr, (128,), gr, br, 1e-6).backward(dy)
    xg = x.to(dev).requires_grad_(True)
    gg = g.to(dev).requires_grad_(True) if affine else None
    bg = b.to(dev).requires_grad_(True) if affine else None
    y = LayerNormFn.apply(xg, gg, bg, 1e-6)
    y.backward(dy.to(dev))
    torch.testing.assert_close(y.detach().cpu(), F.layer_norm(x, (128,), g, b, 1e-6), rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(xg.grad.cpu(), xr.grad, rtol=1e-3, atol=1e-5)
    if affine:
        torch.testing.assert_close(gg.grad.cpu(), gr.grad, rtol=1e-3, atol=1e-4)
        torch.testing.assert_close(bg.grad.cpu(), br.grad, rtol=1e-3, atol=1e-4)


@pytest.mark.parametrize('B,M,H,Dk,Dv,T,drop', [(2, 50, 8, 16, 16, 4, 0.0), (3, 10, 8, 64, 64, 1, 0.3), (1, 700, 8, 64, 64, 1, 0.1)])
def test_cross_attention_fn(dev, B, M, H, Dk, Dv, T, drop):
    from ips_b200.autograd import CrossAttentionFn
    q = _rand(T, H * Dk, seed=76) / math.sqrt(Dk)
    k, v = _rand(B, M, H * Dk, seed=77), _rand(B, M, H * Dv, seed=78)
    dout = _rand(B, T, H * Dv, seed=79)
    mask = (torch.rand(B, H, T, M) >= drop).float() if drop > 0 else None
    keep = 1.0 / (1.0 - drop)
    qr, kr, vr = q.clone().requires_grad_(True), k.clone().requires_grad_(True), v.clone().requires_grad_(True)
    a = torch.softmax(torch.matmul(qr.view(1, T, H, Dk).transpose(1, 2), kr.view(B, M, H, Dk).transpose(1, 2).transpose(2, 3)), -1)
    if mask is not None:
        a = a * mask * keep
    ref = torch.matmul(a, vr.view(B, M, H, Dv).transpose(1, 2)).transpose(1, 2).reshape(B, T, H * Dv)
    ref.backward(dout)
    qg, kg, vg = (t.to(dev).requires_grad_(True) for t in (q, k, v))
    out = CrossAttentionFn.apply(qg, kg, vg, None if mask is None else mask.to(dev), keep, H, Dk, Dv)
    out.backward(dout.to(dev))
    torch.testing.assert_close(out.detach().cpu(), ref.detach(), rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(kg.grad.cpu(), kr.grad, rtol=1e-3, atol=1e-5)
    torch.testing.assert_close(vg.grad.cpu(), vr.grad, rtol=1e-3, atol=1e-5)
    torch.testing.assert_close(qg.grad.cpu(), qr.grad, rtol=1e-3, atol=1e-4)


def _stem_s2d_weights(w):
    """(64, C, 7, 7) -> (64, 256) with k = (a*4+b)*16 + (dy*2+dx)*4 + c = w[2a+dy-1, 2b+dx-1, c]."""
    Cout, C = w.shape[:2]
    out = torch.zeros(Cout, 4, 4, 2, 2, 4)
    for a in range(4):
        for dy in range(2):
            r = 2 * a + dy - 1
            if not 0 <= r < 7:
                continue
            for b in range(4):
                for dx in range(2):
                    s = 2 * b + dx - 1
                    if 0 <= s < 7:
                        out[:, a, b, dy, dx, :C] = w[:, :, r, s]
    return out.reshape(Cout, 256)


@pytest.mark.parametrize('C,H,W,P', [(3, 100, 100, 9), (1, 50, 50, 37), (3, 20, 36, 5)])
def test_stem_s2d(dev, C, H, W, P):
    """7x7/2 stem as a shifted-window 4x4 convolution on the space-to-depth frame + strided max-pool."""
    from ips_b200 import ops
    x = _rand(P, C, H, W, seed=80)
    w = _rand(64, C, 7, 7, seed=81, scale=math.sqrt(2.0 / (49 * C))).to(torch.bfloat16)
    scale, shift = torch.rand(64) + 0.5, _rand(64, seed=82, scale=0.1)
    frame = ops.stage_patches_s2d(x.to(dev), P, C, H, W)
    y = ops.conv_stem_s2d(frame, _stem_s2d_weights(w.float()).to(torch.bfloat16).to(dev), scale.to(dev), shift.to(dev), P, H, W)
    Ho, Wo, Wp = H // 2, W // 2, W // 2 + 3
    Sp = (Ho + 3) * Wp
    got = y.view(P, Sp, 64)[:, :Ho * Wp].view(P, Ho, Wp, 64)[:, :, :Wo].cpu().float()
    xb = x.to(torch.bfloat16).float()
    ref = torch.relu(F.conv2d(xb, w.float(), stride=2, padding=3) * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)).permute(0, 2, 3, 1)
    torch.testing.assert_close(got, ref, rtol=1e-2, atol=1e-2)
    pooled = ops.maxpool3x3s2_pf_strided(y, P, Ho, Wo, 64, Wp, Sp)
    Hq, Wq = (Ho - 1) // 2 + 1, (Wo - 1) // 2 + 1
    refp = F.max_pool2d(got.permute(0, 3, 1, 2), 3, 2, 1).permute(0, 2, 3, 1).to(torch.bfloat16)
    assert torch.equal(ops.from_pf(pooled, P, Hq, Wq).cpu(), refp)


@pytest.mark.parametrize('C,H,W,P', [(3, 100, 100, 9), (1, 50, 50, 37), (3, 20, 36, 5), (3, 100, 100, 160)])
def test_stem_pool_fused(dev, C, H, W, P):
    """Fused stem + max-pool kernel == unfused stem kernel followed by the strided max-pool, bit for bit."""
    from ips_b200 import ops
    x = _rand(P, C, H, W, seed=90)
    w = _stem_s2d_weights(_rand(64, C, 7, 7, seed=91, scale=math.sqrt(2.0 / (49 * C)))).to(torch.bfloat16).to(dev)
    scale, shift = (torch.rand(64) + 0.5).to(dev), _rand(64, seed=92, scale=0.1).to(dev)
    frame = ops.stage_patches_s2d(x.to(dev), P, C, H, W)
    Ho, Wo, Wp = H // 2, W // 2, W // 2 + 3
    y = ops.conv_stem_s2d(frame, w, scale, shift, P, H, W)
    ref = ops.maxpool3x3s2_pf_strided(y, P, Ho, Wo, 64, Wp, (Ho + 3) * Wp)
    got = ops.stem_pool_s2d(frame, w, scale, shift, P, H, W)
    assert got.shape == ref.shape
    assert torch.equal(got, ref)


@pytest.mark.parametrize('P,H,W,C', [(3, 50, 50, 64), (2, 7, 9, 8), (1, 2, 2, 4), (5, 25, 24, 12)])
def test_maxpool_and_add_relu_autograd_fns(dev, P, H, W, C):
    """MaxPoolFn / AddReluFn (grad-mode encoder glue on the library's kernels) against torch autograd, bit for bit --
    including the tie rule of the pooling backward (post-ReLU maps are full of equal zeros: the first maximum in scan
    order takes the gradient, like nn.MaxPool2d)."""
    from ips_b200.autograd import MaxPoolFn, AddReluFn
    g = torch.Generator().manual_seed(P * 100 + H)
    x = torch.relu(torch.randn(P, H, W, C, generator=g)).to(dev)          # ~half zeros: ties in most windows
    x[0, :, :, 0] = 1.0                                                   # a constant channel: every window is one big tie
    xa = x.clone().requires_grad_(True)
    xb = x.clone().requires_grad_(True)
    ya = MaxPoolFn.apply(xa)
    yb = torch.nn.functional.max_pool2d(xb.permute(0, 3, 1, 2), 3, 2, 1).permute(0, 2, 3, 1)
    assert torch.equal(ya, yb)
    dy = torch.randn(ya.shape, generator=g).to(dev)
    ya.backward(dy)
    yb.backward(dy)
    assert torch.equal(xa.grad, xb.grad)
    a = torch.randn(P, H, W, C, generator=g).to(dev)
    b = torch.randn(P, H, W, C, generator=g).to(dev)
    a1, b1, a2, b2 = (t.clone().requires_grad_(True) for t in (a, b, a, b))
    y1 = AddReluFn.apply(a1, b1)
    y2 = torch.relu(a2 + b2)
    assert torch.equal(y1, y2)
    dy = torch.randn(y1.shape, generator=g).to(dev)
    y1.backward(dy)
    y2.backward(dy)
    assert torch.equal(a1.grad, a2.grad) and torch.equal(b1.grad, b2.grad)


@pytest.mark.parametrize('case', [(64, 64, 3, 1, 1, 13, 13), (64, 128, 3, 2, 1, 13, 13), (64, 128, 1, 2, 0, 13, 13),
                                  (128, 256, 3, 2, 1, 7, 7), (256, 256, 3, 1, 1, 4, 4), (256, 512, 1, 2, 0, 7, 7),
                                  (64, 128, 3, 2, 1, 14, 14), (64, 128, 1, 2, 0, 14, 10)])      # even maps under stride 2
def test_conv_autograd_fn(dev, case):
    """ConvFn: forward, grad input (same kernel, flipped weights, zero-dilated dy for stride 2) and grad weight
    (im2col x dy, TN GEMM) against torch autograd on the same bf16-rounded operands."""
    from ips_b200.autograd import ConvFn
    Cin, Cout, k, s, p, H, W = case
    P = 11
    x = _rand(P, H, W, Cin, seed=60).to(torch.bfloat16).float()
    w = _rand(Cout, Cin, k, k, seed=61, scale=math.sqrt(2.0 / (Cin * k * k))).to(torch.bfloat16).float()
    xr = x.clone().requires_grad_(True)
    wr = w.clone().requires_grad_(True)
    yr = F.conv2d(xr.permute(0, 3, 1, 2), wr, stride=s, padding=p).permute(0, 2, 3, 1)
    dy = _rand(*yr.shape, seed=62).to(torch.bfloat16).float()
    yr.backward(dy)
    xd = x.to(dev).requires_grad_(True)
    wd = w.to(dev).requires_grad_(True)
    y = ConvFn.apply(xd, wd, s, p)
    y.backward(dy.to(dev))
    torch.testing.assert_close(y.detach().cpu(), yr.detach(), rtol=1e-2, atol=1e-2)
    for got, ref in ((xd.grad.cpu(), xr.grad), (wd.grad.cpu(), wr.grad)):
        assert got.shape == ref.shape
        err = (got - ref).abs().max() / ref.abs().max()
        assert err < 1e-2, err


@pytest.mark.parametrize('C,H,W', [(3, 100, 100), (1, 50, 50), (3, 20, 36)])
def test_stem_conv_autograd_fn(dev, C, H, W):
    from ips_b200.autograd import StemConvFn
    P = 7
    x = _rand(P, C, H, W, seed=63).to(torch.bfloat16).float()
    w = _rand(64, C, 7, 7, seed=64, scale=math.sqrt(2.0 / (49 * C))).to(torch.bfloat16).float()
    wr = w.clone().requires_grad_(True)
    yr = F.conv2d(x, wr, stride=2, padding=3).permute(0, 2, 3, 1)
    dy = _rand(*yr.shape, seed=65).to(torch.bfloat16).float()
    yr.backward(dy)
    wd = w.to(dev).requires_grad_(True)
    y = StemConvFn.apply(x.to(dev), wd)
    y.backward(dy.to(dev))
    torch.testing.assert_close(y.detach().cpu(), yr.detach(), rtol=1e-2, atol=1e-2)
    err = (wd.grad.cpu() - wr.grad).abs().max() / wr.grad.abs().max()
    assert err < 1e-2, err


@pytest.mark.parametrize('Cin,Cout,H,W,P', [(256, 256, 7, 7, 256), (256, 512, 7, 7, 512), (512, 256, 5, 9, 256)])
def test_conv_im2col_pair(dev, Cin, Cout, H, W, P):
    """Dense stride-1 3x3 convolution through the im2col-mode CTA-pair kernel (128 consecutive pixels per tile)."""
    from ips_b200 import ops
    x = _rand(P, H, W, Cin, seed=70).to(torch.bfloat16)
    w = _rand(Cout, Cin, 3, 3, seed=71, scale=math.sqrt(2.0 / (Cin * 9))).to(torch.bfloat16)
    scale, shift = torch.rand(Cout) + 0.5, _rand(Cout, seed=72, scale=0.1)
    res = _rand(P, H, W, Cout, seed=73).to(torch.bfloat16)
    w_nk = w.permute(0, 2, 3, 1).reshape(Cout, 9 * Cin).contiguous().to(dev)
    assert (P * H * W) % 256 == 0                                   # eligible: an even number of full tiles
    for use_res, relu in ((False, True), (True, True), (False, False)):
        ref = _conv_ref(x.float(), w.float(), scale, shift, res.float() if use_res else None, 1, 1, relu)
        y = ops.conv_bf16_pf(x.to(dev), w_nk, scale.to(dev), shift.to(dev), res.to(dev) if use_res else None, P, H, W, Cout, 3, 3, 1, 1,
                             relu, False, False)
        torch.testing.assert_close(y.cpu().float(), ref, rtol=1e-2, atol=1e-2)

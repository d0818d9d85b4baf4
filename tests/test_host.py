"""CPU-side checks: the C-ABI library loads and exports every declared symbol, the drop-in
module keeps the reference's surface, and the grad-mode forward reproduces the golden
train step.  No kernel is launched here."""
import os
import sys
import re

import numpy as np
import pytest
import torch

import ips_oracle as O
from golden_util import CASE_NAMES, load_case

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from ips_b200 import _lib, build
    build.build()
    lib = _lib.load()
    header = open(os.path.join(ROOT, 'include', 'ips_b200.h')).read()
    declared = set(re.findall(r'IPSB_API\s+[\w\s\*]+?\b(ipsb_\w+)\s*\(', header))
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.ipsb_abi_version() == 1


def test_no_cpu_fallback():
    from ips_b200 import IPSNet, Struct, ops
    conf = O.preset('camelyon', M=4, I=4)
    net = IPSNet(torch.device('cpu'), Struct(**conf.__dict__))
    with pytest.raises(RuntimeError):
        net.ips(torch.randn(1, 16, 2048))
    with pytest.raises(RuntimeError):
        ops.topm_stable(torch.rand(2, 8), 3)


@pytest.mark.parametrize('pre', ['mnist', 'traffic', 'camelyon'])
def test_state_dict_surface(pre):
    from ips_b200 import IPSNet, Struct
    conf = O.preset(pre)
    net = IPSNet(torch.device('cpu'), Struct(**conf.__dict__))
    want = O.param_shapes(conf)
    got = [(k, tuple(v.shape)) for k, v in net.state_dict().items()]
    assert got == [(k, tuple(s)) for k, s in want]
    assert ('pos_enc' not in net.state_dict())
    assert (net.pos_enc is None) == (not conf.use_pos)
    for attr in ('M', 'I', 'D', 'encoder', 'transf', 'score_and_select', 'ips', 'output_layers'):
        assert hasattr(net, attr)


def test_camelyon_config_needs_no_image_keys():
    from ips_b200 import IPSNet, Struct
    conf = O.preset('camelyon')
    for k in ('N', 'n_res_blocks', 'patch_size', 'patch_stride'):
        assert not hasattr(conf, k)
    IPSNet(torch.device('cpu'), Struct(**conf.__dict__))


def test_shortcut_when_memory_covers_all_patches():
    from ips_b200 import IPSNet, Struct
    z, meta, conf, sd, patches = load_case('camelyon_shortcut')
    net = IPSNet(torch.device('cpu'), Struct(**conf.__dict__))
    mem_patch, mem_pos = net.ips(patches)
    assert torch.equal(mem_patch, patches) and mem_pos is None


def test_scan_order_consumes_rng_like_reference():
    from ips_b200.utils import scan_order
    torch.manual_seed(9)
    perm, per = scan_order(True, 'batch', 4, 50, torch.device('cpu'))
    torch.manual_seed(9)
    assert torch.equal(perm[0], torch.randperm(50)) and not per
    torch.manual_seed(9)
    perm, per = scan_order(True, 'instance', 4, 50, torch.device('cpu'))
    torch.manual_seed(9)
    assert torch.equal(perm, torch.rand(4, 50).argsort(1)) and per
    assert scan_order(False, 'batch', 4, 50, torch.device('cpu')) == (None, False)


@pytest.mark.parametrize('name', [n for n in CASE_NAMES if 'shortcut' not in n])
def test_grad_mode_forward_matches_golden(name):
    """IPSNet.forward + the reference's loss on the golden selection (train-mode BN, dropout 0)."""
    from ips_b200 import IPSNet, Struct
    z, meta, conf, sd, patches = load_case(name)
    net = IPSNet(torch.device('cpu'), Struct(**conf.__dict__))
    net.load_state_dict(sd)
    net.train()
    torch.manual_seed(meta['rng_seed'])
    mem_patch, mem_pos, _ = O.ips(sd, conf, patches, perm='draw', tie='topk')
    preds = net(mem_patch, mem_pos)
    loss = O.loss_fn(conf, preds, O.make_labels(conf, meta['B'], meta['label_seed']))
    loss.backward()
    assert abs(loss.item() - float(z['loss'])) <= 1e-5 * max(1.0, abs(float(z['loss'])))
    g = net.transf.crs_attn.q.grad.reshape(-1)[:256].numpy()
    np.testing.assert_allclose(g, z['grad_transf.crs_attn.q'], rtol=1e-3, atol=1e-7)


def test_feature_bag_roundtrip(tmp_path):
    """Flat feature-bag file: write fp32 / bf16 bags, read them back memory-mapped with the item layout
    of the reference's CamelyonFeatures."""
    import torch
    from ips_b200.io import FeatureBagWriter, FeatureBags
    g = torch.Generator().manual_seed(3)
    slides = [torch.randn(n, 64, generator=g) for n in (5, 130, 1)]
    tasks = {'task0': {'name': 'metastases'}}
    for dtype in ('fp32', 'bf16'):
        path = str(tmp_path / ('bags_%s.bin' % dtype))
        with FeatureBagWriter(path, 64, dtype) as w:
            for i, s in enumerate(slides):
                w.add('slide%d' % i, s, label=i % 2)
        ds = FeatureBags(path, tasks)
        assert len(ds) == 3 and ds.slide_names == ['slide0', 'slide1', 'slide2']
        for i, s in enumerate(slides):
            item = ds[i]
            want = s if dtype == 'fp32' else s.to(torch.bfloat16)
            assert item['input'].dtype == want.dtype and torch.equal(item['input'], want)
            assert item['metastases'] == i % 2


def test_auto_chunk_policy_and_image_geometry():
    """Host-side planning: rows are split evenly over the lanes under the workspace cap; the patch grid of an image
    follows the reference's unfold (n = (H - ph) // sh + 1 rows of (W - pw) // sw + 1 patches)."""
    import torch
    import ips_oracle as O
    from ips_b200 import IPSNet, Struct, ops
    conf = O.preset('traffic')
    net = IPSNet(torch.device('cpu'), Struct(**conf.__dict__))
    assert net.lanes == 2
    assert net._auto_chunk((3, 100, 100), 3072) == 1536               # two lanes, one chunk each
    assert net._auto_chunk((3, 100, 100), 3000) == 1504               # rounded up to a multiple of 8
    assert net._auto_chunk((3, 100, 100), 20000) <= 2048              # capped by the activation workspace
    assert net._auto_chunk((3, 100, 100), 10) == 32                   # floor
    net.chunk_patches = 77
    assert net._auto_chunk((3, 100, 100), 3072) == 77                 # explicit override
    img = torch.zeros(2, 3, 1200, 1600)
    geo, n = ops.image_geo(img, (100, 100), (100, 100))
    assert n == 192 and geo.n_per_image == 192 and (geo.img_h, geo.img_w, geo.stride_h, geo.stride_w) == (1200, 1600, 100, 100)
    geo, n = ops.image_geo(torch.zeros(1, 1, 250, 300), (50, 50), (25, 25))
    assert n == 9 * 11


def test_ips_out_buffer_validation():
    import torch
    import ips_oracle as O
    from ips_b200 import IPSNet, Struct
    conf = O.preset('camelyon', M=8, I=8)
    net = IPSNet(torch.device('cpu'), Struct(**conf.__dict__))
    good = torch.zeros(4, 8, conf.n_chan_in)
    a, b = net._out_views((good, None), 1, 2, 8)
    assert a.shape == (2, 8, conf.n_chan_in) and b is None and a.data_ptr() == good[1].data_ptr()
    import pytest
    with pytest.raises(ValueError):
        net._out_views((good, None), 3, 2, 8)                         # rows 3..4 do not fit 4 rows
    with pytest.raises(ValueError):
        net._out_views((torch.zeros(4, 7, conf.n_chan_in), None), 0, 2, 8)


def test_interval_union_of_overlapping_lanes():
    """bench.py's roofline takes the UNION of a kernel family's launch intervals over the concurrent lanes."""
    from ips_b200 import ops
    assert ops.busy_ms([]) == 0.0
    assert ops.busy_ms([(0.0, 1.0), (0.5, 2.0), (3.0, 4.0)]) == pytest.approx(3.0)
    assert ops.busy_ms([(2.0, 3.0), (0.0, 5.0)]) == pytest.approx(5.0)
    assert ops.busy_ms([(0.0, 1.0), (1.0, 2.0)]) == pytest.approx(2.0)


def test_reference_copy_is_the_unmodified_reference():
    """baseline/_ref (what `bench.py --impl reference` and the call-site tests import) is a byte-for-byte copy."""
    import filecmp
    sys.path.insert(0, os.path.join(ROOT, 'baseline'))
    import install_ref
    import ref_harness
    if install_ref.ref_path() is None:
        pytest.skip('baseline/_ref not installed (run baseline/install_ref.py where /root/reference exists)')
    if os.path.isdir(install_ref.SRC):
        for rel in ('architecture/ips_net.py', 'architecture/transformer.py', 'utils/utils.py', 'training/iterative.py'):
            assert filecmp.cmp(os.path.join(install_ref.SRC, rel), os.path.join(install_ref.DST, rel), shallow=False), rel
    IPSNet, Struct, iterative = ref_harness.load_reference()
    conf = ref_harness.reference_conf('camelyon', M=4, I=4)
    net = IPSNet(torch.device('cpu'), conf)
    assert os.path.abspath(sys.modules[IPSNet.__module__].__file__).startswith(os.path.abspath(install_ref.DST))
    mp_, pos = net.ips(torch.randn(1, 16, 2048))
    assert mp_.shape == (1, 4, 2048) and pos is None and hasattr(iterative, 'train_one_epoch')


def test_pretrained_is_honoured_or_raises():
    """conf.pretrained must not be dropped silently (ips_net.py:19-27): either the ImageNet weights load or the constructor raises."""
    from ips_b200 import IPSNet, Struct
    conf = O.preset('traffic', pretrained=True)
    try:
        net = IPSNet(torch.device('cpu'), Struct(**conf.__dict__))
    except RuntimeError as e:
        assert 'pretrained' in str(e)
        return
    ref = IPSNet(torch.device('cpu'), Struct(**O.preset('traffic').__dict__))
    assert not torch.equal(net.encoder[0].weight, ref.encoder[0].weight)      # not a random init


def test_sharded_exchange_layout_is_identical_on_all_ranks():
    """Section offsets of the peer exchange buffer depend only on the global shapes, never on the rank."""
    from ips_b200.distributed import _align, shard_bounds
    assert _align(1) == 256 and _align(256) == 256 and _align(257) == 512
    for N, R in ((50000, 8), (200000, 8), (20001, 2), (7, 3)):
        b = shard_bounds(N, R)
        assert b[0][0] == 0 and b[-1][1] == N and all(x[1] == y[0] for x, y in zip(b, b[1:]))
        assert max(hi - lo for lo, hi in b) - min(hi - lo for lo, hi in b) <= 1


def test_keyed_scan_order_oracle_is_a_uniform_looking_bijection():
    """The device-side scan order's CPU restatement: a permutation for every N; over many keys every element lands on
    every position equally often (chi-square on a 32 x 32 table) and a chunk's membership looks like a random sample."""
    import numpy as np
    for N in (1, 2, 3, 5, 64, 1000, 4097):
        p = O.keyed_scan_order(0x1234 + N, -77, 2, N)
        assert torch.equal(p.sort(1).values, torch.arange(N).expand(2, N))
    N, T = 32, 6000
    cnt = np.zeros((N, N))
    for t in range(T):
        p = O.keyed_scan_order(99, t, 1, N)[0].numpy()
        cnt[np.arange(N), p] += 1
    chi = ((cnt - T / N) ** 2 / (T / N)).sum()
    dof = (N - 1) ** 2
    assert abs(chi - dof) < 5 * (2 * dof) ** 0.5, chi
    fr = [(O.keyed_scan_order(5, t, 1, 20000)[0][:2000] < 10000).double().mean().item() for t in range(30)]
    assert abs(np.mean(fr) - 0.5) < 0.01 and np.std(fr) < 0.02


def test_streamed_select_shape_gate():
    """ops.streamed_select_ok mirrors ipsb_select_loop_scan's -2 cases: only the 8-CTA cluster loop (M + I >= 2048, H*T a
    power of two, slices inside shared memory) runs side by side with the projector."""
    from ips_b200 import ops
    assert ops.streamed_select_ok(1, 50000, 8, 5000, 5000)          # CAMELYON
    assert ops.streamed_select_ok(16, 6250, 8, 5000, 5000)          # a rank's slice of 16 slides on 8 GPUs (one iteration)
    assert not ops.streamed_select_ok(16, 192, 1, 10, 32)           # traffic: tiny buffers -> rank-counting kernel
    assert not ops.streamed_select_ok(1, 5000, 8, 5000, 5000)       # M >= N: the caller's shortcut
    assert not ops.streamed_select_ok(1, 50000, 12, 5000, 5000)     # H*T not a power of two
    assert not ops.streamed_select_ok(1, 400000, 32, 60000, 60000)  # slices beyond shared memory

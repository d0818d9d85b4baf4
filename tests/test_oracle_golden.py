"""Pin the oracle to outputs of the unmodified reference (tests/golden, made by
oracle/gen_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

import ips_oracle as O
from golden_util import CASE_NAMES, load_case


@pytest.mark.parametrize('name', CASE_NAMES)
def test_ips_matches_reference(name):
    z, meta, conf, sd, patches = load_case(name)
    torch.manual_seed(meta['rng_seed'])
    mem_patch, mem_pos, src = O.ips(sd, conf, patches, perm='draw', tie='topk')
    assert list(mem_patch.shape) == list(z['mem_patch_shape'])
    assert mem_patch.double().sum().item() == float(z['mem_patch_sum'])
    if 'mem_patch' in z:
        assert np.array_equal(mem_patch.numpy(), z['mem_patch'])
    if 'mem_src' in z and z['mem_src'].size:
        assert np.array_equal(src.numpy(), z['mem_src'])
    if 'mem_pos' in z:
        assert np.array_equal(mem_pos.numpy(), z['mem_pos'])
    else:
        assert mem_pos is None


@pytest.mark.parametrize('name', CASE_NAMES)
def test_units_match_reference(name):
    z, meta, conf, sd, patches = load_case(name)
    B = meta['B']
    g = torch.Generator().manual_seed(101)
    L = min(conf.M + conf.I, 64)
    emb = torch.randn(B, L, conf.D, generator=g)
    s = O.attn_scores(sd, conf, emb)
    assert np.array_equal(s.numpy(), z['unit_emb_scores'])
    assert torch.allclose(s.sum(-1), torch.ones(B), atol=1e-5)
    x = patches[:, :3].reshape(-1, *patches.shape[2:])
    assert np.array_equal(O.encode(sd, conf, x).numpy(), z['unit_encode'])
    idx = torch.arange(L).unsqueeze(0).expand(B, -1)
    _, si, _ = O.score_and_select(sd, conf, emb, None, max(1, L // 3), idx)
    assert np.array_equal(si.numpy(), z['unit_select_idx'])
    # where no tie straddles the boundary the stable contract picks the same set
    _, ss, sc = O.score_and_select(sd, conf, emb, None, max(1, L // 3), idx, tie='stable')
    assert np.array_equal(np.sort(ss.numpy(), -1), np.sort(si.numpy(), -1))


@pytest.mark.parametrize('name', CASE_NAMES)
def test_train_step_matches_reference(name):
    z, meta, conf, sd, patches = load_case(name)
    torch.manual_seed(meta['rng_seed'])
    mem_patch, mem_pos, _ = O.ips(sd, conf, patches, perm='draw', tie='topk')
    for k, v in sd.items():
        if v.is_floating_point() and 'running' not in k:
            v.requires_grad_(True)
    preds = O.forward(sd, conf, mem_patch, mem_pos, train=True)
    loss = O.loss_fn(conf, preds, O.make_labels(conf, meta['B'], meta['label_seed']))
    loss.backward()
    assert abs(loss.item() - float(z['loss'])) <= 1e-6 * max(1.0, abs(float(z['loss'])))
    for task in conf.tasks.values():
        np.testing.assert_allclose(preds[task['name']].detach().numpy().squeeze(-1) if conf.n_class == 1
                                   else preds[task['name']].detach().numpy(),
                                   z['pred_' + task['name']], rtol=1e-5, atol=1e-7)
    for key in z.files:
        if key.startswith('grad_'):
            got = sd[key[5:]].grad.reshape(-1)[:256].numpy()
            np.testing.assert_allclose(got, z[key], rtol=1e-4, atol=1e-7)
    bn = 'encoder.1.running_mean' if conf.is_image else 'encoder.2.running_mean'
    np.testing.assert_allclose(sd[bn].numpy(), z['bn_running_mean_after'], rtol=1e-5, atol=1e-7)


def test_sharded_schedule_single_rank_equals_sequential():
    z, meta, conf, sd, patches = load_case('camelyon_batch')
    a = O.ips(sd, conf, patches, perm=None, tie='stable')
    b = O.ips_sharded(sd, conf, patches, R=1, perm=None, tie='stable')
    assert torch.equal(a[2], b[2]) and torch.equal(a[0], b[0])


def test_sharded_exact_when_single_head_token():
    """with H = n_token = 1 the sharded schedule selects the same set."""
    conf = O.preset('camelyon', M=16, I=16, H=1, D_k=64, D_v=64, D=64, n_chan_in=96, D_inner=128)
    sd = O.make_state(conf, 3, q_gain=12.0)
    x = O.make_patches(conf, 2, 200, 4)
    a = O.ips(sd, conf, x, perm=None, tie='stable')[2]
    for R in (2, 4):
        b = O.ips_sharded(sd, conf, x, R=R, perm=None, tie='stable')[2]
        assert torch.equal(a.sort(-1)[0], b.sort(-1)[0])

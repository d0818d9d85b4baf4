"""Call-site parity: the reference's OWN training loop -- `train_one_epoch` / `evaluate` with
`init_batch`, `fill_batch`, `shrink_batch`, `compute_loss` (training/iterative.py:7-231, run unmodified from
baseline/_ref) -- drives `ips_b200.IPSNet` exactly as it drives the reference class, and logs the same losses.

Loader batches are synthetic dicts of the layout the reference datasets emit (`input` + one entry per task).  The
run covers a full train batch assembled from several `ips()` calls, the short last batch (`shrink_batch`, :149-151)
and the `M >= N` shortcut whose zero padding is encoded and attended (:37-38)."""
import os
import sys

import pytest
import torch

import ips_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'baseline'))

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


class _Log:
    """Stand-in for utils.Logger: keeps what the loop reports per optimizer step."""

    def __init__(self):
        self.losses, self.preds = [], []

    def update(self, task_losses, task_preds, task_labels):
        self.losses.append(dict(task_losses))
        self.preds.append({k: v.copy() for k, v in task_preds.items()})


def _loader(conf, n_batches, B_seq, N, seed):
    out = []
    for i in range(n_batches):
        x = O.make_patches(conf, B_seq, N, seed + i)
        batch = {'input': x}
        batch.update(O.make_labels(conf, B_seq, seed + 100 + i))
        out.append(batch)
    return out


def _criterions(conf):
    return {t['name']: (torch.nn.NLLLoss() if t['act_fn'] == 'softmax' else torch.nn.BCELoss()) for t in conf.tasks.values()}


def _run(make_net, iterative, Struct, conf, sd, loader, train):
    torch.manual_seed(0)
    net = make_net(Struct(**conf.__dict__))
    net.load_state_dict(sd, strict=True)
    log = _Log()
    torch.manual_seed(17)                                   # the 'batch' shuffle draws randperm from the global CPU generator
    if train:
        opt = torch.optim.AdamW(net.parameters(), lr=0, weight_decay=conf.wd)          # main.py:57
        iterative.train_one_epoch(net, _criterions(conf), loader, opt, torch.device(DEV), 0, log, conf)
    else:
        iterative.evaluate(net, _criterions(conf), loader, torch.device(DEV), log, conf)
    return log, net


def _reference():
    import ref_harness
    try:
        return ref_harness.load_reference()
    except ref_harness.ReferenceUnavailable as e:
        pytest.skip(str(e))


@pytest.mark.parametrize('pre,over,N', [
    ('mnist', dict(N=36, M=8, I=8, B=4, B_seq=2), 36),              # 3 loader batches: one full train batch + a short one
    ('mnist', dict(N=6, M=8, I=8, B=4, B_seq=2), 6),                # M >= N: the shortcut, zero-padded rows in the train batch
    ('camelyon', dict(M=32, I=48, B=4, B_seq=1), 333),              # B_seq = 1: four ips() calls per train batch
    ('traffic', dict(N=12, M=4, I=5, B=2, B_seq=2), 12),
])
@pytest.mark.parametrize('train', [True, False])
def test_reference_training_loop_drives_the_drop_in(pre, over, N, train):
    RefNet, Struct, iterative = _reference()
    from ips_b200 import IPSNet
    torch.backends.cudnn.allow_tf32 = False                 # the reference's fp32 convolutions stay fp32 on the GPU
    torch.backends.cuda.matmul.allow_tf32 = False
    conf = O.preset(pre, attn_dropout=0.0, dropout=0.0, n_epoch=2, n_epoch_warmup=1, precision='fp32', **over)
    sd = O.make_state(conf, 11, q_gain=12.0)
    n_batches = 3 if conf.B_seq > 1 else 6
    loader = _loader(conf, n_batches, conf.B_seq, N, 300)
    dev = torch.device(DEV)
    ref_log, ref_net = _run(lambda c: RefNet(dev, c).to(dev), iterative, Struct, conf, sd, loader, train)
    our_log, our_net = _run(lambda c: IPSNet(dev, c).to(dev), iterative, Struct, conf, sd, loader, train)
    assert len(our_log.losses) == len(ref_log.losses) >= 2
    for step, (a, b) in enumerate(zip(our_log.losses, ref_log.losses)):
        for task in b:
            tol = 2e-4 if step == 0 else 2e-3               # later steps compare two optimizers' fp32 trajectories
            assert abs(a[task] - b[task]) <= tol * max(1.0, abs(b[task])), (step, task, a[task], b[task])
    for a, b in zip(our_log.preds, ref_log.preds):           # the batch shapes (incl. the shrunken last batch) agree
        for task in b:
            assert a[task].shape == b[task].shape
    if train:                                               # both nets took the same optimizer steps
        ours = our_net.state_dict()
        for k, v in ref_net.state_dict().items():
            w = ours[k]
            if v.dtype.is_floating_point:
                assert float((v - w).abs().max()) <= 5e-3 * (1.0 + float(v.abs().max())), k

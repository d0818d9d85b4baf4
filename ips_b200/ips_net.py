"""Drop-in `IPSNet` whose selection path runs on hand-written sm_100a kernels.

Same constructor, `ips` / `forward` signatures, attribute names and state_dict
keys as architecture/ips_net.py:85-283 of the reference, so `main.py` and
`training/iterative.py` can use it unchanged.

How `ips` differs inside (results are the same):
  * a patch's pre-softmax logits are a fixed property of the patch, so every patch
    is encoded and projected ONCE into a (B, N, H*T) logit table; the sequential
    M+I loop (ips_net.py:218-241) then runs inside one kernel on that table;
  * shuffling (utils/utils.py:33-58) is a scan order handed to the kernels, not a
    copy of the patch tensor; the same RNG calls are made so seeds reproduce it;
  * the memory buffer is persistent shared memory, not torch.cat + gather.
"""
import math
import os

import torch
from torch import nn

from . import ops
from .autograd import Linear, LayerNormFn, BatchNormTrainFn, HeadLossFn, conv_encoder_train
from .transformer import Transformer, pos_enc_1d
from .utils import scan_order

_BN_EPS = 1e-5


def _on_device(fn):
    """Run a method with the module's CUDA device current: the library launches on the CURRENT device's stream and keeps
    per-device internal streams, so `IPSNet(device='cuda:1')` must work without `torch.cuda.set_device(1)`."""
    import functools

    @functools.wraps(fn)
    def wrapped(self, *a, **kw):
        dev = torch.device(self.device)
        if dev.type != 'cuda' or (dev.index is not None and dev.index == torch.cuda.current_device()):
            return fn(self, *a, **kw)
        with torch.cuda.device(dev):
            return fn(self, *a, **kw)
    return wrapped


# ----------------------------------------------------------------------------------------
# parameter containers (state_dict compatible with the truncated torchvision ResNet)
# ----------------------------------------------------------------------------------------

class _BasicBlock(nn.Module):
    def __init__(self, cin, cout, stride):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, cout, 3, stride, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(cout)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = nn.Conv2d(cout, cout, 3, 1, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(cout)
        self.downsample = None
        if stride != 1 or cin != cout:
            self.downsample = nn.Sequential(nn.Conv2d(cin, cout, 1, stride, bias=False), nn.BatchNorm2d(cout))
        self.stride = stride

    def forward(self, x):                      # grad-mode path (train step)
        idt = x if self.downsample is None else self.downsample(x)
        y = self.relu(self.bn1(self.conv1(x)))
        return self.relu(self.bn2(self.conv2(y)) + idt)


def _imagenet_resnet18_state():
    """ImageNet weights of torchvision's resnet18 (`ResNet18_Weights.IMAGENET1K_V1`, ips_net.py:19-27), renamed to the
    children of the truncated trunk.  Raises -- never continues with random weights -- when they cannot be had."""
    try:
        from torchvision.models import resnet18, ResNet18_Weights
        sd = resnet18(weights=ResNet18_Weights.IMAGENET1K_V1).state_dict()
    except Exception as e:
        raise RuntimeError('ips_b200: conf.pretrained is set but the ImageNet ResNet-18 weights are not available '
                           '(torchvision missing, or no cached checkpoint and no network): %s.  Set `pretrained: False` '
                           'to train the patch encoder from random initialisation.' % (e,)) from e
    ren = {'conv1': '0', 'bn1': '1', 'layer1': '4', 'layer2': '5', 'layer3': '6', 'layer4': '7'}
    return {ren[k.split('.', 1)[0]] + '.' + k.split('.', 1)[1]: v for k, v in sd.items() if k.split('.', 1)[0] in ren}


def _conv_patch_encoder(enc_type, n_chan_in, n_res_blocks, pretrained=False):
    """Children 0,1,2,3,4,5[,6,7],avgpool exactly as ips_net.py:34-50 composes them."""
    if enc_type != 'resnet18':
        raise NotImplementedError("only enc_type 'resnet18' is built (resnet50 is unused by the shipped configs)")
    layers = [nn.Conv2d(n_chan_in, 64, 7, 2, 3, bias=False), nn.BatchNorm2d(64), nn.ReLU(inplace=True),
              nn.MaxPool2d(3, 2, 1)]
    cin = 64
    for li, w in enumerate([64, 128, 256, 512][:2 if n_res_blocks != 4 else 4]):
        s = 1 if li == 0 else 2
        layers.append(nn.Sequential(_BasicBlock(cin, w, s), _BasicBlock(w, w, 1)))
        cin = w
    layers.append(nn.AdaptiveAvgPool2d((1, 1)))
    enc = nn.Sequential(*layers)
    for m in enc.modules():                    # torchvision's ResNet initialisation
        if isinstance(m, nn.Conv2d):
            nn.init.kaiming_normal_(m.weight, mode='fan_out', nonlinearity='relu')
    if pretrained:                             # ImageNet weights for every child that is kept (:19-27)
        have = enc.state_dict()
        enc.load_state_dict({k: v for k, v in _imagenet_resnet18_state().items() if k in have}, strict=True)
    if n_chan_in == 1:                         # the reference swaps in a default-initialised stem (:29-31)
        enc[0].reset_parameters()
    return enc


def _projector(n_chan_in, D):
    return nn.Sequential(nn.LayerNorm(n_chan_in, eps=1e-05, elementwise_affine=False), Linear(n_chan_in, D),
                         nn.BatchNorm1d(D), nn.ReLU())


def _fold_bn(bn):
    scale = bn.weight.detach().float() * torch.rsqrt(bn.running_var.detach().float() + bn.eps)
    shift = bn.bias.detach().float() - bn.running_mean.detach().float() * scale
    return scale.contiguous(), shift.contiguous()


class IPSNet(nn.Module):
    """Patch encoder + IPS + cross-attention aggregator + heads (ips_net.py:10-283)."""

    def get_output_layers(self, tasks):
        out = nn.ModuleDict()
        for task in tasks.values():
            act = nn.Softmax(dim=-1) if task['act_fn'] == 'softmax' else nn.Sigmoid()
            out[task['name']] = nn.Sequential(Linear(self.D, self.n_class), act)
        return out

    def __init__(self, device, conf):
        super().__init__()
        self.device = device
        self.n_class = conf.n_class
        self.M = conf.M
        self.I = conf.I
        self.D = conf.D
        self.use_pos = conf.use_pos
        self.tasks = conf.tasks
        self.shuffle = conf.shuffle
        self.shuffle_style = conf.shuffle_style
        self.is_image = conf.is_image
        # optional, not a reference key: 'bf16' (tcgen05, default), 'fp32' (CUDA-core exact mode) or 'bf16x3' (near-fp32 on
        # the tensor cores: operands as hi + lo bf16 pairs, three MMAs per product, fp32 between layers; csrc/split3.cu)
        self.precision = os.environ.get('IPS_B200_PRECISION', getattr(conf, 'precision', 'bf16'))
        if self.precision not in ('bf16', 'fp32', 'bf16x3'):
            raise ValueError(f'unknown precision {self.precision!r}')

        if self.is_image:
            self.encoder = _conv_patch_encoder(conf.enc_type, conf.n_chan_in, conf.n_res_blocks, bool(getattr(conf, 'pretrained', False)))
        else:
            self.encoder = _projector(conf.n_chan_in, self.D)
        self.transf = Transformer(conf.n_token, conf.H, conf.D, conf.D_k, conf.D_v, conf.D_inner,
                                  conf.attn_dropout, conf.dropout)
        self.pos_enc = pos_enc_1d(conf.D, conf.N).unsqueeze(0).to(device) if conf.use_pos else None
        self.output_layers = self.get_output_layers(conf.tasks)

        for m in self.modules():                     # train-step GEMMs follow the same precision switch as ips()
            if isinstance(m, Linear):
                m.precision = 'bf16' if self.precision == 'bf16' else 'fp32'      # (grad-mode GEMMs: bf16x3 trains in fp32)
        self._plan = None
        self._plan_key = None
        self.last_mem_idx = None      # (B,M) original-order indices of the last ips() call (notebook cell 9)
        self.chunk_patches = int(os.environ.get('IPS_B200_CHUNK', getattr(conf, 'chunk_patches', 0)))   # 0 = auto
        # chunks in flight inside the native executor (internal streams); 1 = strictly sequential
        self.lanes = int(os.environ.get('IPS_B200_LANES', getattr(conf, 'lanes', 2)))
        # data-parallel train step: synchronise BatchNorm statistics over the ranks (True = default process group) so a
        # batch split over R GPUs normalises like the whole batch in one process; off = per-rank statistics
        self.sync_bn = getattr(conf, 'sync_bn', False)
        # grad-mode conv encoder: 'native' = library kernels forward and backward (bf16 precision), 'torch' = cuDNN autograd
        self.train_encoder = os.environ.get('IPS_B200_TRAIN_ENCODER', getattr(conf, 'train_encoder', 'native'))
        # 'native': one C++ call runs the whole encoder; 'python': one library call per layer (per-kernel timing)
        self.executor = os.environ.get('IPS_B200_EXECUTOR', 'native')
        self._ws_cache = {}
        # feature bags: run the projector kernel and the selection loop side by side (ops.projector_select); False = one
        # after the other through the original-order logit table
        self.streamed_select = os.environ.get('IPS_B200_STREAMED_SELECT', '1') != '0' and getattr(conf, 'streamed_select', True)
        self.last_stream_status = None
        # 'reference': the scan order is drawn with the reference's RNG calls (CPU randperm: seeds reproduce its order);
        # 'device': drawn on the GPU (same distribution, other stream; no host work, capturable in a CUDA graph)
        self.scan_order_rng = os.environ.get('IPS_B200_SCAN_ORDER_RNG', getattr(conf, 'scan_order_rng', 'reference'))
        # lazy loading (`eager: False`): host inputs up to this size also keep a device copy for the final gather; larger
        # ones stream through a three-chunk ring so device memory stays O(chunk), independent of N
        self.lazy_resident_bytes = int(os.environ.get('IPS_B200_LAZY_RESIDENT_BYTES', getattr(conf, 'lazy_resident_bytes', 8 << 30)))
        # bf16 stem: 4 = shifted-window kernel on the space-to-depth frame, 3 = TMA-fed im2col rows (both need even
        # patch sizes), 1 = cp.async gather
        ps = getattr(conf, 'patch_size', [0, 0])
        self.patch_size = tuple(ps)
        stem = os.environ.get('IPS_B200_STEM', 's2d')
        even = self.is_image and ps[0] % 2 == 0 and ps[1] % 2 == 0
        self.stem_mode = {'s2d': 4, 'tma': 3}.get(stem, 1) if even else 1
        if self.stem_mode == 4 and ps[1] > 240:      # the shifted-window stem keeps 128 + 3*(W/2+3) + 3 frame rows per tile in two TMA boxes
            self.stem_mode = 1
        if self.stem_mode == 3 and max(ps) > 480:
            self.stem_mode = 1
        self.stem_tma = self.stem_mode >= 3
        ops.register_custom_ops()

    # ------------------------------------------------------------------ derived parameters
    def _state_versions(self):
        vs = [self.precision]
        for t in list(self.encoder.parameters()) + list(self.encoder.buffers()):
            vs.append((t.data_ptr(), t._version))
        for t in (self.transf.crs_attn.q, self.transf.crs_attn.q_w.weight, self.transf.crs_attn.k_w.weight):
            vs.append((t.data_ptr(), t._version))
        return tuple(vs)

    def _conv_entry(self, conv, bn, stem=False):
        """Geometry + destination buffers of one folded convolution and its `ipsb_fold_item` (filled by `ipsb_fold_plan`)."""
        from . import _lib
        w = conv.weight
        cout, cin, kh, kw = w.shape
        dev, bf16 = w.device, self.precision == 'bf16'
        e = dict(cin=4 if stem else cin, cout=cout, kh=kh, kw=kw, stride=conv.stride[0], pad=conv.padding[0],
                 mode=(self.stem_mode if bf16 else 1) if stem else 0)
        e['scale'] = torch.empty(cout, dtype=torch.float32, device=dev)
        e['shift'] = torch.empty(cout, dtype=torch.float32, device=dev)
        if stem and bf16:                                         # channels padded to 4; 8 x 8 taps
            layout = _lib.FOLD_STEM_S2D if e['mode'] == 4 else _lib.FOLD_STEM_8X8
            e['w'] = torch.empty((cout, 256), dtype=torch.bfloat16, device=dev)
            cin_pad = 4
            if e['mode'] == 4:
                e['cin'] = 16
        elif bf16:
            layout, cin_pad = _lib.FOLD_KMAJOR, cin
            e['w'] = torch.empty((cout, kh * kw * cin), dtype=torch.bfloat16, device=dev)
        else:
            layout, cin_pad = _lib.FOLD_KN, (4 if stem else cin)
            e['w'] = torch.empty((kh * kw * cin_pad, cout), dtype=torch.float32, device=dev)
        e['item'] = _lib.FoldItem(w.data_ptr(), e['w'].data_ptr(), bn.weight.data_ptr(), bn.bias.data_ptr(),
                                  bn.running_mean.data_ptr(), bn.running_var.data_ptr(), e['scale'].data_ptr(), e['shift'].data_ptr(),
                                  e['w'].numel(), cout, cin, kh, kw, cin_pad, layout, int(bf16), float(bn.eps))
        e['src'] = (w, bn.weight, bn.bias, bn.running_mean, bn.running_var)
        return e

    def _fold_static(self):
        """Destination buffers and the device-resident item table of the conv encoder's folded parameters: built once per
        set of parameter storages (keyed on data pointers), refilled by ONE kernel launch whenever the values change."""
        enc = self.encoder
        convs = [(enc[0], enc[1], True)]
        for child in list(enc.children())[4:-1]:
            for blk in child:
                convs.append((blk.conv1, blk.bn1, False))
                convs.append((blk.conv2, blk.bn2, False))
                if blk.downsample is not None:
                    convs.append((blk.downsample[0], blk.downsample[1], False))
        key = (self.precision, self.stem_mode) + tuple(t.data_ptr() for c, b, _ in convs
                                                       for t in (c.weight, b.weight, b.bias, b.running_mean, b.running_var))
        st = getattr(self, '_fold_cache', None)
        if st is not None and st['key'] == key:
            return st
        for c, b, _ in convs:
            for t in (c.weight, b.weight, b.bias, b.running_mean, b.running_var):
                if t.dtype != torch.float32 or not t.is_contiguous() or not t.is_cuda:
                    raise RuntimeError('ips_b200: encoder parameters must be contiguous fp32 CUDA tensors')
        entries = [self._conv_entry(c, b, stem) for c, b, stem in convs]
        import ctypes
        arr = (type(entries[0]['item']) * len(entries))(*[e['item'] for e in entries])
        host = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8)
        it = iter(entries)
        stem = next(it)
        blocks = []
        for child in list(enc.children())[4:-1]:
            for blk in child:
                b = dict(c1=next(it), c2=next(it), ds=None)
                if blk.downsample is not None:
                    b['ds'] = next(it)
                blocks.append(b)
        st = dict(key=key, items=host.to(entries[0]['w'].device), n=len(entries), stem=stem, blocks=blocks,
                  blocks_per_item=max(1, min(64, max(e['w'].numel() for e in entries) // 4096)))
        self._fold_cache = st
        return st

    def _build_plan(self):
        plan = {}
        ca = self.transf.crs_attn
        plan['U'] = ca.score_basis()
        plan['posU'] = None
        if self.use_pos:
            plan['posU'] = ops.logits(self.pos_enc[0].contiguous().float(), plan['U'])      # (N, HT)
        if self.precision == 'bf16x3':
            self._build_plan_x3(plan)
        elif self.is_image:
            st = self._fold_static()
            ops.fold_plan(st['items'], st['n'], st['blocks_per_item'])      # every layer's weights + BatchNorm in one launch
            plan['stem'], plan['blocks'] = st['stem'], st['blocks']
        else:
            lin, bn = self.encoder[1], self.encoder[2]
            scale, bshift = _fold_bn(bn)
            plan['p_scale'] = scale
            plan['p_shift'] = (lin.bias.detach().float() * scale + bshift).contiguous()
            w = lin.weight.detach().float().contiguous()
            plan['p_w'] = w.to(torch.bfloat16).contiguous() if self.precision == 'bf16' else w
            K, HT = w.shape[1], plan['U'].shape[1]
            plan['p_table'] = None                        # one-kernel path: features -> logits (csrc/umma_projector.cu)
            if (self.precision == 'bf16' and self.D in (256, 512) and HT <= 8 and K % 128 == 0 and K >= 256 and not self.use_pos
                    and not os.environ.get('IPSB_NO_FUSED_PROJECTOR')):
                plan['p_table'] = ops.projector_table(scale, plan['p_shift'], plan['p_w'], plan['U'])
        return plan

    # ------------------------------------------------------------------ bf16x3: near-fp32 on the tensor cores
    @staticmethod
    def _split_w(w):
        hi = w.to(torch.bfloat16)
        return hi.contiguous(), (w - hi.float()).to(torch.bfloat16).contiguous()

    def _build_plan_x3(self, plan):
        """Weights as hi / lo bf16 in the tcgen05 K-major layouts, arranged [wh | wl | wh] along the input channels (the
        partner of the activations' [hi | hi | lo] layout); BatchNorm(eval) as fp32 scale / shift."""
        def entry(conv, bn):
            w = conv.weight.detach().float()
            cout, cin, kh, kw = w.shape
            e = dict(cin=cin, cout=cout, kh=kh, kw=kw, stride=conv.stride[0], pad=conv.padding[0])
            e['scale'], e['shift'] = _fold_bn(bn)
            wt = w.permute(0, 2, 3, 1).contiguous()                          # (cout, kh, kw, cin)
            wh, wl = self._split_w(wt)
            e['w3'] = torch.cat([wh, wl, wh], dim=3).reshape(cout, kh * kw * 3 * cin).contiguous()
            return e
        if self.is_image:
            enc = self.encoder
            ps = self.patch_size
            if enc[0].out_channels != 64 or ps[0] % 2 or ps[1] % 2 or max(ps) > 480:
                raise NotImplementedError("precision 'bf16x3' needs even patch sizes up to 480 and the 64-channel stem")
            conv, bn = enc[0], enc[1]                                         # stem: (cout, 256), k = r*32 + (s+1)*4 + c, as hi and lo
            w = conv.weight.detach().float()
            cout, cin, kh, kw = w.shape
            st = dict(cout=cout)
            st['scale'], st['shift'] = _fold_bn(bn)
            st['zero'] = torch.zeros_like(st['shift'])
            w4 = torch.zeros((cout, 4, kh, kw), device=w.device)
            w4[:, :cin] = w
            wp = torch.zeros((cout, 8, 8, 4), device=w.device)
            wp[:, :kh, 1:kw + 1] = w4.permute(0, 2, 3, 1)
            st['wh'], st['wl'] = self._split_w(wp.reshape(cout, 256).contiguous())
            plan['stem'] = st
            blocks = []
            for child in list(enc.children())[4:-1]:
                for blk in child:
                    b = dict(c1=entry(blk.conv1, blk.bn1), c2=entry(blk.conv2, blk.bn2), ds=None)
                    if blk.downsample is not None:
                        b['ds'] = entry(blk.downsample[0], blk.downsample[1])
                    blocks.append(b)
            plan['blocks'] = blocks
        else:
            lin, bn = self.encoder[1], self.encoder[2]
            scale, bshift = _fold_bn(bn)
            plan['p_scale'] = scale
            plan['p_shift'] = (lin.bias.detach().float() * scale + bshift).contiguous()
            wh, wl = self._split_w(lin.weight.detach().float().contiguous())
            plan['p_w3'] = torch.cat([wh, wl, wh], dim=1).contiguous()        # (D, 3K)
            plan['p_table'] = None

    def _embed_x3(self, plan, flat, first_row, n_rows):
        """Eval-mode embeddings (rows, D) fp32 in the bf16x3 precision: every convolution / the projector GEMM as ONE
        fp32-output tensor-core launch over the tripled input channels, sums / residual / ReLU / pooling in fp32."""
        P = n_rows
        if not self.is_image:
            rows = flat[first_row:first_row + n_rows].contiguous().float()
            K = rows.shape[1]
            if K % 64 or self.D % 64:
                raise NotImplementedError("precision 'bf16x3' needs feature and embedding widths that are multiples of 64")
            a3 = ops.tri_from_f32(ops.layernorm_rows(rows, 1e-5))                                  # (P, 3K)
            y = ops.conv_f32out(a3.view(P, 1, 1, 3 * K), plan['p_w3'], plan['p_scale'], plan['p_shift'], self.D, 1, 1, 1, 0)
            return ops.sum_split_tri([y], relu=True, C=self.D, want_f32=True, want_tri=False)[1].view(P, self.D)
        _, C, H, W = flat.shape
        e = plan['stem']
        fh, fl = ops.stage_patches_padded_split(flat, P, C, H, W, first_row=first_row)
        parts = [ops.conv_f32out(f, w, e['scale'], sh, e['cout'], 7, 7, 2, 3, mode=3)
                 for f, w, sh in ((fh, e['wh'], e['shift']), (fh, e['wl'], e['zero']), (fl, e['wh'], e['zero']))]
        x3 = ops.sum3_maxpool_tri(parts, relu=True)                           # conv1 + bn1 + relu + maxpool -> (P, h, w, 192)
        del parts, fh, fl
        out = None

        def conv(x3, c):
            return ops.conv_f32out(x3, c['w3'], c['scale'], c['shift'], c['cout'], c['kh'], c['kw'], c['stride'], c['pad'])
        for bi, b in enumerate(plan['blocks']):
            last = bi == len(plan['blocks']) - 1
            idt_tri, idt_f32 = x3, None
            if b['ds'] is not None:
                idt_tri, idt_f32 = None, conv(x3, b['ds'])                    # downsample branch stays fp32
            y3, _ = ops.sum_split_tri([conv(x3, b['c1'])], relu=True, C=b['c1']['cout'])
            x3, out = ops.sum_split_tri([conv(y3, b['c2'])], relu=True, C=b['c2']['cout'], res_f32=idt_f32, res_tri=idt_tri,
                                        want_f32=last, want_tri=not last)
        return ops.avgpool(out, ops.F32)

    def invalidate_plan(self):
        """Drop the folded-parameter plan.  `ips()` notices in-place parameter updates through the tensors' version
        counters; updates that bypass them -- a CUDA-graph replay of the optimizer step (`GraphedTrainStep`), writes
        through `.data` -- must call this, or the next `ips()` would select with stale weights."""
        self._plan = None
        self._plan_key = None

    def _get_plan(self):
        key = self._state_versions()
        if self._plan is None or key != self._plan_key:
            self._plan = self._build_plan()
            self._plan_key = key
        return self._plan

    # ------------------------------------------------------------------ no-grad embedding (CUDA kernels)
    def _conv(self, x, e, res=None, relu=True):
        f = ops.conv_bf16 if self.precision == 'bf16' else ops.conv_f32
        if self.precision == 'bf16':
            return f(x, e['w'], e['scale'], e['shift'], res, e['cout'], e['kh'], e['kw'], e['stride'], e['pad'], relu, e['mode'])
        return f(x, e['w'], e['scale'], e['shift'], res, e['cout'], e['kh'], e['kw'], e['stride'], e['pad'], relu)

    @_on_device
    @torch.no_grad()
    def embed(self, flat, row_idx=None, first_row=0, n_rows=None):
        """Eval-mode patch embeddings (rows, D) fp32 of `flat` = (rows, C, ph, pw) or (rows, F) on the GPU.
        Equivalent to `encoder.eval()(x).view(rows, -1)` of the reference (ips_net.py:209,227)."""
        plan = self._get_plan()
        dt = ops.BF16 if self.precision == 'bf16' else ops.F32
        if n_rows is None:
            n_rows = flat.shape[0] - first_row if row_idx is None else row_idx.numel()
        if self.precision == 'bf16x3':
            if row_idx is not None:
                flat, first_row = flat[row_idx].contiguous(), 0
            return self._embed_x3(plan, flat.contiguous(), first_row, n_rows)
        if self.is_image:
            _, C, H, W = flat.shape
            if plan['stem']['mode'] >= 3:
                return self._embed_pf(plan, flat, row_idx, first_row, n_rows, C, H, W)
            x = ops.stage_patches(flat, n_rows, C, H, W, dt, row_idx=row_idx, first_row=first_row)
            x = self._conv(x, plan['stem'])
            x = ops.maxpool3x3s2(x, dt)
            for b in plan['blocks']:
                idt = x if b['ds'] is None else self._conv(x, b['ds'], relu=False)
                y = self._conv(x, b['c1'])
                x = self._conv(y, b['c2'], res=idt)
            return ops.avgpool(x, dt)
        rows = flat[first_row:first_row + n_rows] if row_idx is None else flat[row_idx]
        if self.precision == 'bf16':
            a = ops.rows_to_bf16(rows.contiguous(), layernorm=True, eps=1e-5)
            return ops.linear_bf16(a, plan['p_w'], plan['p_scale'], plan['p_shift'], relu=True)
        a = ops.layernorm_rows(rows.contiguous().float(), 1e-5)
        return ops.linear_f32(a, plan['p_w'], plan['p_scale'], plan['p_shift'], relu=True)

    def _embed_pf(self, plan, flat, row_idx, first_row, n_rows, C, H, W):
        """bf16 encoder on padded-flat activations (ips_b200/csrc/pf.cuh), one library call per layer;
        the same kernel sequence the native executor issues."""
        P = n_rows
        if plan['stem']['mode'] == 4:
            e = plan['stem']
            x = ops.stage_patches_s2d(flat, P, C, H, W, row_idx=row_idx, first_row=first_row)
            h, w = H // 2, W // 2
            if e['cout'] == 64 and not os.environ.get('IPSB_STEM_UNFUSED'):
                x = ops.stem_pool_s2d(x, e['w'], e['scale'], e['shift'], P, H, W)          # stem output stays on chip
            else:
                x = ops.conv_stem_s2d(x, e['w'], e['scale'], e['shift'], P, H, W, e['cout'])   # wide row order
                x = ops.maxpool3x3s2_pf_strided(x, P, h, w, e['cout'], w + 3, (h + 3) * (w + 3))
        else:
            x = ops.stage_patches_padded(flat, P, C, H, W, row_idx=row_idx, first_row=first_row)
            x = self._conv(x, plan['stem'])                              # dense (P, H/2, W/2, 64)
            h, w = x.shape[1], x.shape[2]
            x = ops.maxpool3x3s2_pf(x)
        h, w = (h - 1) // 2 + 1, (w - 1) // 2 + 1

        def conv(x, e, h, w, res=None, relu=True):
            return ops.conv_bf16_pf(x, e['w'], e['scale'], e['shift'], res, P, h, w, e['cout'], e['kh'], e['kw'],
                                    e['stride'], e['pad'], relu, True, True)
        for b in plan['blocks']:
            s = b['c1']['stride']
            ho, wo = (h + 2 - 3) // s + 1, (w + 2 - 3) // s + 1
            idt = x if b['ds'] is None else conv(x, b['ds'], h, w, relu=False)
            y = conv(x, b['c1'], h, w)
            x = conv(y, b['c2'], ho, wo, res=idt)
            h, w = ho, wo
        return ops.avgpool_pf(x, P, h, w)

    def _auto_chunk(self, patch_shape, rows=None):
        """Patches per encoder launch.  Large chunks waste less of each persistent kernel's last wave of tiles; with
        `lanes` chunks in flight the rows are split evenly over the lanes, capped by the activation workspace
        (~1.3 GB per 1024 patches of 100x100)."""
        if self.chunk_patches:
            return self.chunk_patches
        if not self.is_image:
            return 16384
        px = patch_shape[-1] * patch_shape[-2]
        cap = max(32, min(8192, (2048 * 10000) // max(px, 1)))
        if rows is None:
            return cap
        per_lane = -(-rows // max(1, self.lanes))
        n = -(-per_lane // cap)                                        # chunks per lane
        return max(32, -(-(-(-per_lane // n)) // 8) * 8)

    @_on_device
    @torch.no_grad()
    def patch_logits(self, patches, pos_offset=0):
        """(B,N,...) -> (B,N,H*T) fp32 logit table on `self.device`, original patch order.
        Host-resident input (lazy loading, conf.eager=False) is streamed chunk by chunk.
        `pos_offset`: index of patches[:, 0] in the full sequence (sequence-sharded runs)."""
        plan = self._get_plan()
        B, N = patches.shape[:2]
        rows = B * N
        flat = patches.reshape(rows, *patches.shape[2:])
        HT = plan['U'].shape[1]
        chunk = self._auto_chunk(patches.shape, rows)
        if self.is_image and flat.is_cuda and self.executor == 'native' and pos_offset == 0 and self.precision != 'bf16x3':
            if 'desc' not in plan:
                plan['desc'] = ops.make_resnet_desc(plan, ops.BF16 if self.precision == 'bf16' else ops.F32, self.D, HT)
            z, _ = ops.resnet_logits(plan['desc'], flat.contiguous(), N, chunk, self._ws_cache, lanes=self.lanes)
            return z.view(B, N, HT)
        if not self.is_image and flat.is_cuda and plan.get('p_table') is not None and flat.dtype in (torch.float32, torch.bfloat16):
            return ops.projector_logits(flat.contiguous(), plan['p_w'], plan['p_table'], HT, 1e-5).view(B, N, HT)
        z = torch.empty((rows, HT), dtype=torch.float32, device=self.device)
        pos_idx = None
        if self.use_pos:
            pos_idx = (torch.arange(rows, device=self.device) % N + pos_offset).contiguous()
        for lo in range(0, rows, chunk):
            n = min(chunk, rows - lo)
            if flat.is_cuda:
                emb = self.embed(flat, first_row=lo, n_rows=n)
            else:
                emb = self.embed(flat[lo:lo + n].to(self.device, non_blocking=True).contiguous())
            z[lo:lo + n] = ops.logits(emb, plan['U'], plan['posU'], None if pos_idx is None else pos_idx[lo:lo + n].contiguous())
        return z.view(B, N, HT)

    @torch.no_grad()
    def _stream_in(self, patches):
        """Host-resident patches (lazy loading, `conf.eager: False`): copy them to the device chunk by chunk on a
        side stream while the encoder already works on the chunks that have arrived.  (The reference moves each chunk
        with `.to(device)` inside its loop, ips_net.py:206,223.)  Returns (device copy or None, logit table).

        Memory: IPS's guarantee is O(M + I) device memory whatever N.  A whole-tensor device copy is
        kept only when it fits `lazy_resident_bytes` (the final gather then runs on the device); larger inputs stream
        through a ring of three chunk buffers, only the logit table (N * H*T * 4 bytes) survives, and the M winners are
        gathered from the host tensor afterwards like the reference does (:244-247)."""
        plan = self._get_plan()
        B, N = patches.shape[:2]
        rows = B * N
        flat_h = patches.reshape(rows, *patches.shape[2:])
        keep_bf16 = (not self.is_image) and flat_h.dtype == torch.bfloat16 and self.precision == 'bf16'   # bf16 feature bags
        if flat_h.dtype != torch.float32 and not keep_bf16:
            flat_h = flat_h.float()
        HT = plan['U'].shape[1]
        # smaller chunks than the resident path: the encoder starts as soon as the first chunk has arrived and only the
        # last chunk's compute is exposed after the copy (traffic, 369 MB over PCIe: raw copy 6.7 ms, ips() 8.0 ms)
        chunk = self.chunk_patches or (max(32, min(4096, (512 * 10000) // max(patches.shape[-1] * patches.shape[-2], 1)))
                                       if self.is_image else 16384)
        chunk = min(chunk, rows)
        main = torch.cuda.current_stream(self.device)
        if getattr(self, '_copy_stream', None) is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
        cs = self._copy_stream
        resident = flat_h.numel() * flat_h.element_size() <= self.lazy_resident_bytes
        n_chunks = -(-rows // chunk)
        ring = 1 if resident else min(3, n_chunks)
        dev = torch.empty((rows if resident else ring * chunk, *flat_h.shape[1:]), dtype=flat_h.dtype, device=self.device)
        dev.record_stream(cs)
        cs.wait_stream(main)
        z = torch.empty((rows, HT), dtype=torch.float32, device=self.device)
        native = self.is_image and self.executor == 'native' and self.precision != 'bf16x3'
        if native and 'desc' not in plan:
            plan['desc'] = ops.make_resnet_desc(plan, ops.BF16 if self.precision == 'bf16' else ops.F32, self.D, HT)
        pos_idx = (torch.arange(rows, device=self.device) % N).contiguous() if self.use_pos else None
        arrived, consumed = [], []
        # chunk boundaries; with a whole-tensor device copy and the automatic chunk size the LAST chunk is split into
        # 1/2 + 1/4 + 1/4: only the last piece's encoder time is exposed after the final copy
        edges = list(range(0, rows, chunk)) + [rows]
        if resident and not self.chunk_patches and self.is_image and edges[-1] - edges[-2] >= 128:
            a, b = edges[-2], edges[-1]
            q = (b - a) // 4
            edges = edges[:-1] + [a + 2 * q, a + 3 * q, b]
        n_chunks = len(edges) - 1

        def issue_copy(ci):
            lo, n = edges[ci], edges[ci + 1] - edges[ci]
            slot = lo if resident else (ci % ring) * chunk
            with torch.cuda.stream(cs):
                if not resident and ci >= ring:
                    cs.wait_event(consumed[ci - ring])        # the encoder is done with this slot's previous chunk
                dev[slot:slot + n].copy_(flat_h[lo:lo + n], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(cs)
                arrived.append(ev)

        ahead = n_chunks if resident else ring
        for ci in range(min(ahead, n_chunks)):
            issue_copy(ci)
        for ci in range(n_chunks):
            lo, n = edges[ci], edges[ci + 1] - edges[ci]
            slot = lo if resident else (ci % ring) * chunk
            main.wait_event(arrived[ci])
            if native:
                ops.resnet_logits(plan['desc'], dev, N, chunk, self._ws_cache, first_row=lo, n_rows=n, z=z, src_first_row=lo - slot)
            elif not self.is_image and plan.get('p_table') is not None:
                z[lo:lo + n] = ops.projector_logits(dev[slot:slot + n], plan['p_w'], plan['p_table'], HT, 1e-5)
            else:
                emb = self.embed(dev, first_row=slot, n_rows=n)
                z[lo:lo + n] = ops.logits(emb, plan['U'], plan['posU'], None if pos_idx is None else pos_idx[lo:lo + n].contiguous())
            if not resident:
                ev = torch.cuda.Event()
                ev.record(main)
                consumed.append(ev)
                if ci + ring < n_chunks:
                    issue_copy(ci + ring)
        return (dev.view(patches.shape) if resident else None), z.view(B, N, HT)

    def _draw_scan_order(self, B, N, data_device, device=None):
        """Scan order of one call as a DEVICE tensor: (perm (1|B, N) int64 on the device or None, per_instance).  Same RNG
        calls as the reference (utils.scan_order).  The host-drawn 'batch' order is written straight into a ring of pinned
        buffers and copied asynchronously: a copy from pageable memory would first synchronise the stream, i.e. serialise
        the host's work for call k+1 (the draw and ~50 kernel launches) with the GPU work of call k: traffic 3.07 -> 2.90 ms
        per call, mnist 2.50 -> 2.42 ms.  Long sequences (N > 16 384) keep the plain copy: there the host's randperm itself
        (0.4 ms for 50 000 patches) is longer than the GPU work and the extra host calls of the ring only add to it
        (measured 0.48 -> 0.56 ms); `scan_order_rng: device` is the answer for those."""
        device = self.device if device is None else device
        if (self.shuffle and self.scan_order_rng == 'reference' and self.shuffle_style == 'batch'
                and torch.device(device).type == 'cuda' and N <= 16384 and os.environ.get('IPS_B200_PINNED_ORDER', '1') != '0'):
            ring = getattr(self, '_perm_ring', None)
            if ring is None or ring['n'] < N:
                ring = self._perm_ring = {'n': N, 'i': 0, 'slots': [[torch.empty(N, dtype=torch.int64).pin_memory(), None]
                                                                   for _ in range(4)]}
            slot = ring['slots'][ring['i'] % 4]
            ring['i'] += 1
            if slot[1] is not None:
                slot[1].synchronize()                                # its previous copy has left the buffer (long ago)
            buf = slot[0][:N]
            torch.randperm(N, out=buf)                               # global CPU generator, utils.py:38
            perm = torch.empty((1, N), dtype=torch.int64, device=device)
            perm[0].copy_(buf, non_blocking=True)
            slot[1] = torch.cuda.Event()
            slot[1].record(torch.cuda.current_stream(device))
            return perm, False
        perm, per_inst = scan_order(self.shuffle, self.shuffle_style, B, N, data_device, self.scan_order_rng, device)
        if perm is not None:
            perm = perm.to(device).contiguous()
        return perm, per_inst

    # ------------------------------------------------------------------ reference API
    def do_shuffle(self, patches, pos_enc):
        """Kept for API compatibility (ips_net.py:118-134): returns permuted copies like the
        reference; `ips` itself never copies, it passes the scan order to the kernels."""
        B, N = patches.shape[:2]
        perm, per_inst = scan_order(True, self.shuffle_style, B, N, patches.device)
        if perm is None:
            return patches, pos_enc
        perm = perm.to(patches.device)
        idx = perm.expand(B, -1)
        patches = torch.stack([patches[b, idx[b]] for b in range(B)])
        if torch.is_tensor(pos_enc):
            pos_enc = torch.stack([pos_enc[b, idx[b].to(pos_enc.device)] for b in range(B)])
        return patches, pos_enc

    @_on_device
    @torch.no_grad()
    def score_and_select(self, emb, emb_pos, M, idx):
        """Scores embeddings and keeps the top M (ips_net.py:136-155); stable tie-break."""
        to_score = emb_pos if torch.is_tensor(emb_pos) else emb
        scores = self.transf.get_scores(to_score)
        _, top = ops.topm_stable(scores.contiguous(), M)
        D = emb.shape[2]
        mem_emb = torch.gather(emb, 1, top.unsqueeze(-1).expand(-1, -1, D))
        mem_idx = torch.gather(idx, 1, top)
        return mem_emb, mem_idx

    def get_preds(self, embeddings):
        preds = {}
        for task in self.tasks.values():
            preds[task['name']] = self.output_layers[task['name']](embeddings[:, task['id']])
        return preds

    @_on_device
    @torch.no_grad()
    def ips(self, patches, out=None, row_offset=0):
        """Iterative Patch Selection (ips_net.py:169-262): returns (mem_patch, mem_pos).

        `out=(mem_patch_buf, mem_pos_buf)`, `row_offset=n_prep`: write the winners straight into rows
        [n_prep, n_prep + B) of the train-step buffers of `init_batch` instead of allocating and copying
        (`fill_batch`, training/iterative.py:31-50).  The returned tensors are views of the buffers."""
        M, I, D = self.M, self.I, self.D
        device = self.device
        B, N = patches.shape[:2]
        if M >= N:                                               # shortcut, :185-188
            pos_enc = self.pos_enc.expand(B, -1, -1) if self.use_pos else None
            self.last_mem_idx = None
            mem_patch = patches.to(device)
            if out is not None:
                out[0][row_offset:row_offset + B, :N] = mem_patch
                mem_patch = out[0][row_offset:row_offset + B, :N]
                if self.use_pos:
                    out[1][row_offset:row_offset + B, :N] = pos_enc
                    pos_enc = out[1][row_offset:row_offset + B, :N]
            return mem_patch, pos_enc
        if torch.device(device).type != 'cuda':
            raise RuntimeError('ips_b200.IPSNet.ips needs a CUDA device: there is no CPU implementation')

        perm, per_inst = self._draw_scan_order(B, N, patches.device, device)

        ca = self.transf.crs_attn
        if not patches.is_cuda:                                   # lazy loading: overlap H2D with the encoder
            dev_copy, z = self._stream_in(patches)
            if dev_copy is not None:
                patches = dev_copy
        else:
            z = None
            if not self.is_image and patches.dtype in (torch.float32, torch.bfloat16) and self.streamed_select:
                plan = self._get_plan()
                if plan.get('p_table') is not None and ops.streamed_select_ok(B, N, ca.H * ca.n_token, M, I):
                    # projector and selection loop side by side: the loop consumes the scan-ordered logits tile by tile
                    _, mem_src, _, _, self.last_stream_status = ops.projector_select(
                        patches.reshape(B * N, -1).contiguous(), plan['p_w'], plan['p_table'], B, N, perm, per_inst,
                        ca.H, ca.n_token, M, I)
                    z = False
            if z is None:
                z = self.patch_logits(patches)                    # encode + project every patch once
        if z is not False:
            _, mem_src, _ = ops.select_loop(z, perm, per_inst, ca.H, ca.n_token, M, I)
        self.last_mem_idx = mem_src

        o_patch, o_pos = self._out_views(out, row_offset, B, M)
        if patches.is_cuda and patches.dtype == torch.bfloat16:   # bf16 feature bag: the winners go to the train step as fp32
            mem_patch = ops.gather_rows(patches.contiguous(), mem_src, N).float()
            if o_patch is not None:
                o_patch.copy_(mem_patch)
                mem_patch = o_patch
        elif patches.is_cuda:
            mem_patch = ops.gather_rows(patches.contiguous(), mem_src, N, out=o_patch)
        else:                                                    # lazy loading: gather on the host, :244-247
            host_idx = mem_src.cpu()
            mem_patch = torch.stack([patches[b, host_idx[b]] for b in range(B)]).to(device).float()
            if o_patch is not None:
                o_patch.copy_(mem_patch)
                mem_patch = o_patch
        mem_pos = ops.gather_rows(self.pos_enc[0].contiguous(), mem_src, 0, out=o_pos) if self.use_pos else None
        return mem_patch, mem_pos

    def _out_views(self, out, row_offset, B, M):
        """Rows [row_offset, row_offset + B) of the caller's (B_total, M, ...) buffers, or (None, None)."""
        if out is None:
            return None, None
        buf_patch, buf_pos = out
        if buf_patch.shape[1] != M or row_offset + B > buf_patch.shape[0]:
            raise ValueError('ips_b200: out buffer of shape %s cannot take %d x %d winners at row %d'
                             % (tuple(buf_patch.shape), B, M, row_offset))
        o_pos = buf_pos[row_offset:row_offset + B] if (self.use_pos and buf_pos is not None) else None
        return buf_patch[row_offset:row_offset + B], o_pos

    @_on_device
    @torch.no_grad()
    def ips_image(self, images, patch_size=None, patch_stride=None, out=None, row_offset=0):
        """`ips` on whole images (B, C, Himg, Wimg): the patch grid the reference's data loaders cut on the CPU
        (`unfold(1, ph, sh).unfold(2, pw, sw)`, mnist_dataset.py:47-53, traffic_dataset.py:337-343; patch n = grid row
        n // n_cols, column n % n_cols) is read straight out of the images by the staging kernel and by the final
        gather.  Same result as `ips(patchify(images))`, same RNG consumption."""
        if not self.is_image:
            raise ValueError('ips_image needs an image configuration')
        ph, pw = patch_size or self.patch_size
        sh, sw = patch_stride or (ph, pw)
        if not images.is_cuda:
            images = images.to(self.device, non_blocking=True)
        images = images.float().contiguous()
        B, C = images.shape[:2]
        geo, N = ops.image_geo(images, (ph, pw), (sh, sw))
        M, I = self.M, self.I
        plan = self._get_plan()
        direct = (self.precision == 'bf16' and self.executor == 'native' and plan['stem']['mode'] == 4
                  and plan['stem']['cout'] == 64 and M < N)
        if not direct:                                            # other modes: patchify on the device, then the usual path
            return self.ips(ops.gather_patches_image(images, geo, None, (ph, pw)), out=out, row_offset=row_offset)
        perm, per_inst = self._draw_scan_order(B, N, images.device)
        ca = self.transf.crs_attn
        HT = plan['U'].shape[1]
        if 'desc' not in plan:
            plan['desc'] = ops.make_resnet_desc(plan, ops.BF16, self.D, HT)
        chunk = self._auto_chunk((ph, pw), B * N)
        z, _ = ops.resnet_logits(plan['desc'], images, N, chunk, self._ws_cache, lanes=self.lanes, geo=geo, patch_size=(ph, pw))
        _, mem_src, _ = ops.select_loop(z.view(B, N, HT), perm, per_inst, ca.H, ca.n_token, M, I)
        self.last_mem_idx = mem_src
        o_patch, o_pos = self._out_views(out, row_offset, B, M)
        mem_patch = ops.gather_patches_image(images, geo, mem_src, (ph, pw), out=o_patch)
        mem_pos = ops.gather_rows(self.pos_enc[0].contiguous(), mem_src, 0, out=o_pos) if self.use_pos else None
        return mem_patch, mem_pos

    @torch.no_grad()
    def _forward_inference(self, mem_patch, mem_pos):
        """`forward` in eval mode under no-grad (training/iterative.py:193-231 `evaluate`): every operator is a
        library kernel -- encoder as in `ips`, then projections, cross-attention, LayerNorm, MLP, heads."""
        ca, mlp = self.transf.crs_attn, self.transf.mlp
        B, M = mem_patch.shape[:2]
        f = lambda t: t.detach().float().contiguous()
        emb = self.embed(mem_patch.reshape(B * M, *mem_patch.shape[2:]).contiguous())          # (B*M, D) fp32
        if torch.is_tensor(mem_pos):
            emb = ops.add(emb, mem_pos.reshape(B * M, -1).float().contiguous())
        k = ops.linear_f32(emb, f(ca.k_w.weight)).view(B, M, -1)
        v = ops.linear_f32(emb, f(ca.v_w.weight)).view(B, M, -1)
        q_tok = f(ca.q[0])                                                                      # (T, D)
        inv_t = torch.full((ca.H * ca.D_k,), 1.0 / ca.attention.temperature, device=emb.device)
        q = ops.linear_f32(q_tok, f(ca.q_w.weight), scale=inv_t)                                # q / sqrt(D_k)
        att = ops.cross_attention(q, k, v, ca.H, ca.D_k, ca.D_v)                                # (B, T, H*Dv)
        T = q_tok.shape[0]
        o = ops.linear_f32(att.view(B * T, -1), f(ca.fc.weight))
        o = ops.residual_layernorm(o, q_tok, f(ca.layer_norm.weight), f(ca.layer_norm.bias), ca.layer_norm.eps)
        hdn = ops.linear_f32(o, f(mlp.w_1.weight), shift=f(mlp.w_1.bias), relu=True)
        h2 = ops.linear_f32(hdn, f(mlp.w_2.weight), shift=f(mlp.w_2.bias))
        tok = ops.residual_layernorm(h2, o, f(mlp.layer_norm.weight), f(mlp.layer_norm.bias), mlp.layer_norm.eps).view(B, T, -1)
        preds = {}
        for task in self.tasks.values():
            lin = self.output_layers[task['name']][0]
            zl = ops.linear_f32(tok[:, task['id']].contiguous(), f(lin.weight), shift=f(lin.bias))
            preds[task['name']] = ops.head_activation(zl, task['act_fn'])
        return preds

    @_on_device
    def forward(self, mem_patch, mem_pos=None, tokens_only=False):
        """Encode + aggregate the selected patches (ips_net.py:264-283).  Inference (eval mode, no grad) runs on
        the library's kernels; the grad-mode train step uses PyTorch autograd on the same parameters."""
        if not self.training and not torch.is_grad_enabled() and mem_patch.is_cuda and not tokens_only:
            return self._forward_inference(mem_patch, mem_pos)
        shape = mem_patch.shape
        B, M = shape[:2]
        if not self.is_image and mem_patch.is_cuda and self.training:
            # feature projector on the library's kernels, forward and backward: LayerNorm -> Linear -> BatchNorm1d
            # (batch statistics, running-stat update) + ReLU
            ln, lin, bn = self.encoder[0], self.encoder[1], self.encoder[2]
            h = LayerNormFn.apply(mem_patch.reshape(B * M, -1), None, None, ln.eps)
            h = lin(h)
            mem_emb = BatchNormTrainFn.apply(h, bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.momentum, bn.eps, True,
                                             self.sync_bn)
            with torch.no_grad():
                bn.num_batches_tracked += 1
            mem_emb = mem_emb.view(B, M, -1)
        elif (self.is_image and mem_patch.is_cuda and self.training and self.precision == 'bf16'
              and self.train_encoder == 'native' and self.encoder[0].out_channels == 64):
            # conv encoder on the library's kernels, forward and backward (ips_b200/autograd.py::conv_encoder_train)
            mem_emb = conv_encoder_train(self.encoder, mem_patch.reshape(-1, *shape[2:]), self.sync_bn).view(B, M, -1)
        else:
            mem_emb = self.encoder(mem_patch.reshape(-1, *shape[2:])).view(B, M, -1)
        if torch.is_tensor(mem_pos):
            mem_emb = mem_emb + mem_pos
        tok = self.transf(mem_emb)
        return tok if tokens_only else self.get_preds(tok)

    @_on_device
    def loss(self, mem_patch, mem_pos, labels, eps=1e-6):
        """`compute_loss(net, mem_patch, mem_pos, criterions, labels, conf)` of training/iterative.py:65-100 with the
        head activations, the losses and their gradients fused into one kernel per task: mean over
        tasks of NLLLoss(log(softmax + eps)) / BCELoss(sigmoid).  Returns the scalar loss (on the device, no sync)."""
        tok = self.forward(mem_patch, mem_pos, tokens_only=True)
        if not tok.is_cuda:
            raise RuntimeError('ips_b200.IPSNet.loss needs a CUDA device')
        total = 0
        for task in self.tasks.values():
            z = self.output_layers[task['name']][0](tok[:, task['id']])
            total = total + HeadLossFn.apply(z, labels[task['name']], task['act_fn'], eps)
        return total / len(self.tasks)

"""Host-side helpers mirroring utils/utils.py:10-12,33-58 of the reference."""
import torch


class Struct:
    """Attribute bag built from the YAML config (utils/utils.py:10-12)."""

    def __init__(self, **entries):
        self.__dict__.update(entries)


def scan_order(shuffle, style, B, N, data_device, rng='reference', device=None, generator=None):
    """The reference shuffles by materialising permuted copies of the patches
    (utils/utils.py:39,56).  Here only the permutation is drawn -- with the SAME
    RNG calls, so seeds reproduce the reference's order -- and handed to the
    kernels as the scan order.  Returns (perm int64 (1|B, N) on CPU/`data_device`, per_instance).

    rng='device' (conf.scan_order_rng): the order is drawn on `device` instead: two int64 from the CUDA generator
    (optional `generator`) key a bijection of [0, N) evaluated by one kernel (ops.keyed_scan_order; one key per instance
    for 'instance').  A shuffle, NOT the reference's random stream.  The host draw costs 0.4 ms for 50 000 patches and
    1.6 ms for 200 000 -- more than the whole selection on the GPU -- and cannot be captured in a CUDA graph."""
    if not shuffle:
        return None, False
    if rng == 'device':
        if style not in ('batch', 'instance'):
            return None, False
        from . import ops
        key = torch.randint(-2 ** 63, 2 ** 63 - 1, (2,), dtype=torch.int64, device=device, generator=generator)
        return ops.keyed_scan_order(key, B if style == 'instance' else 1, N), style == 'instance'
    if style == 'batch':
        return torch.randperm(N).unsqueeze(0), False            # global CPU generator, utils.py:38
    if style == 'instance':
        return torch.rand((B, N), device=data_device).argsort(1), True   # utils.py:48
    return None, False                                          # unknown style: reference does nothing

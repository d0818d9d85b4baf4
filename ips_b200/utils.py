"""Host-side helpers mirroring utils/utils.py:10-12,33-58 of the reference."""
import torch


class Struct:
    """Attribute bag built from the YAML config (utils/utils.py:10-12)."""

    def __init__(self, **entries):
        self.__dict__.update(entries)


def scan_order(shuffle, style, B, N, data_device):
    """The reference shuffles by materialising permuted copies of the patches
    (utils/utils.py:39,56).  Here only the permutation is drawn -- with the SAME
    RNG calls, so seeds reproduce the reference's order -- and handed to the
    kernels as the scan order.  Returns (perm int64 (1|B, N) on CPU/`data_device`, per_instance)."""
    if not shuffle:
        return None, False
    if style == 'batch':
        return torch.randperm(N).unsqueeze(0), False            # global CPU generator, utils.py:38
    if style == 'instance':
        return torch.rand((B, N), device=data_device).argsort(1), True   # utils.py:48
    return None, False                                          # unknown style: reference does nothing

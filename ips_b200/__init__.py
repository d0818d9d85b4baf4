"""ips_b200 -- B200-native (sm_100a) implementation of the IPS selection hot path.

Public surface mirrors the reference: ``IPSNet(device, conf)`` with ``ips`` /
``forward``, ``Transformer``, ``Struct``.  All compute of ``ips`` goes through the
C-ABI library ``ips_b200/csrc/libips_b200.so`` (include/ips_b200.h); there is no
CPU or PyTorch fallback for it.
"""
from .utils import Struct  # noqa: F401
from .transformer import Transformer, pos_enc_1d  # noqa: F401
from .ips_net import IPSNet  # noqa: F401

__all__ = ['IPSNet', 'Transformer', 'Struct', 'pos_enc_1d']

"""Import and drive the UNMODIFIED reference from baseline/_ref (see install_ref.py).  BASELINE / TEST INFRASTRUCTURE.

Used by `bench.py --impl reference` (the reference's own CPU path, timed), by the `gpu_library_baseline` leg of the GPU
arm (the same unmodified code on the B200 through torch's cuDNN / cuBLAS) and by the call-site tests, which run the
reference's `training/iterative.py` with `ips_b200.IPSNet` in place of the reference class.  Nothing under `ips_b200/`
imports this file.
"""
import importlib
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import install_ref  # noqa: E402


class ReferenceUnavailable(RuntimeError):
    pass


def load_reference():
    """(IPSNet, Struct, iterative module) of the unmodified reference.  baseline/_ref first (travels to the GPU box),
    else /root/reference where it exists (build container)."""
    path = install_ref.ref_path()
    if path is None and os.path.isdir(os.path.join(install_ref.SRC, 'architecture')):
        path = install_ref.SRC
    if path is None:
        raise ReferenceUnavailable('baseline/_ref is empty and /root/reference does not exist: run baseline/install_ref.py '
                                   'in the build container')
    if path not in sys.path:
        sys.path.insert(0, path)
    ips_net = importlib.import_module('architecture.ips_net')
    utils = importlib.import_module('utils.utils')
    iterative = importlib.import_module('training.iterative')
    src = os.path.abspath(ips_net.__file__)
    if not src.startswith(os.path.abspath(path)):
        raise ReferenceUnavailable(f'`architecture` resolved to {src}, not to the reference under {path}')
    return ips_net.IPSNet, utils.Struct, iterative


def reference_conf(preset, **over):
    """The reference's own YAML (baseline/_ref/config) as its Struct, `pretrained` off (no network for ImageNet
    weights; north_star: random init) plus overrides -- exactly how main.py:25-28 builds `conf`."""
    import yaml
    _, Struct, _ = load_reference()
    path = install_ref.ref_path() or install_ref.SRC
    with open(os.path.join(path, 'config', f'{preset}_config.yml')) as f:
        c = yaml.safe_load(f)
    c['pretrained'] = False
    c.update(over)
    return Struct(**c)


def build_reference_net(conf, device, state_dict=None, seed=0):
    """Reference IPSNet on `device`, in train() mode like training/iterative.py:109; optional weights."""
    import torch
    IPSNet, _, _ = load_reference()
    torch.manual_seed(seed)
    net = IPSNet(device, conf).to(device)
    if state_dict is not None:
        net.load_state_dict(state_dict, strict=True)
    net.train()
    return net

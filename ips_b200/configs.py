"""Load the shipped configs (config/*.yml; same keys as the reference's YAML files)."""
import os

import yaml

from .utils import Struct

_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'config')


def load_config(name, **overrides):
    """name in {'mnist', 'traffic', 'camelyon'} -> Struct, as main.py:25-28 builds it."""
    with open(os.path.join(_DIR, f'{name}_config.yml')) as f:
        c = yaml.safe_load(f)
    c.update(overrides)
    return Struct(**c)

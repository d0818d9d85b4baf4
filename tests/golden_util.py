"""Rebuild a golden case's config / weights / inputs from its stored seeds."""
import json
import os

import numpy as np

import ips_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
CASE_NAMES = sorted(f[:-4] for f in os.listdir(GOLDEN) if f.endswith('.npz'))


def load_case(name):
    z = np.load(os.path.join(GOLDEN, name + '.npz'), allow_pickle=False)
    meta = json.loads(str(z['meta']))
    conf = O.preset(meta['preset'], **meta['over'])
    conf.B = meta['B']
    sd = O.make_state(conf, meta['state_seed'], q_gain=meta['q_gain'])
    patches = O.make_patches(conf, meta['B'], meta['N'], meta['data_seed'], sparse=meta['sparse'])
    return z, meta, conf, sd, patches

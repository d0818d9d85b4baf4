"""Flat feature-bag files for pre-extracted patch features.

The reference stores CAMELYON16 features as one gzip-9 HDF5 dataset per slide (`data/camelyon/extract_feat.py:168-171`)
and reads a whole slide with `slide['img'][:]` per item (`data/camelyon/camelyon_dataset.py:129-143`): every access
decompresses on a CPU worker.  This module keeps the same dataset interface over ONE flat file that is memory-mapped:

    [8 B magic 'IPSBAG01'][8 B little-endian offset of the index][slide 0 rows][slide 1 rows] ... [index: JSON, utf-8]

Rows are `n_features` values of fp32 or bf16 (bf16 halves the bytes over PCIe and in HBM; `IPSNet.ips` accepts bf16
feature bags directly -- LayerNorm statistics are taken in fp32 on the exactly-upcast values).  Slides start at
4096-byte boundaries so a slide can be handed to direct I/O unchanged.
"""
import json
import os

import numpy as np
import torch

MAGIC = b'IPSBAG01'
_ALIGN = 4096


def _to_u16(t):
    """bf16 tensor -> uint16 numpy view of the same bits."""
    return t.contiguous().view(torch.int16).numpy().view(np.uint16)


class FeatureBagWriter:
    """with FeatureBagWriter(path, n_features, dtype='bf16') as w: w.add(name, features (N, F) tensor, label)"""

    def __init__(self, path, n_features, dtype='bf16'):
        if dtype not in ('bf16', 'fp32'):
            raise ValueError("dtype must be 'bf16' or 'fp32'")
        self.path, self.F, self.dtype = path, int(n_features), dtype
        self.f = open(path, 'wb')
        self.f.write(MAGIC + (0).to_bytes(8, 'little'))
        self.index = {'n_features': self.F, 'dtype': dtype, 'slides': []}

    def add(self, name, features, label=None, pos=None):
        features = torch.as_tensor(features)
        if features.dim() != 2 or features.shape[1] != self.F:
            raise ValueError('features must be (N, %d)' % self.F)
        pad = (-self.f.tell()) % _ALIGN
        self.f.write(b'\0' * pad)
        off = self.f.tell()
        if self.dtype == 'bf16':
            self.f.write(_to_u16(features.to(torch.bfloat16)).tobytes())
        else:
            self.f.write(features.float().contiguous().numpy().tobytes())
        entry = {'name': str(name), 'offset': off, 'n_patches': int(features.shape[0])}
        if label is not None:
            entry['label'] = int(label)
        if pos is not None:
            entry['pos'] = np.asarray(pos).tolist()
        self.index['slides'].append(entry)

    def close(self):
        off = self.f.tell()
        self.f.write(json.dumps(self.index).encode('utf-8'))
        self.f.seek(len(MAGIC))
        self.f.write(off.to_bytes(8, 'little'))
        self.f.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


class FeatureBags(torch.utils.data.Dataset):
    """Memory-mapped reader with the item layout of the reference's `CamelyonFeatures` (camelyon_dataset.py:106-143):
    `ds[i]` -> {'input': (N_i, F) tensor (a zero-copy view of the file), <task name>: label, ...}."""

    def __init__(self, path, tasks=None):
        self.path, self.tasks = path, tasks or {}
        with open(path, 'rb') as f:
            head = f.read(16)
            if head[:8] != MAGIC:
                raise ValueError('%s is not a feature-bag file' % path)
            f.seek(int.from_bytes(head[8:16], 'little'))
            self.index = json.loads(f.read().decode('utf-8'))
        self.F = self.index['n_features']
        self.dtype = self.index['dtype']
        self._map = None

    def __len__(self):
        return len(self.index['slides'])

    @property
    def slide_names(self):
        return [s['name'] for s in self.index['slides']]

    def _mm(self):
        if self._map is None:                       # opened lazily so DataLoader workers map the file themselves
            self._map = np.memmap(self.path, dtype=np.uint8, mode='r')
        return self._map

    def features(self, i):
        s = self.index['slides'][i]
        es = 2 if self.dtype == 'bf16' else 4
        raw = self._mm()[s['offset']: s['offset'] + s['n_patches'] * self.F * es]
        if self.dtype == 'bf16':
            t = torch.from_numpy(raw.view(np.int16).reshape(s['n_patches'], self.F)).view(torch.bfloat16)
        else:
            t = torch.from_numpy(raw.view(np.float32).reshape(s['n_patches'], self.F))
        return t

    def __getitem__(self, i):
        s = self.index['slides'][i]
        item = {'input': self.features(i)}
        for task in self.tasks.values():
            item[task['name']] = s.get('label')
        return item


def convert_hdf5(h5_path, out_path, dtype='bf16'):
    """Rewrite the reference's HDF5 feature file (one group per slide: 'img', 'pos', attrs['label']) as a flat bag."""
    try:
        import h5py
    except ImportError as e:                         # h5py is not part of this image
        raise RuntimeError('convert_hdf5 needs h5py') from e
    with h5py.File(h5_path, 'r') as h5:
        names = list(h5.keys())
        F = h5[names[0]]['img'].shape[1]
        with FeatureBagWriter(out_path, F, dtype) as w:
            for name in names:
                g = h5[name]
                w.add(name, torch.from_numpy(g['img'][:]), int(g.attrs['label']), g['pos'][:] if 'pos' in g else None)
    return out_path

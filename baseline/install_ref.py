"""Place the UNMODIFIED reference's hot-path files under baseline/_ref/ (git-ignored, travels with gpurun).

The reference (benbergner/ips, MIT) is pure Python with no setup.py / pyproject, so there is nothing for
`pip install` to build; the base contract's `baseline/_ref` is therefore filled by a byte-for-byte copy of the files
the path needs: architecture/, utils/, training/, config/ and the package markers.  Nothing is edited.
Runs in the build container only (where /root/reference exists); `__graft_entry__.build()` calls it.

    python baseline/install_ref.py
"""
import filecmp
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, '_ref')
SRC = os.environ.get('IPS_REFERENCE', '/root/reference')
PARTS = ['architecture', 'utils', 'training', 'config', '__init__.py', 'LICENSE']


def install(verbose=True):
    """Copy the reference files; returns True when baseline/_ref is usable afterwards."""
    if not os.path.isdir(SRC):
        return os.path.isdir(os.path.join(DST, 'architecture'))
    os.makedirs(DST, exist_ok=True)
    for part in PARTS:
        s, d = os.path.join(SRC, part), os.path.join(DST, part)
        if os.path.isdir(s):
            if os.path.isdir(d):
                cmp = filecmp.dircmp(s, d)
                if not (cmp.left_only or cmp.diff_files or cmp.funny_files):
                    continue
                shutil.rmtree(d)
            shutil.copytree(s, d, ignore=shutil.ignore_patterns('__pycache__', '*.pyc'))
        elif os.path.isfile(s):
            shutil.copyfile(s, d)
    if verbose:
        print('reference files copied to', DST)
    return True


def ref_path():
    """Directory to put on sys.path to import the unmodified reference, or None."""
    return DST if os.path.isfile(os.path.join(DST, 'architecture', 'ips_net.py')) else None


if __name__ == '__main__':
    sys.exit(0 if install() else 1)

"""Per-kernel SASS mnemonic counts of the built objects (evidence of tcgen05 / TMA / TMEM / cluster instructions).

    python tools/sass_summary.py > profiles/r02_sass_summary.txt

Runs `cuobjdump -sass` on ips_b200/csrc/*.o and counts, per kernel, the mnemonics B200_PROFILING.md lists:
UTCHMMA (tcgen05.mma, `.2CTA` = cta_group::2), UTCBAR (tcgen05.commit), LDTM (tcgen05.ld), UTMALDG / UTMASTG (TMA tensor
loads / stores, `.IM2COL` mode), UTMAPF (bulk prefetch), SYNCS (mbarrier), UCGABAR (cluster barrier), LDG.E.*.256 (256-bit
global loads), MATCH / ATOMS (radix histogram), HMMA (legacy mma.sync: expected 0)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, 'ips_b200', 'csrc')
KEYS = ['UTCHMMA', 'UTCHMMA.2CTA', 'UTCBAR', 'LDTM', 'UTMALDG', 'UTMALDG.IM2COL', 'UTMASTG', 'UTMAPF', 'SYNCS', 'UCGABAR', 'LDG.256',
        'MATCH', 'ATOMS', 'HMMA', 'FFMA', 'MUFU.EX2']


def demangle(name):
    try:
        return subprocess.run(['c++filt', name], capture_output=True, text=True).stdout.strip() or name
    except Exception:
        return name


def main():
    print('# SASS mnemonic counts per kernel (cuobjdump -sass, sm_100a objects); see tools/sass_summary.py')
    print('%-64s %s' % ('kernel', ' '.join('%s' % k for k in KEYS)))
    for obj in sorted(f for f in os.listdir(CSRC) if f.endswith('.o')):
        out = subprocess.run(['cuobjdump', '-sass', os.path.join(CSRC, obj)], capture_output=True, text=True).stdout
        cur, counts = None, collections.OrderedDict()
        for line in out.splitlines():
            m = re.match(r'\s*Function : (\S+)', line)
            if m:
                cur = m.group(1)
                counts[cur] = collections.Counter()
                continue
            if cur is None:
                continue
            m = re.match(r'\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
            if not m:
                continue
            op = m.group(1)
            c = counts[cur]
            if op.startswith('UTCHMMA'):
                c['UTCHMMA'] += 1
                if '.2CTA' in op:
                    c['UTCHMMA.2CTA'] += 1
            elif op.startswith('UTCBAR'):
                c['UTCBAR'] += 1
            elif op.startswith('LDTM'):
                c['LDTM'] += 1
            elif op.startswith('UTMALDG'):
                c['UTMALDG'] += 1
                if 'IM2COL' in op:
                    c['UTMALDG.IM2COL'] += 1
            elif op.startswith('UTMASTG'):
                c['UTMASTG'] += 1
            elif op.startswith('UTMAPF'):
                c['UTMAPF'] += 1
            elif op.startswith('SYNCS'):
                c['SYNCS'] += 1
            elif op.startswith('UCGABAR'):
                c['UCGABAR'] += 1
            elif op.startswith('LDG') and '.256' in op:
                c['LDG.256'] += 1
            elif op.startswith('MATCH'):
                c['MATCH'] += 1
            elif op.startswith('ATOMS'):
                c['ATOMS'] += 1
            elif op.startswith('HMMA'):
                c['HMMA'] += 1
            elif op.startswith('FFMA'):
                c['FFMA'] += 1
            elif op.startswith('MUFU.EX2'):
                c['MUFU.EX2'] += 1
        print('## ' + obj)
        for fn, c in counts.items():
            name = demangle(fn)
            name = re.sub(r'\(anonymous namespace\)::', '', name)
            name = re.sub(r'\(.*', '', name)[:62]
            print('%-64s %s' % (name, ' '.join('%*d' % (len(k), c[k]) for k in KEYS)))


if __name__ == '__main__':
    sys.exit(main())

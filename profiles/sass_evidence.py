"""Static SASS instruction counts per kernel of libmome.so (cuobjdump -sass): which kernels are tcgen05 / TMEM / TMA code.
    python profiles/sass_evidence.py > profiles/r02_sass_evidence.md"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OPS = [('UTCHMMA', 'tcgen05.mma'), ('UTCBAR', 'tcgen05.commit'), ('LDTM', 'tcgen05.ld'), ('STTM', 'tcgen05.st'), ('UTMALDG', 'TMA tensor load'),
       ('UCGABAR', 'cluster barrier'), ('USETMAXREG', 'setmaxnreg'), ('HMMA', 'mma.sync'), ('LDSM', 'ldmatrix'), ('LDGSTS', 'cp.async'),
       ('MUFU.EX2', 'ex2.approx'), ('REDG', 'red.global.add')]


def main():
    sass = subprocess.run(['cuobjdump', '-sass', os.path.join(ROOT, 'exploremultimodal_b200', 'libmome.so')], capture_output=True, text=True).stdout
    cur, counts = None, collections.OrderedDict()
    for line in sass.split('\n'):
        m = re.search(r'Function : (\S+)', line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None or '/*' not in line:
            continue
        for op, _ in OPS:
            if re.search(r'(?<![A-Z0-9_.])' + re.escape(op) + r'(?![A-Z0-9_])', line):
                counts[cur][op] += 1
    print('# SASS evidence (cuobjdump -sass exploremultimodal_b200/libmome.so, sm_100a), end of round 2\n')
    print('Static instruction counts per kernel: ' + ', '.join(f'`{op}` = {what}' for op, what in OPS) + '.\n')
    print('| kernel | ' + ' | '.join(op for op, _ in OPS) + ' |')
    print('|---|' + '---:|' * len(OPS))
    for k, c in counts.items():
        if not (c['UTCHMMA'] or c['HMMA'] or 'attn' in k):
            continue
        d = subprocess.run(['c++filt', k], capture_output=True, text=True).stdout.strip()
        d = re.sub(r'mome::\(anonymous namespace\)::', '', d)
        d = re.sub(r'\((mome|\(anonymous).*', '', d)
        print(f'| `{d[:100]}` | ' + ' | '.join(str(c[op]) for op, _ in OPS) + ' |')


if __name__ == '__main__':
    main()

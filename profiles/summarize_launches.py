"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total time
and share of the captured step.   python profiles/summarize_launches.py launches.csv > summary.md"""
import collections
import csv
import re
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith('==')]
    agg = collections.defaultdict(lambda: [0, 0.0])
    total = 0.0
    for row in csv.DictReader(lines):
        v = float(row['Metric Value'].replace(',', ''))
        v *= {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0}.get(row['Metric Unit'], 1e-6)
        name = re.sub(r'\(.*', '', row['Kernel Name'])
        name = re.sub(r'^void ', '', name)[:100]
        agg[name][0] += 1
        agg[name][1] += v
        total += v
    print(f'captured launches: {sum(n for n, _ in agg.values())}, summed device time {total:.2f} ms '
          '(ncu serialises launches, cold cache: compare shares, not absolutes)\n')
    print('| kernel | launches | total ms | share | avg us |')
    print('|---|---:|---:|---:|---:|')
    for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:30]:
        print(f'| `{k}` | {n} | {t:.3f} | {100 * t / total:.1f}% | {1e3 * t / n:.1f} |')


if __name__ == '__main__':
    main(sys.argv[1])

"""Summarise an `ncu --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv` launch
list: per-kernel count, total time, share of the captured step and (when present) DRAM traffic.
    python profiles/summarize_launches.py launches.csv > summary.md"""
import collections
import csv
import json
import re
import sys

SCALE = {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}


def load(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith('==')]
    agg = collections.defaultdict(lambda: dict(n=0, ms=0.0, rd=0.0, wr=0.0))
    for row in csv.DictReader(lines):
        v = float(row['Metric Value'].replace(',', '')) * SCALE.get(row['Metric Unit'], 1.0)
        name = re.sub(r'\(.*', '', row['Kernel Name'])
        name = re.sub(r'^void ', '', name)[:100]
        a = agg[name]
        m = row['Metric Name']
        if m.startswith('gpu__time_duration'):
            a['n'] += 1
            a['ms'] += v
        elif m.startswith('dram__bytes_read'):
            a['rd'] += v
        elif m.startswith('dram__bytes_write'):
            a['wr'] += v
    return agg


def main(path, json_out=None):
    agg = load(path)
    total = sum(a['ms'] for a in agg.values())
    has_dram = any(a['rd'] + a['wr'] > 0 for a in agg.values())
    print(f'captured launches: {sum(a["n"] for a in agg.values())}, summed device time {total:.2f} ms '
          '(ncu serialises launches, cold cache: compare shares, not absolutes)\n')
    print('| kernel | launches | total ms | share | avg us |' + (' DRAM read GB | DRAM write GB | GB/s |' if has_dram else ''))
    print('|---|---:|---:|---:|---:|' + ('---:|---:|---:|' if has_dram else ''))
    for k, a in sorted(agg.items(), key=lambda x: -x[1]['ms'])[:30]:
        line = f'| `{k}` | {a["n"]} | {a["ms"]:.3f} | {100 * a["ms"] / total:.1f}% | {1e3 * a["ms"] / a["n"]:.1f} |'
        if has_dram:
            line += f' {a["rd"] / 1e9:.2f} | {a["wr"] / 1e9:.2f} | {(a["rd"] + a["wr"]) / 1e9 / (a["ms"] * 1e-3):.0f} |'
        print(line)
    if json_out:
        gem = [a for k, a in agg.items() if 'gemm_pair_kernel' in k]
        n = sum(a['n'] for a in gem)
        out = dict(source=path, kernel='gemm_pair_kernel (all instantiations)', launches=n,
                   dram_bytes_per_launch=(sum(a['rd'] + a['wr'] for a in gem) / n) if n else None,
                   ms_total=sum(a['ms'] for a in gem), share_of_step=sum(a['ms'] for a in gem) / total)
        with open(json_out, 'w') as f:
            json.dump(out, f, indent=1)


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)

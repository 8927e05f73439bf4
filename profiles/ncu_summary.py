"""Extract the judged metrics from `ncu --set full` reports into a markdown table.
    python profiles/ncu_summary.py gpurun_out/prof_a.ncu-rep [more.ncu-rep ...] > profiles/rNN_ncu_full_summary.md"""
import csv
import io
import subprocess
import sys

WANT = [
    ('gpu__time_duration.sum', 'time us'),
    ('launch__registers_per_thread', 'regs'),
    ('launch__grid_size', 'grid'),
    ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor pipe %'),
    ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue active %'),
    ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps active %'),
    ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'DRAM %'),
    ('dram__bytes_read.sum', 'DRAM read MB'),
    ('dram__bytes_write.sum', 'DRAM write MB'),
    ('lts__t_sector_hit_rate.pct', 'L2 hit %'),
]
SCALE = {'Kbyte': 1e-3, 'Mbyte': 1.0, 'Gbyte': 1e3, 'byte': 1e-6, 'us': 1.0, 'ms': 1e3, 'ns': 1e-3}


def main(paths):
    print('| kernel | ' + ' | '.join(n for _, n in WANT) + ' |')
    print('|---|' + '---:|' * len(WANT))
    for path in paths:
        raw = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        hdr, units = rows[0], rows[1]
        for r in rows[2:]:
            name = r[hdr.index('Kernel Name')].split('(')[0][-70:]
            cells = []
            for key, _ in WANT:
                if key not in hdr:
                    cells.append('-')
                    continue
                i = hdr.index(key)
                try:
                    v = float(r[i].replace(',', '')) * SCALE.get(units[i], 1.0)
                    cells.append(f'{v:.1f}' if v < 1e5 else f'{v:.0f}')
                except ValueError:
                    cells.append(r[i])
            print(f'| `{name}` | ' + ' | '.join(cells) + ' |')


if __name__ == '__main__':
    main(sys.argv[1:])

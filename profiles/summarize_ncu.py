"""Turns an .ncu-rep (ncu --set full) into the short per-kernel markdown table we
commit under profiles/.  Usage: python profiles/summarize_ncu.py rep.ncu-rep > out.md"""
import csv
import io
import subprocess
import sys

KEYS = [
    ('gpu__time_duration.sum', 'duration'),
    ('dram__bytes_read.sum', 'DRAM read'),
    ('dram__bytes_write.sum', 'DRAM write'),
    ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'DRAM % of peak'),
    ('lts__t_sector_hit_rate.pct', 'L2 hit %'),
    ('l1tex__t_sector_hit_rate.pct', 'L1 hit %'),
    ('l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'L1 % of peak'),
    ('lts__throughput.avg.pct_of_peak_sustained_elapsed', 'L2 % of peak'),
    ('sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'FP64 pipe % active'),
    ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue slots % busy'),
    ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps active % of max'),
    ('launch__registers_per_thread', 'registers / thread'),
    ('launch__block_size', 'block size'),
    ('launch__grid_size', 'grid size'),
    ('launch__waves_per_multiprocessor', 'waves / SM'),
    ('smsp__inst_executed.sum', 'warp instructions'),
]


def main(rep):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    print(f'# ncu summary of `{rep.split("/")[-1]}` (ncu --set full --clock-control none)\n')
    for r in rows[2:]:
        name = r[hdr.index('Kernel Name')]
        print(f'## `{name[:110]}`\n')
        print('| metric | value |\n|---|---|')
        for k, label in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f'| {label} (`{k}`) | {r[i]} {units[i]} |')
        try:
            rd = float(r[hdr.index('dram__bytes_read.sum')])
            wr = float(r[hdr.index('dram__bytes_write.sum')])
            u = units[hdr.index('dram__bytes_read.sum')]
            print(f'| **DRAM traffic (read+write)** | {rd + wr:.3f} {u} |')
        except Exception:
            pass
        print()


if __name__ == '__main__':
    main(sys.argv[1])

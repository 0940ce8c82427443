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
        for i, k in enumerate(hdr):      # whatever tensor-pipe counters this ncu version exposes
            if 'pipe_tensor' in k and ('pct_of_peak_sustained_active' in k or 'cycles_active.avg' in k) and r[i] not in ('', '0'):
                print(f'| tensor pipe (`{k}`) | {r[i]} {units[i]} |')
        try:
            mult = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}
            ird, iwr = hdr.index('dram__bytes_read.sum'), hdr.index('dram__bytes_write.sum')
            tot = float(r[ird]) * mult[units[ird]] + float(r[iwr]) * mult[units[iwr]]
            print(f'| **DRAM traffic (read+write)** | {tot / 1e9:.3f} Gbyte |')
            dur = float(r[hdr.index('gpu__time_duration.sum')]) * {'ns': 1e-9, 'us': 1e-6, 'ms': 1e-3, 's': 1.0}[
                units[hdr.index('gpu__time_duration.sum')]]
            print(f'| **DRAM GB/s over the launch** | {tot / dur / 1e9:.1f} |')
        except Exception:
            pass
        print()


if __name__ == '__main__':
    main(sys.argv[1])

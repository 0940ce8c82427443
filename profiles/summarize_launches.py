"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list
(cold-cache, serialised launches: compare SHARES, not absolutes).
Usage: python profiles/summarize_launches.py list.csv "title" > out.md"""
import collections
import csv
import re
import sys


def main(path, title):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ik, iv, iu = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    unit = {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 'nsecond': 1e-6, 'usecond': 1e-3, 'msecond': 1.0}
    agg = collections.OrderedDict()
    for r in rows[1:]:
        ms = float(r[iv].replace(',', '')) * unit.get(r[iu], 1e-6)
        key = re.sub(r'\(.*', '', r[ik]).replace('<unnamed>::', '')[:90]
        a = agg.setdefault(key, [0, 0.0])
        a[0] += 1
        a[1] += ms
    tot = sum(a[1] for a in agg.values())
    print(f'# {title}\n')
    print('`ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised: shares, not absolutes).\n')
    print('| kernel | launches | total ms | share | avg us |\n|---|---|---|---|---|')
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        if a[1] / tot < 0.002:
            continue
        print(f'| `{k}` | {a[0]} | {a[1]:.3f} | {100 * a[1] / tot:.1f}% | {1e3 * a[1] / a[0]:.1f} |')
    print(f'\ntotal {tot:.2f} ms over {sum(a[0] for a in agg.values())} launches')


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else sys.argv[1])

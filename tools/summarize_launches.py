"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into a markdown table.

    python tools/summarize_launches.py gpurun_out/launches.csv "title" > profiles/rNN_ncu_launches.md
"""
import csv
import re
import sys


def short(name):
    name = re.sub(r'\(anonymous namespace\)::|<unnamed>::', '', name)
    name = re.sub(r'^void ', '', name)
    m = re.match(r'([A-Za-z_0-9]+(<[^(]*>)?)', name)
    s = m.group(1) if m else name
    s = re.sub(r'\(int\)|\(bool\)', '', s)
    s = re.sub(r'StencilMat<[^>]*>', 'StencilMat', s)
    return s


def main():
    path, title = sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else ''
    rows = []
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get('Metric Name') == 'gpu__time_duration.sum':
            v = float(r['Metric Value'].replace(',', ''))
            if r.get('Metric Unit') == 'us':
                v *= 1e3
            rows.append((short(r['Kernel Name']), v))
    tot = sum(v for _, v in rows)
    agg = {}
    for k, v in rows:
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1; a[1] += v
    print('# ncu launch list summary: %s\n' % title)
    print('Source: `ncu --metrics gpu__time_duration.sum --clock-control none --csv` (per-launch times are cold-cache and '
          'serialised: compare SHARES).')
    print('Total kernel time %.1f ms over %d launches.\n' % (tot / 1e6, len(rows)))
    print('| kernel | launches | total us | avg us | share |')
    print('|---|---:|---:|---:|---:|')
    for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        if v / tot < 0.0005:
            continue
        print('| `%s` | %d | %.1f | %.1f | %.1f%% |' % (k, c, v / 1e3, v / 1e3 / c, 100 * v / tot))


if __name__ == '__main__':
    main()

"""Per-source-line instruction counts / stall samples of a kernel in an .ncu-rep (cuda,sass correlated view).
Usage: ncu_lines.py <rep> <kernel-regex> [N]"""
import collections, csv, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
N = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--print-source', 'cuda,sass', '--csv', '--kernel-name', 'regex:' + rx],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
fname, agg, src = None, collections.OrderedDict(), {}
hdr = None
seen_kernel = 0
for r in rows:
    if not r:
        continue
    if r[0] == 'File Path':
        fname = r[1].split('/')[-1]
        continue
    if r[0] == 'Function Name':
        continue
    if r[0] == 'Line No':
        hdr = r
        ci, cs = hdr.index('Instructions Executed'), hdr.index('# Samples')
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    line = r[0]
    if line:
        cur = (fname, int(line))
        src[cur] = r[1]
    try:
        inst, smp = int(r[ci] or 0), int(r[cs] or 0)
    except ValueError:
        continue
    a = agg.setdefault(cur, [0, 0])
    a[0] += inst
    a[1] += smp
ti, ts = sum(a[0] for a in agg.values()), sum(a[1] for a in agg.values())
print(f'total warp instructions {ti}, samples {ts}')
top = sorted(agg.items(), key=lambda kv: -kv[1][0])[:N]
for (f, l), (i, s) in sorted(top):
    print(f'{f}:{l:<5d} inst {100 * i / ti:5.1f}%  samples {100 * s / max(ts, 1):5.1f}%  {src[(f, l)].strip()[:110]}')

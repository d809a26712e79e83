"""Top SASS instructions of a kernel in an .ncu-rep by stall samples (dev aid): ncu_hot.py <rep> <kernel-regex> [N]"""
import csv, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
N = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', 'regex:' + rx], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
for h in hi[:1]:
    hdr = rows[h]
    body = [r for r in rows[h + 1:] if len(r) == len(hdr)]
    c = {k: hdr.index(k) for k in ('Source', '# Samples', 'Instructions Executed', 'L1 Wavefronts Shared', 'L1 Wavefronts Shared Ideal')}
    tot = sum(int(r[c['# Samples']] or 0) for r in body)
    toti = sum(int(r[c['Instructions Executed']] or 0) for r in body)
    print(f'total samples {tot}, warp instructions {toti}, sass lines {len(body)}')
    stalls = [k for k in hdr if k.startswith('stall_') and 'Not Issued' not in k]
    agg = {k: sum(int(r[hdr.index(k)] or 0) for r in body) for k in stalls}
    print('stalls:', ', '.join(f'{k[6:]} {v*100//max(tot,1)}%' for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
    idx = sorted(range(len(body)), key=lambda i: -int(body[i][c['# Samples']] or 0))[:N]
    for i in sorted(idx):
        r = body[i]
        top = max(stalls, key=lambda k: int(r[hdr.index(k)] or 0))
        print(f"{i:5d} {int(r[c['# Samples']]):6d} {int(r[c['Instructions Executed']]):9d} wf {r[c['L1 Wavefronts Shared']]:>9}/{r[c['L1 Wavefronts Shared Ideal']]:<9} {top[6:]:12s} {r[c['Source']].strip()[:90]}")

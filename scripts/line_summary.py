"""per-source-line instruction counts and stall samples of an ncu report (needs -lineinfo + --import-source on)
   python scripts/line_summary.py <report.ncu-rep> [min share] [kernel-name substring]"""
import csv, subprocess, sys, collections
def fl(x):
    try: return float(x)
    except ValueError: return 0.0
rep = sys.argv[1]; thresh = float(sys.argv[2]) if len(sys.argv) > 2 else 0.01
only = sys.argv[3] if len(sys.argv) > 3 else None
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
kernels = collections.OrderedDict(); cur = None; fpath = None; hdr = None; line = None
for r in rows:
    if r and r[0] == 'File Path': fpath = r[1].split('/')[-1]; continue
    if r and r[0] == 'Function Name':
        cur = kernels.setdefault(r[1][:70], dict(per=collections.OrderedDict(), src={})); hdr = None; continue
    if r and r[0] == 'Line No':
        hdr = r; ia = hdr.index('Instructions Executed'); ism = hdr.index('# Samples'); ith = hdr.index('Thread Instructions Executed')
        stall = [(i, h) for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]; continue
    if hdr is None or len(r) != len(hdr) or cur is None: continue
    if r[0]:
        line = (fpath, int(r[0])); cur['src'][line] = r[1]
    if r[2] and line:
        v = cur['per'].setdefault(line, [0.0, 0.0, 0, 0.0, collections.Counter()]); v[0] += fl(r[ia]); v[1] += fl(r[ism]); v[2] += 1; v[3] += fl(r[ith])
        for i, h in stall:
            if fl(r[i]): v[4][h[6:]] += fl(r[i])
for name, k in kernels.items():
    if only and only not in name: continue
    per, src = k['per'], k['src']
    tot = sum(v[0] for v in per.values()); tots = sum(v[1] for v in per.values())
    print('== %s: total inst %.3e samples %d' % (name, tot, tots))
    for key, v in per.items():
        if v[0] / max(tot, 1) > thresh or v[1] / max(tots, 1) > thresh:
            top = ','.join(f'{a}:{int(b)}' for a, b in v[4].most_common(2))
            print(f'{key[0]:>20}:{key[1]:<4} inst {v[0] / tot * 100:5.1f}%  samples {v[1] / max(tots, 1) * 100:5.1f}%  thr/inst {v[3] / max(v[0], 1):4.1f} sass {v[2]:3d} {top:28s}| {src[key].strip()[:90]}')

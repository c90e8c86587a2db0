"""per-source-line instruction counts of an ncu report (needs -lineinfo + --import-source on)"""
import csv, subprocess, sys, collections
def fl(x):
    try: return float(x)
    except ValueError: return 0.0
rep = sys.argv[1]; thresh = float(sys.argv[2]) if len(sys.argv) > 2 else 0.01
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
per = collections.OrderedDict(); src = {}; fpath = None; hdr = None; line = None
for r in rows:
    if r and r[0] == 'File Path': fpath = r[1].split('/')[-1]; continue
    if r and r[0] == 'Function Name': continue
    if r and r[0] == 'Line No': hdr = r; ia = hdr.index('Instructions Executed'); ism = hdr.index('# Samples'); continue
    if hdr is None or len(r) != len(hdr): continue
    if r[0]:
        line = (fpath, int(r[0])); src[line] = r[1]
    if r[2] and line:
        v = per.setdefault(line, [0.0, 0.0, 0]); v[0] += fl(r[ia]); v[1] += fl(r[ism]); v[2] += 1
tot = sum(v[0] for v in per.values()); tots = sum(v[1] for v in per.values())
print('total inst %.3e samples %d' % (tot, tots))
for k, v in per.items():
    if v[0] / tot > thresh or v[1] / max(tots, 1) > thresh:
        print(f'{k[0]:>20}:{k[1]:<4} inst {v[0] / tot * 100:5.1f}%  samples {v[1] / max(tots, 1) * 100:5.1f}%  sass {v[2]:3d} | {src[k].strip()[:100]}')

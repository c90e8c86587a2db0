"""hot SASS regions of a kernel in an ncu report: every instruction counted once (the source page repeats an instruction under
every line of its inline chain), grouped by the set of source lines it is attributed to
   python scripts/sass_hot.py <report.ncu-rep> <kernel substring> [top n]"""
import csv, subprocess, sys, collections
def fl(x):
    try: return float(x)
    except ValueError: return 0.0
rep, only = sys.argv[1], sys.argv[2]; top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
addr = {}; fpath = None; hdr = None; line = None; use = False; srcs = {}
for r in rows:
    if r and r[0] == 'File Path': fpath = r[1].split('/')[-1]; continue
    if r and r[0] == 'Function Name': use = only in r[1]; hdr = None; continue
    if r and r[0] == 'Line No':
        hdr = r; ia = hdr.index('Instructions Executed'); ism = hdr.index('# Samples'); ith = hdr.index('Thread Instructions Executed'); continue
    if not use or hdr is None or len(r) != len(hdr): continue
    if r[0]: line = f'{fpath}:{r[0]}'; srcs[line] = r[1].strip()[:70]
    if r[2] and line:
        a = addr.setdefault(r[2], dict(sass=r[3].strip(), inst=fl(r[ia]), smp=fl(r[ism]), thr=fl(r[ith]), lines=[]))
        a['lines'].append(line)
tot = sum(a['inst'] for a in addr.values()); tots = sum(a['smp'] for a in addr.values())
print(f'unique SASS {len(addr)}  total warp inst {tot:.4e}  samples {int(tots)}')
grp = collections.OrderedDict()
for k in sorted(addr):
    a = addr[k]; key = ' < '.join(a['lines'][:3])
    g = grp.setdefault(key, [0.0, 0.0, 0, 0.0]); g[0] += a['inst']; g[1] += a['smp']; g[2] += 1; g[3] += a['thr']
for key, g in sorted(grp.items(), key=lambda kv: -kv[1][0])[:top]:
    first = key.split(' < ')[0]
    print(f'{g[0] / tot * 100:5.1f}% inst {g[1] / max(tots, 1) * 100:5.1f}% smp  thr/inst {g[3] / max(g[0], 1):4.1f} sass {g[2]:3d}  {key[:80]:80s} | {srcs.get(first, "")}')

"""summarise an ncu report's SASS page: opcode mix and execution-count plateaus (loop levels)"""
import csv, collections, subprocess, sys
import numpy as np
rep = sys.argv[1]
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
blocks = []; cur = None
for r in rows:
    if r and r[0] == 'Kernel Name':
        cur = {'name': r[1], 'rows': []}; blocks.append(cur); continue
    if cur is not None:
        cur['rows'].append(r)
for b in blocks:
    hdr = b['rows'][0]; data = [r for r in b['rows'][1:] if len(r) == len(hdr)]
    ia = hdr.index('Instructions Executed'); isrc = hdr.index('Source'); it = hdr.index('Thread Instructions Executed')
    c = np.array([float(r[ia] or 0) for r in data]); tot = c.sum(); th = np.array([float(r[it] or 0) for r in data])
    print(b['name'][:60], 'total inst %.3e' % tot, 'n sass', len(data), 'avg active threads %.1f' % (th.sum() / tot))
    ops = collections.Counter()
    for r, v in zip(data, c):
        t = r[isrc].split(); op = t[1] if t[0].startswith('@') else t[0]
        ops[op.split('.')[0]] += v
    print(' '.join(f'{k}:{v / tot * 100:.1f}%' for k, v in ops.most_common(24)))
    vals, n = np.unique(np.round(c / 1e6), return_counts=True)
    for v, nn in sorted(zip(vals, n), key=lambda t: -t[1] * t[0])[:16]:
        print(f'  exec={v:9.0f}M  x{nn} sass -> {v * nn * 1e6 / tot * 100:5.1f}% of instr')

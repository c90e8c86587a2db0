"""instruction share per code region (function-level line ranges) of an ncu report with -lineinfo source:
   python scripts/region_summary.py <report.ncu-rep> [kernel substring]"""
import csv, subprocess, sys, collections
rep = sys.argv[1]; pick = sys.argv[2] if len(sys.argv) > 2 else ''
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
per = collections.defaultdict(lambda: collections.OrderedDict()); src = {}
fpath = fn = None; hdr = None
for r in rows:
    if r and r[0] == 'File Path': fpath = r[1].split('/')[-1]; continue
    if r and r[0] == 'Function Name': fn = r[1]; continue
    if r and r[0] == 'Line No': hdr = r; ia = hdr.index('Instructions Executed'); ism = hdr.index('# Samples'); continue
    if hdr is None or len(r) != len(hdr) or not r[0].isdigit(): continue
    try: v = float(r[ia]); s = float(r[ism])
    except ValueError: continue
    k = (fpath, int(r[0])); src[k] = r[1]
    a = per[fn].setdefault(k, [0.0, 0.0]); a[0] += v; a[1] += s
REGIONS = [('nearest.cuh', 52, 72, 'cone fns'), ('nearest.cuh', 75, 110, 'lane_query_setup (stage 1)'), ('nearest.cuh', 112, 122, 'shfl_window'),
           ('nearest.cuh', 132, 150, 'sweep_cell'), ('nearest.cuh', 154, 173, 'phase1_chunk'), ('nearest.cuh', 176, 192, 'phase2_chunk'),
           ('nearest.cuh', 198, 274, 'nearest_w body'), ('nearest.cuh', 279, 322, 'batch loop'), ('rsgpu_internal.cuh', 125, 143, 'exact math'),
           ('rsgpu_internal.cuh', 156, 185, 'make_window'), ('rsgpu_internal.cuh', 188, 208, 'axis_gap/window_cell')]
for fn, lines in per.items():
    if pick not in fn: continue
    tot = sum(v[0] for v in lines.values()); ts = sum(v[1] for v in lines.values())
    print(fn[:90], 'inst %.3e' % tot)
    agg = collections.OrderedDict()
    for (f, l), v in lines.items():
        for rf, a, b, nm in REGIONS:
            if f == rf and a <= l <= b: key = nm; break
        else: key = f
        x = agg.setdefault(key, [0.0, 0.0]); x[0] += v[0]; x[1] += v[1]
    for k, v in sorted(agg.items(), key=lambda t: -t[1][0]): print(f'   inst {v[0] / tot * 100:5.1f}%  samples {v[1] / max(ts, 1) * 100:5.1f}%  {k}')

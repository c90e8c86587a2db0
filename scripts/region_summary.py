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
# (file, first line, last line, name) - line ranges of the sources as of the end of round 1; adjust after edits
REGIONS = [('nearest_group.cuh', 70, 100, 'stage1_test (pass A)'), ('nearest_group.cuh', 102, 106, 'normal_dot_call'),
           ('nearest_group.cuh', 108, 125, 'group_min / group_sum'), ('nearest_group.cuh', 130, 166, 'group_search: window + gaps'),
           ('nearest_group.cuh', 167, 205, 'phase 1: cell table'), ('nearest_group.cuh', 206, 243, 'phase 2: sweeps'),
           ('nearest_group.cuh', 244, 276, 'phase 3: rank count'), ('nearest_group.cuh', 278, 305, 'group_round glue'),
           ('rsgpu_internal.cuh', 130, 138, 'xf_apply'), ('rsgpu_internal.cuh', 140, 145, 'dist2_exact'), ('rsgpu_internal.cuh', 147, 150, 'dot3_exact'),
           ('rsgpu_internal.cuh', 152, 195, 'make_window'), ('nearest.cuh', 40, 80, 'cone tests'),
           ('score.cu', 120, 175, 'score: setup + pass A loop'), ('score.cu', 176, 210, 'score: pass B glue / query_of'),
           ('score.cu', 211, 235, 'score: pass C terms + sum')]
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

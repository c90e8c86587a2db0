import os, re, subprocess, sys, time, shutil
ROOT = "/root/repo"
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "integration"))
import make_dropin_case
folder = "/tmp/pp_exp"
shutil.rmtree(folder, ignore_errors=True)
db, scan, out, scan1 = make_dropin_case.write_case(folder, n_objects=10, n_static=2, room=(7.0, 2.6, 5.0), target_points=200000)
exe = os.path.join(ROOT, "integration", "_build", "pose_proposal_rsgpu")
for env in ({}, {}, {"RSGPU_DENSE_CAP": "16777216"}, {"RSGPU_DENSE_CAP": "4194304"}, {"RSGPU_DENSE_SERIAL": "0"}):
    t0 = time.perf_counter()
    r = subprocess.run([exe, db, scan, out, "-v"], capture_output=True, text=True, env=dict(os.environ, **env))
    wall = time.perf_counter() - t0
    lines = [l for l in r.stdout.splitlines() if re.search(r"Done in|Computed poses|IO:.*ms|took|Loading", l)]
    print(env, round(wall, 2), [l.strip()[:90] for l in lines][:8], flush=True)

"""diagnostic: GPU ICP vs oracle, iteration by iteration"""
import sys, numpy as np
sys.path.insert(0, '.')
from oracle import orcbind as O
from rescan_b200 import api, synth
from tests import common
from tests.test_gpu_parity import _pose_delta
scene = common.small_scene()
p2, n2 = scene.scan.pos(2), scene.scan.nor(2)
grid = api.HashGrid(p2, 0.05, normals=n2)
og = O.OrcGrid(p2, 0.10)
rng = np.random.default_rng(31)
ang = np.float32(np.deg2rad(60.0))
for o in scene.objects[:2]:
    starts = [common.colmajor(m) for _, m in common.perturbed_poses(rng, type("S", (), {"objects": [o]})(), 2, 0.03, 0.08)]
    cloud = api.PointCloud(o.cloud.pos(2), o.cloud.nor(2))
    for s in starts:
        To, eo, ito = O.icp_align(o.cloud.pos(2), o.cloud.nor(2), p2, n2, s, 0.10, ang)
        Tg, eg, itg = api.icp_align(cloud, grid, s[None], 0.10, ang)
        print('full: iters', ito, itg[0], 'err', eo, eg[0], 'delta', _pose_delta(Tg[0], To))
        # step by step with the oracle
        T = s.copy(); md = np.float32(0.10)
        for it in range(1, 12):
            c = O.icp_find_corrs(o.cloud.pos(2), o.cloud.nor(2), og, p2, n2, T, md, ang)
            T, e = O.icp_pt2pl(c[0], c[2], c[3], c[4], T)
            Tg, eg, _ = api.icp_align(cloud, grid, s[None], 0.10, ang, max_iter=it)
            print('  it', it, 'ncorr', len(c[4]), 'w0', int((c[4] == 0).sum()), 'err', e, eg[0], 'delta', _pose_delta(Tg[0], T))
            md = np.float32(max(float(md) * 0.95, 0.05))

"""CPU: host-side logic of the product (pose grids, sharding, merging) and the C-ABI surface of librsgpu.so."""
import ctypes
import os
import re

import numpy as np
import pytest

from oracle import orcbind as O
from rescan_b200 import api, pipeline, posegrid, synth
from tests import common

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    """the C-ABI library loads and exports every function include/rsgpu.h declares (no compute call is made)"""
    hdr = open(os.path.join(ROOT, "include", "rsgpu.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(rsgpu_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 30
    lib = ctypes.CDLL(api.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"librsgpu.so does not export {name}"
    assert declared == set(api.SIGNATURES), declared ^ set(api.SIGNATURES)


def test_no_device_fails_loudly():
    if api.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(api.RsgpuError, match="no CUDA device"):
        api.HashGrid(np.zeros((4, 3), np.float32), 0.05)


def test_segment_transfer_dropin_without_device_exits_like_the_reference(tmp_path):
    """the drop-in executable has no CPU path either: without a device its first replaced call (the scan rasterisation)
    ends the run with the reference's own exit(-1) and the rsgpu error text"""
    import subprocess
    import sys
    exe = os.path.join(ROOT, "integration", "_build", "segment_transfer_rsgpu")
    if api.device_count() > 0 or not os.path.exists(exe):
        pytest.skip("a CUDA device is present, or integration/_build is not built")
    sys.path.insert(0, os.path.join(ROOT, "integration"))
    import make_dropin_case
    folder = str(tmp_path)
    db, scan, _, _ = make_dropin_case.write_case(folder)
    out = make_dropin_case.write_pose_proposal_output(folder, db, scan, os.path.join(ROOT, "tests", "golden", "dropin_pp.bin"))
    r = subprocess.run([exe, out, "-o", os.path.join(folder, "out", "x.rsdb")], capture_output=True, text=True, timeout=300)
    assert r.returncode != 0
    assert "rsgpu drop-in" in r.stderr and "no CUDA device" in r.stderr


@pytest.mark.parametrize("variant", ["segment_transfer_fake", "segment_transfer_fake_planes", "segment_transfer_fake_all"])
def test_segment_transfer_shim_host_logic_against_cpu_reference(tmp_path, variant):
    """integration/rsgpu_dropin_st.cpp (and, for the _planes variant, integration/rsgpu_dropin_planes.cpp: the RANSAC rounds of
    the wall / floor detector drawn first and counted in one call each; for _all also integration/rsgpu_dropin_levels.cpp: level
    building at its call site) linked with the reference's unmodified segment_transfer sources and, in place of
    librsgpu.so, tests/fake_rsgpu (the rsgpu entry points the shim uses, backed by the CPU oracle): the shim's host logic -
    placement order and the two labelling passes, the per-placement mask cache of the coverage term under 25 000 annealing
    moves, edge de-duplication, label maps - must lead to what the pure-CPU reference build decided
    (tests/golden/dropin_st.npz): same placements, poses within 1e-5 m / 1e-5 rad, identical per-vertex labels."""
    import subprocess
    import sys
    if not os.path.isdir("/root/reference/apps/segment_transfer"):
        pytest.skip("needs the reference tree to compile segment_transfer")
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "tests", "fake_rsgpu")], stdout=subprocess.DEVNULL)
    exe = os.path.join(ROOT, "tests", "fake_rsgpu", "_build", variant)
    sys.path.insert(0, os.path.join(ROOT, "integration"))
    import make_dropin_case
    folder = str(tmp_path)
    db, scan, _, _ = make_dropin_case.write_case(folder)
    out = make_dropin_case.write_pose_proposal_output(folder, db, scan, os.path.join(ROOT, "tests", "golden", "dropin_pp.bin"))
    stdout, rows, ply = make_dropin_case.run_segment_transfer(exe, out, folder)
    assert "(GPU)" in stdout  # the shim's replaced stages ran (here on the stand-in)
    assert ("candidates counted in one call each" in stdout) == (variant != "segment_transfer_fake")
    assert ("selected by rsgpu_poisson_level" in stdout) == variant.endswith("_all")
    g = np.load(os.path.join(ROOT, "tests", "golden", "dropin_st.npz"))
    assert [r[0] for r in rows] == list(g["placement_uidx"]) and [r[2] for r in rows] == list(g["object_idx"])
    for r, score, pose in zip(rows, g["score"], g["pose"]):
        A, B = r[4].astype(np.float64), pose.astype(np.float64)
        R = A[:3, :3] @ B[:3, :3].T
        assert np.linalg.norm(A[:3, 3] - B[:3, 3]) < 1e-5 and np.linalg.norm(R - R.T) / (2 * np.sqrt(2)) < 1e-5
        assert abs(r[3] - score) <= 1e-4 * max(abs(score), 1e-3)
    assert len(ply) == len(g["x"]) and (np.asarray(ply["x"], np.float32) == g["x"]).all()
    assert (np.asarray(ply["class_idx"], np.int32) == g["class_idx"]).all()
    assert (np.asarray(ply["instance_idx"], np.int32) == g["instance_idx"]).all()


def test_product_never_imports_oracle():
    for root, _, files in os.walk(os.path.join(ROOT, "rescan_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(root, f)).read()
                bad = re.search(r"^\s*(from|import)\s+oracle\b|orcbind|refbind|librescan_oracle|librescan_ref|rescan_oracle\.c", src, flags=re.M)
                assert bad is None, f"{f} references the oracle: {bad.group(0)}"


def test_posegrid_matches_reference_arithmetic():
    rng = np.random.default_rng(0)
    for a in list(posegrid.rotation_angles(36)) + list(rng.uniform(-7, 7, 500).astype(np.float32)):
        assert (posegrid.make_pose(a, 1.5, 0.0, -2.25) == O.make_pose(a, 1.5, 0.0, -2.25)).all()
    assert len(posegrid.rotation_angles(10)) == 10  # the reference's own setting (pose_proposal.cpp:28)
    assert posegrid.rotation_xforms(36).shape == (36, 16) and posegrid.rotation_xforms(72).shape == (72, 16)
    z, scan, _ = common.golden()
    t = posegrid.reference_translation_grid(z["scan_bbox"][:3], z["scan_bbox"][3:], 0.10)
    assert (t == synth.reference_translation_grid(type("C", (), {"bbox_min": z["scan_bbox"][:3], "bbox_max": z["scan_bbox"][3:]})(), 0.10)).all()
    g = posegrid.pose_grid(posegrid.rotation_xforms(4), t[:5])
    assert g.shape == (5, 4, 16) and (g[3, 2, 12:15] == t[3]).all() and (g[:, :, 15] == 1).all()


def test_shard_range_partitions():
    for n in (0, 1, 7, 2048, 20000):
        for world in (1, 2, 3, 8):
            spans = [pipeline.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_merge_topk_is_deterministic_and_order_independent():
    rng = np.random.default_rng(1)
    props = rng.uniform(0, 1, (40, 17)).astype(np.float32)
    props[5, 16] = props[9, 16] = props[30, 16] = 0.75  # ties broken by pose id
    ids = rng.permutation(1000)[:40].astype(np.int64)
    a = pipeline.merge_topk([props[:13], props[13:]], [ids[:13], ids[13:]], 10)
    b = pipeline.merge_topk([props[25:], props[:25]], [ids[25:], ids[:25]], 10)
    assert (a[0] == b[0]).all() and (a[1] == b[1]).all()
    assert (np.diff(a[0][:, 16]) <= 0).all()
    full = pipeline.merge_topk([props], [ids], 0)
    assert len(full[0]) == 40


def test_cloud_centroid_matches_oracle():
    """rs_pointcloud_centroid (rs_pointcloud.h:1319-1339): the product's host helper against the oracle's restatement"""
    from oracle import orcbind as O
    from rescan_b200 import posegrid
    rng = np.random.default_rng(3)
    for n in (1, 7, 1000, 50_000):
        p = (rng.standard_normal((n, 3)) * 3 + 1).astype(np.float32)
        assert (posegrid.cloud_centroid(p) == O.centroid(p)).all()
    assert (posegrid.cloud_centroid(np.zeros((0, 3), np.float32)) == 0).all()


def test_proposal_bin_round_trip(tmp_path):
    """the proposal wire format (apps/pose_proposal/main.cpp:61-89): re-writing the file the reference's CPU executable
    wrote (tests/golden/dropin_pp.bin) reproduces it byte for byte"""
    from rescan_b200 import rsio
    src = os.path.join(ROOT, "tests", "golden", "dropin_pp.bin")
    props = rsio.read_proposals(src)
    out = rsio.write_proposals(str(tmp_path / "pp.bin"), props)
    assert open(out, "rb").read() == open(src, "rb").read()
    empty = rsio.write_proposals(str(tmp_path / "e.bin"), [np.zeros((0, 17), np.float32)] * 3)
    assert [len(p) for p in rsio.read_proposals(empty)] == [0, 0, 0]


def test_host_wait_policy():
    """lane threads spin on short waits unless lanes x local ranks outnumber the cores"""
    pytest.importorskip("ctypes")
    if not os.path.exists(api.LIB_PATH):
        pytest.skip("librsgpu.so not built")
    cores = os.cpu_count() or 1
    assert pipeline.configure_host_waits(1, local_world=1) == ((1 + 2) > cores)
    assert pipeline.configure_host_waits(8, local_world=cores) is True
    api.set_option("sync", None)


def test_readers_on_reference_written_files(tmp_path):
    """.rsdb and PLY as the reference's own CPU executable wrote them (tests/golden/dropin_pp.rsdb, dropin_obj_102.ply, from
    integration/make_dropin_case.py --golden) are read back; our writers' files round-trip through the same readers"""
    from rescan_b200 import rsio
    gold = os.path.join(ROOT, "tests", "golden")
    db = rsio.read_database(os.path.join(gold, "dropin_pp.rsdb"))
    assert db["version"] == "1.0" and len(db["classes"]) == len(synth.CLASS_NAMES)
    assert [c[0] for c in db["classes"]] == list(synth.CLASS_NAMES)
    assert [o[1] for o in db["objects"]] == [100, 101, 102] and db["n_arrangements"] == 1
    assert len(db["scenes"]) == 2 and db["scenes"][0][3] is None and db["scenes"][1][3].endswith("scan1_pp.bin")
    assert len(db["poses"]) == 3
    for _, arr, obj, score, m in db["poses"]:
        assert arr == 0 and score == 1.0 and (m[3] == [0, 0, 0, 1]).all()
        assert np.allclose(m[:3, :3] @ m[:3, :3].T, np.eye(3), atol=1e-5) and m[1, 3] == 0.0  # yaw + translation on the floor
    v = rsio.read_ply(os.path.join(gold, "dropin_obj_102.ply"))
    assert len(v) == 3069 and set(v.dtype.names) >= {"x", "y", "z", "nx", "ny", "nz", "radius", "class_idx", "instance_idx"}
    nn = np.stack([v["nx"], v["ny"], v["nz"]], axis=1)
    assert np.allclose(np.linalg.norm(nn, axis=1), 1.0, atol=1e-4) and (v["instance_idx"] == 102).all()
    # our writers -> our readers
    scene = common.tiny_scene()
    p = str(tmp_path / "scan.ply")
    rsio.write_ply(p, scene.scan.pos(0), scene.scan.nor(0), scene.scan_class, scene.scan_instance)
    w = rsio.read_ply(p)
    assert (np.stack([w["x"], w["y"], w["z"]], axis=1) == scene.scan.pos(0)).all() and (w["class_idx"] == scene.scan_class).all()
    path = rsio.write_database(str(tmp_path), "db0", scene, p, [(i, o.pose) for i, o in enumerate(scene.objects)])
    d2 = rsio.read_database(path)
    assert len(d2["objects"]) == len(scene.objects) and len(d2["poses"]) == len(scene.objects)
    for (_, _, oi, _, m), o in zip(d2["poses"], scene.objects):
        assert np.allclose(m, o.pose, atol=1e-6)


def test_spatial_order_is_a_permutation_that_shortens_the_walk():
    rng = np.random.default_rng(0)
    t = rng.uniform(0, 7, (3000, 3)).astype(np.float32)
    o = posegrid.spatial_order(t)
    assert sorted(o.tolist()) == list(range(len(t)))
    step = lambda a: np.linalg.norm(np.diff(a[:, [0, 2]], axis=0), axis=1).mean()
    assert step(t[o]) < 0.2 * step(t)
    assert len(posegrid.spatial_order(np.zeros((0, 3), np.float32))) == 0
    assert list(posegrid.spatial_order(np.zeros((1, 3), np.float32))) == [0]


def test_header_is_plain_c_and_library_fails_loudly_without_a_device(tmp_path):
    """include/rsgpu.h compiles as C99, a C program links against librsgpu.so, and - in this container, which has no GPU -
    a compute entry point reports RSGPU_ERR_NO_DEVICE instead of falling back to anything"""
    import shutil
    import subprocess
    if shutil.which("gcc") is None or not os.path.exists(api.LIB_PATH):
        pytest.skip("gcc or librsgpu.so missing")
    src = tmp_path / "abi.c"
    src.write_text(r'''
#include <stdio.h>
#include <string.h>
#include "rsgpu.h"
int main( void )
{
  float pts[9] = { 0, 0, 0, 1, 0, 0, 0, 1, 0 };
  rsgpu_grid_t* g = NULL;
  rsgpu_propose_opts_t o;
  rsgpu_propose_default_opts( &o );
  if( strncmp( rsgpu_version(), "rsgpu", 5 ) != 0 || o.max_n_neigh != 64 || o.translation_ids != NULL ) { return 2; }
  int n = rsgpu_device_count();
  int s = rsgpu_grid_create( pts, 3, 0.05f, &g );
  printf( "%d %d %s\n", n, s, rsgpu_last_error() );
  if( g ) { rsgpu_grid_destroy( g ); }
  return 0;
}
''')
    exe = tmp_path / "abi"
    libdir = os.path.dirname(api.LIB_PATH)
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                           "-L", libdir, "-lrsgpu", f"-Wl,-rpath,{libdir}"])
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    n, status = (int(x) for x in out.stdout.split()[:2])
    if n == 0:
        assert status == -1 and "no CUDA device" in out.stdout  # RSGPU_ERR_NO_DEVICE: there is no CPU fallback
    else:
        assert status == 0


def test_sharding_and_merge_properties():
    """shard_range tiles [0, n) in rank order with sizes differing by at most one; merging per-shard top-k lists gives the
    top-k of the whole (the identity the pose-sharded step rests on), for any split and any k"""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=60, deadline=None)
    @given(n=st.integers(0, 300), world=st.integers(1, 9), k=st.integers(0, 40), seed=st.integers(0, 2 ** 16))
    def check(n, world, k, seed):
        edges = [pipeline.shard_range(n, r, world) for r in range(world)]
        assert edges[0][0] == 0 and edges[-1][1] == n
        sizes = [hi - lo for lo, hi in edges]
        assert all(a[1] == b[0] for a, b in zip(edges, edges[1:])) and max(sizes) - min(sizes) <= 1
        rng = np.random.default_rng(seed)
        props = np.zeros((n, 17), np.float32)
        props[:, 16] = rng.choice(np.array([-1.0, 0.3, 0.5, 0.5, 0.9], np.float32), n)  # many ties
        ids = np.arange(n, dtype=np.int64) * 3
        whole = pipeline.merge_topk([props], [ids], k)
        parts = [pipeline.merge_topk([props[lo:hi]], [ids[lo:hi]], k) for lo, hi in edges]
        merged = pipeline.merge_topk([p[0] for p in parts], [p[1] for p in parts], k)
        assert (merged[1] == whole[1]).all() and (merged[0] == whole[0]).all()
    check()


def test_coverage_grid_host_arithmetic_matches_oracle():
    """api.coverage_grid (isect_grid3d_init in float32 on the host) against the oracle's restatement"""
    rng = np.random.default_rng(2)
    for _ in range(50):
        mn = rng.uniform(-5, 5, 3).astype(np.float32)
        mx = (mn + rng.uniform(0.1, 9, 3)).astype(np.float32)
        for voxel in (0.05, 0.15, 0.1):
            res, origin = api.coverage_grid(mn, mx, voxel)
            ro, oo, _ = O.cov_grid(mn, mx, voxel)
            assert (res == ro).all() and (origin == oo).all()
    assert api.coverage_score(np.zeros((0, 3), np.uint32), 10) == 0 and api.coverage_score(np.array([[1, 0], [2, 0]], np.uint32), 4) == np.float32(0.5)


def test_shard_translations_partitions_every_translation_once():
    """the pose-sharded step's shares: blocks of the Z-order curve dealt round the ranks - a partition of the caller's ids for every
    world size, balanced to one block, the whole curve for one rank, the caller's order when spatial ordering is off"""
    from rescan_b200 import pipeline, posegrid
    rng = np.random.default_rng(4)
    for n in (0, 1, 255, 256, 257, 5000):
        t = rng.uniform(0, 9, (n, 3)).astype(np.float32)
        assert (pipeline.shard_translations(t, 0, 1) == posegrid.spatial_order(t)).all()
        for world in (2, 3, 8):
            parts = [pipeline.shard_translations(t, r, world) for r in range(world)]
            allids = np.concatenate(parts) if parts else np.zeros(0, np.int64)
            assert sorted(allids.tolist()) == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= pipeline.SHARD_BLOCK
            plain = [pipeline.shard_translations(t, r, world, spatial=False) for r in range(world)]
            assert all((np.diff(p) > 0).all() for p in plain if len(p) > 1)


def test_nms_pair_table_gives_the_greedy_result():
    """the claim behind rsgpu_nms's one-launch path (csrc/nms.cu), checked with the oracle on the CPU: every overlap factor the
    greedy loop of mgs_non_maxima_suppresion (pose_proposal.cpp:371-452) can ask for belongs to a pair {i < j : both scores >=
    0.01, centroids >= dist_threshold apart}, the factor is symmetric in the pair, and the loop run from that table keeps what
    the reference's round-by-round loop keeps"""
    from oracle import orcbind as O
    from rescan_b200 import synth
    from tests import common
    scene = common.tiny_scene()
    rng = np.random.default_rng(23)
    thr = 0.2
    for o in scene.objects[:2]:
        n = 14
        props = np.zeros((n, 17), np.float32)
        for j in range(n):
            d = synth.yaw_pose(rng.uniform(-3.1, 3.1), rng.uniform(-0.5, 0.5), rng.uniform(-0.5, 0.5), rng.uniform(-0.02, 0.02))
            props[j, :16] = common.colmajor((d.astype(np.float64) @ o.pose.astype(np.float64)).astype(np.float32))
            props[j, 16] = np.float32(rng.uniform(0.0, 1.0)) if j % 5 else np.float32(-1.0)
        props[3, 16] = props[8, 16]  # a tie: the first index wins
        p3, p1 = o.cloud.pos(3), o.cloud.pos(1)
        cen = O.centroid(o.cloud.pos(0))
        want = O.nms(p3, p1, cen, props, thr)
        # centroid under every pose: msh_mat4_vec3_mul in float, left to right (column-major 4x4)
        cpos = np.zeros((n, 3), np.float32)
        for j in range(n):
            m = props[j, :16]
            for r in range(3):
                s = np.float32(m[r] * cen[0]); s = np.float32(s + np.float32(m[4 + r] * cen[1]))
                s = np.float32(s + np.float32(m[8 + r] * cen[2])); cpos[j, r] = np.float32(s + m[12 + r])

        def dist(a, b):
            v = (cpos[a] - cpos[b]).astype(np.float32)
            s = np.float32(np.float32(np.float32(v[0] * v[0]) + np.float32(v[1] * v[1])) + np.float32(v[2] * v[2]))
            return np.float32(np.sqrt(np.float64(s)))
        table, asked = {}, 0
        for i in range(n):
            for j in range(i + 1, n):
                if props[i, 16] >= 0.01 and props[j, 16] >= 0.01 and not dist(i, j) < thr:
                    table[(i, j)] = O.overlap_factor(p3, p1, props[i, :16], props[j, :16])
        # symmetry of the factor (union box, max of the two counts)
        for (i, j) in list(table)[:6]:
            assert table[(i, j)] == O.overlap_factor(p3, p1, props[j, :16], props[i, :16])
        mark = np.zeros(n, np.int8)
        while (mark == 0).any():
            best, bs = -1, np.float32(-1e9)
            for i in range(n):
                if mark[i] == 0 and props[i, 16] > bs:
                    best, bs = i, props[i, 16]
            mark[best] = 1
            for i in range(n):
                if mark[i]:
                    continue
                if dist(best, i) < thr or props[i, 16] < 0.01:
                    mark[i] = 2
                    continue
                key = (min(best, i), max(best, i))
                assert key in table, "the loop asked for a pair outside the pre-computed set"
                asked += 1
                if table[key] > 0.5:
                    mark[i] = 2
        assert ((mark == 1) == np.asarray(want, bool)).all()
        assert asked > 0 and 0 < (mark == 1).sum() < n


def test_sub_cell_gap_is_a_lower_bound():
    """the pruning bound of radius_search_sub_kernel (csrc/search.cu), restated in numpy float32 on the CPU: the squared gap from
    a query to a 4x4x4 sub-cell's box - shrunk by the margin, low mantissa bits dropped - never exceeds the exact squared
    distance (msh_hash_grid.h:852-855 expression) to any point that sub_key_kernel's float arithmetic assigns to that sub-cell,
    also for points on sub-cell faces and in grids thousands of cells wide (without the margin the same check fails on 1 % of
    the points)"""
    f32 = np.float32
    rng = np.random.default_rng(8)
    worst = 0.0
    for cell, extent in ((0.1, 8.0), (0.2, 40.0), (0.1, 400.0), (0.5, 900.0)):
        mn = rng.uniform(-5, 5, 3).astype(f32)
        n = 20000
        pts = (mn + rng.uniform(0, extent, (n, 3))).astype(f32)
        # a third of the points exactly on sub-cell faces of their axis
        k = n // 3
        sub = f32(cell / 4)
        pts[:k] = (mn + np.round((pts[:k] - mn) / sub).astype(f32) * sub).astype(f32)
        inv_cell_d = 1.0 / np.float64(cell)
        inv_cell_f, cellf = f32(inv_cell_d), f32(cell)
        subf = f32(cellf * f32(0.25))
        W = int(extent / cell) + 2
        margin = f32(cellf * max(f32(1e-3), f32(5e-7) * f32(W)))
        rel = (pts - mn).astype(f32)                                   # float subtraction, like the grid build
        c = np.trunc(rel.astype(np.float64) * inv_cell_d).astype(np.int64)  # the cell of the reference: double product, truncation
        s = np.clip(np.floor(((rel * inv_cell_f).astype(f32) - c.astype(f32)).astype(f32) * f32(4)).astype(np.int64), 0, 3)
        # queries near the points
        q = (pts + rng.uniform(-cell, cell, (n, 3))).astype(f32)
        qrel = (q - mn).astype(f32)
        lo = ((c.astype(f32) * cellf).astype(f32) + (s.astype(f32) * subf).astype(f32)).astype(f32)
        d = np.maximum(np.maximum((lo - qrel).astype(f32), (qrel - (lo + subf).astype(f32)).astype(f32)) - margin, f32(0)).astype(f32)
        gap = ((d[:, 0] * d[:, 0]).astype(f32) + (d[:, 1] * d[:, 1]).astype(f32)).astype(f32)
        gap = (gap + (d[:, 2] * d[:, 2]).astype(f32)).astype(f32)
        gap_bits = gap.view(np.uint32) & np.uint32(0xFFFFFF00)
        v = (pts - q).astype(f32)
        d2 = (((v[:, 0] * v[:, 0]).astype(f32) + (v[:, 1] * v[:, 1]).astype(f32)).astype(f32) + (v[:, 2] * v[:, 2]).astype(f32)).astype(f32)
        assert (gap_bits <= d2.view(np.uint32)).all(), (cell, extent, int((gap_bits > d2.view(np.uint32)).sum()))
        worst = max(worst, float((gap_bits.view(f32) / np.maximum(d2, f32(1e-30))).max()))
    assert worst <= 1.0

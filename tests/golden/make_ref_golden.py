"""Generate tests/golden/ref_factors.npz by running the REFERENCE's own factor functors.

`make -C oracle ref` compiles /root/reference/src/lvio_fusion/include/lvio_fusion/ceres/{visual,lidar,pose}_error.hpp (in
place, never copied) against the stand-in third-party headers in oracle/ref_compat and the oracle's dual numbers; this script
feeds it seeded KITTI-shaped inputs and stores inputs + residuals + Jacobians.  The oracle (tests/test_ref_golden.py, CPU)
and the CUDA path (same file, -m gpu) must reproduce them.  Only runs where /root/reference is mounted.

    python tests/golden/make_ref_golden.py
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from lvio_fusion_b200 import synth  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
N = 24
N_IMU = 12


def _unit_pose(rng, P, sigma_q=0.01, sigma_t=0.05, scale=1.0):
    q = P[:4] + rng.normal(0, sigma_q, 4)
    return np.concatenate([q / np.linalg.norm(q) * scale, P[4:] + rng.normal(0, sigma_t, 3)])


def build_cases(seed=synth.SEED):
    rng = np.random.default_rng(seed)
    d = synth.make_ba_problem(6, 120, with_imu=False, seed=seed)
    P = d["poses"]
    c0, i0 = d["factors"][0]; c1, i1 = d["factors"][1]; c2, i2 = d["factors"][2]
    cases = {"cameras": np.asarray(d["cameras"], dtype=np.float64)}
    # a1: consts (first_ob, ob, w), params rho, T1, T2 -- one pose pair with a non-unit quaternion (QuaternionRotatePoint normalises)
    sel = rng.choice(len(c0), N, replace=False)
    T1 = np.stack([_unit_pose(rng, P[i0[s, 1]], scale=1.0 + 0.01 * (k % 3)) for k, s in enumerate(sel)])
    T2 = np.stack([_unit_pose(rng, P[i0[s, 2]]) for s in sel])
    cases["tf_c"] = c0[sel]; cases["tf_x"] = np.concatenate([d["rho"][i0[sel, 0]][:, None], T1, T2], axis=1)
    sel = rng.choice(len(c1), N, replace=False)
    cases["po_c"] = c1[sel]; cases["po_x"] = np.stack([_unit_pose(rng, P[i1[s, 0]]) for s in sel])
    sel = rng.choice(len(c2), N, replace=False)
    cases["tc_c"] = c2[sel]; cases["tc_x"] = d["rho"][i2[sel, 0]][:, None] * rng.uniform(0.9, 1.1, (N, 1))
    # a5: mode, p, pa, pb, pc, Twc1, rpyxyz, w
    lid = []
    for k in range(N):
        pa = rng.normal(0, 8, 3); pb = pa + rng.normal(0, 0.4, 3); pc = pa + rng.normal(0, 0.4, 3)
        p = rng.normal(0, 8, 3)
        T = _unit_pose(rng, P[k % len(P)], sigma_q=0.05, sigma_t=1.0)
        e = np.concatenate([rng.normal(0, 0.05, 3), rng.normal(0, 0.3, 3)])
        lid.append(np.concatenate([[k % 2], p, pa, pb, pc, T, e, [rng.uniform(0.01, 1.0)]]))
    cases["lidar"] = np.asarray(lid)
    # a6
    pg, pe, pr = [], [], []
    for k in range(N):
        a, b = _unit_pose(rng, P[k % 5]), _unit_pose(rng, P[k % 5 + 1])
        x1, x2 = _unit_pose(rng, P[k % 5], 0.02, 0.1, 1.0 + 0.005 * (k % 2)), _unit_pose(rng, P[k % 5 + 1], 0.02, 0.1)
        pg.append(np.concatenate([a, b, [rng.uniform(1, 100), rng.uniform(0, 2)], x1, x2]))
        pe.append(np.concatenate([a, [rng.uniform(1, 100), rng.uniform(0, 2)], x1]))
        pr.append(np.concatenate([[k % 2], rng.normal(0, 0.3, 6), [rng.uniform(1, 100)], rng.normal(0, 0.3, 3)]))
    cases["pg"] = np.asarray(pg); cases["pe"] = np.asarray(pe); cases["pr"] = np.asarray(pr)
    # a4: noise4, then per case: ba bg acc0 gyr0 n_samples samples[n][7] | pose_i v_i ba_i bg_i pose_j v_j ba_j bg_j
    cases["imu_noise"] = np.array(synth.IMU_NOISE, dtype=np.float64)
    imu = []
    for k in range(N_IMU):
        ns = 6 + (k % 7)
        ba, bg = rng.normal(0, 0.05, 3), rng.normal(0, 0.01, 3)
        acc = rng.normal(0, 0.8, (ns + 1, 3)) + [0.3, 0.1, 9.81]; gyr = rng.normal(0, 0.15, (ns + 1, 3))
        dt = rng.uniform(0.008, 0.012, ns)
        T = dt.sum()
        pose_i = _unit_pose(rng, P[k % len(P)], 0.05, 0.5)
        v_i = np.array([9.5, 0.3, -0.1]) + rng.normal(0, 0.3, 3)
        pose_j = _unit_pose(rng, np.concatenate([pose_i[:4], pose_i[4:] + v_i * T]), 0.02, 0.03)
        v_j = v_i + rng.normal(0, 0.2, 3)
        x = np.concatenate([pose_i, v_i, ba + rng.normal(0, 0.01, 3), bg + rng.normal(0, 0.002, 3), pose_j, v_j, ba + rng.normal(0, 0.01, 3), bg + rng.normal(0, 0.002, 3)])
        imu.append({"ba": ba, "bg": bg, "acc0": acc[0], "gyr0": gyr[0], "samples": np.concatenate([dt[:, None], acc[1:], gyr[1:]], axis=1), "x": x})
    cases["imu_first"] = np.concatenate([[0], np.cumsum([len(c["samples"]) for c in imu])]).astype(np.int32)
    cases["imu_samples"] = np.concatenate([c["samples"] for c in imu])
    for key in ("ba", "bg", "acc0", "gyr0", "x"):
        cases["imu_" + key] = np.stack([c[key] for c in imu])
    return cases


def run_reference(cases):
    subprocess.check_call(["make", "-s", "-j4", "-C", os.path.join(ROOT, "oracle"), "ref"])
    flat = [cases["cameras"], np.array([N] * 7 + [N_IMU], dtype=np.float64)]
    flat += [np.concatenate([cases["tf_c"], cases["tf_x"]], axis=1).ravel(), np.concatenate([cases["po_c"], cases["po_x"]], axis=1).ravel(),
             np.concatenate([cases["tc_c"], cases["tc_x"]], axis=1).ravel(), cases["lidar"].ravel(), cases["pg"].ravel(), cases["pe"].ravel(), cases["pr"].ravel(), cases["imu_noise"]]
    for k in range(N_IMU):
        lo, hi = cases["imu_first"][k], cases["imu_first"][k + 1]
        flat += [cases["imu_ba"][k], cases["imu_bg"][k], cases["imu_acc0"][k], cases["imu_gyr0"][k], np.array([hi - lo], dtype=np.float64), cases["imu_samples"][lo:hi].ravel(), cases["imu_x"][k]]
    with tempfile.TemporaryDirectory() as td:
        np.concatenate(flat).astype(np.float64).tofile(os.path.join(td, "cases.bin"))
        subprocess.check_call([os.path.join(ROOT, "oracle", "_ref", "ref_factors"), os.path.join(td, "cases.bin"), os.path.join(td, "out.bin")])
        o = np.fromfile(os.path.join(td, "out.bin"), dtype=np.float64)
    out, pos = {}, 0
    for name, R, C in (("tf", 2, 15), ("po", 2, 7), ("tc", 2, 1), ("lidar", 1, 3), ("pg", 6, 14), ("pe", 6, 7), ("pr", 3, 3)):
        blk = o[pos:pos + N * (R + R * C)].reshape(N, R + R * C); pos += N * (R + R * C)
        out[name + "_r"] = blk[:, :R].copy(); out[name + "_J"] = blk[:, R:].reshape(N, R, C).copy()
    per = 17 + 225 + 225 + 15 + 105 + 45 * 3 + 105 + 45 * 3 + 2 * (15 + 105 + 45 * 3 + 105 + 45)
    blk = o[pos:pos + N_IMU * per].reshape(N_IMU, per); pos += N_IMU * per
    out["imu_record"] = blk[:, :467].copy()                      # LVB_IMU layout without the two prior slots
    out["imu_r"] = blk[:, 467:482].copy()
    sizes, q, Js = [7, 3, 3, 3, 7, 3, 3, 3], 482, []
    for w in sizes:
        Js.append(blk[:, q:q + 15 * w].reshape(N_IMU, 15, w)); q += 15 * w
    out["imu_J"] = np.concatenate(Js, axis=2)                      # [n, 15, 32] in the block order of imu_error.hpp:12
    out["imuinit_r"] = blk[:, q:q + 15].copy(); q += 15
    Js = []
    for w in [7, 3, 3, 3, 7, 3]:
        Js.append(blk[:, q:q + 15 * w].reshape(N_IMU, 15, w)); q += 15 * w
    out["imuinit_J"] = np.concatenate(Js, axis=2)                  # [n, 15, 26], ImuInitError with priors 1e8 / 1e8
    out["imuinit2_r"] = blk[:, q:q + 15].copy(); q += 15
    Js = []
    for w in [7, 3, 3, 3, 7, 3]:
        Js.append(blk[:, q:q + 15 * w].reshape(N_IMU, 15, w)); q += 15 * w
    out["imuinit2_J"] = np.concatenate(Js, axis=2)                 # the reference's real priors 1e4 / 1e2 (initializer.cpp:62): indefinite cov^-1
    assert q == per
    assert pos == len(o)
    return out


LIDAR_CFG = dict(num_scans=64, horizon_scan=450, ang_res_y=0.427, ang_bottom=24.9, ground_rows=60, min_range=5.0, max_range=30.0)


def run_reference_lidar(scan):
    """ImageProjection::Process (src/projection.cpp) + filter_points_by_distance (utility.h) of the reference on a raw sweep."""
    c = LIDAR_CFG
    with tempfile.TemporaryDirectory() as td:
        with open(os.path.join(td, "in.bin"), "wb") as f:
            np.array([c["num_scans"], c["horizon_scan"], c["ang_res_y"], c["ang_bottom"], c["ground_rows"], c["min_range"], c["max_range"]], dtype=np.float64).tofile(f)
            np.array([len(scan)], dtype=np.int32).tofile(f)
            np.ascontiguousarray(scan[:, :3], dtype=np.float32).tofile(f)
        subprocess.check_call([os.path.join(ROOT, "oracle", "_ref", "ref_lidar"), os.path.join(td, "in.bin"), os.path.join(td, "out.bin")])
        raw = open(os.path.join(td, "out.bin"), "rb").read()
    m = int(np.frombuffer(raw[:4], dtype=np.int32)[0]); off = 4
    R = c["num_scans"]
    out = {}
    out["points"] = np.frombuffer(raw[off:off + 16 * m], dtype=np.float32).reshape(m, 4).copy(); off += 16 * m
    out["range"] = np.frombuffer(raw[off:off + 4 * m], dtype=np.float32).copy(); off += 4 * m
    out["ground"] = np.frombuffer(raw[off:off + m], dtype=np.uint8).copy(); off += m
    out["col"] = np.frombuffer(raw[off:off + 4 * m], dtype=np.int32).copy(); off += 4 * m
    out["start_ring"] = np.frombuffer(raw[off:off + 4 * R], dtype=np.int32).copy(); off += 4 * R
    out["end_ring"] = np.frombuffer(raw[off:off + 4 * R], dtype=np.int32).copy(); off += 4 * R
    out["orientation"] = np.frombuffer(raw[off:off + 12], dtype=np.float32).copy(); off += 12
    assert off == len(raw)
    return out


ASSOC_EXT = [0.01, -0.02, 0.005, 1.0, 0.27, 0.0, 0.08]
W_VISUAL, N_FEATURES_LEFT = 71.8856, 120


def _assoc_cfg(horizon, ext):
    return np.array([64, horizon, 0.427, 24.9, 60, 0.1036, 5, 30, 0.2, 0], dtype=np.float64).tobytes() + np.asarray(ext, dtype=np.float64).tobytes()


def _read_cloud(raw, off):
    m = int(np.frombuffer(raw[off:off + 4], dtype=np.int32)[0]); off += 4
    return np.frombuffer(raw[off:off + 16 * m], dtype=np.float32).reshape(m, 4).copy(), off + 16 * m


def run_reference_extract(scan, horizon):
    """FeatureAssociation::Preprocess .. ExtractFeatures (association.cpp:88-268) of the reference on a raw sweep."""
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_assoc")
    with tempfile.TemporaryDirectory() as td:
        with open(os.path.join(td, "in.bin"), "wb") as f:
            f.write(_assoc_cfg(horizon, ASSOC_EXT)); np.array([len(scan)], dtype=np.int32).tofile(f); np.ascontiguousarray(scan[:, :3], dtype=np.float32).tofile(f)
        subprocess.check_call([exe, "extract", os.path.join(td, "in.bin"), os.path.join(td, "out.bin")])
        raw = open(os.path.join(td, "out.bin"), "rb").read()
    seg, off = _read_cloud(raw, 0)
    curv = np.frombuffer(raw[off:off + 4 * len(seg)], dtype=np.float32).copy(); off += 4 * len(seg)
    ground, off = _read_cloud(raw, off); surf, off = _read_cloud(raw, off)
    assert off == len(raw)
    return {"seg": seg, "curvature": curv, "ground": ground, "surf": surf}


def run_reference_scan2map(sc, mode, e0):
    """FeatureAssociation::ScanToMapWithGround / ScanToMapWithSegmented (association.cpp:270-384) of the reference."""
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_assoc")
    scan = np.ascontiguousarray(sc["scan"][:, :4], dtype=np.float32); mp = np.ascontiguousarray(sc["map"][:, :4], dtype=np.float32)
    w = [W_VISUAL, sc["weight"] if mode == 0 else 1.0, sc["weight"] if mode == 1 else 0.01]
    with tempfile.TemporaryDirectory() as td:
        with open(os.path.join(td, "in.bin"), "wb") as f:
            f.write(_assoc_cfg(1800, [0, 0, 0, 1, 0, 0, 0]))
            np.array([mode], dtype=np.int32).tofile(f)
            for a in (sc["frame_pose"], sc["map_pose"], e0, w):
                np.asarray(a, dtype=np.float64).tofile(f)
            np.array([N_FEATURES_LEFT, 0], dtype=np.int32).tofile(f)
            np.array([len(scan)], dtype=np.int32).tofile(f); scan.tofile(f); np.array([len(mp)], dtype=np.int32).tofile(f); mp.tofile(f)
        subprocess.check_call([exe, "scan2map", os.path.join(td, "in.bin"), os.path.join(td, "out.bin")])
        o = np.fromfile(os.path.join(td, "out.bin"), dtype=np.float64)
    return o[:4].copy(), o[4:].reshape(len(scan), 5).copy()


def assoc_case(kind):
    """Scan / map feature clouds with a share of scan points pushed away from every map surface, so that the gate rejects."""
    sc = synth.make_icp_problem(1200, 9000, seed=synth.SEED + (7 if kind == "ground" else 8), kind=kind)
    rng = np.random.default_rng(99)
    scan = np.array(sc["scan"], dtype=np.float32, copy=True)
    far = rng.choice(len(scan), len(scan) // 6, replace=False)
    scan[far, :3] += rng.normal(0, 1.5, (len(far), 3)).astype(np.float32)
    sc = dict(sc); sc["scan"] = scan
    return sc


def main():
    if not os.path.isdir("/root/reference"):
        raise SystemExit("the reference tree is not mounted here; the committed fixture stays as it is")
    cases = build_cases()
    out = run_reference(cases)
    np.savez_compressed(os.path.join(HERE, "ref_factors.npz"), **cases, **out)
    print("reference golden written:", {k: v.shape for k, v in out.items()})
    scan = synth.make_lidar_scan(seed=synth.SEED + 5, horizon_scan=LIDAR_CFG["horizon_scan"], n_boxes=10)
    lid = run_reference_lidar(scan)
    np.savez_compressed(os.path.join(HERE, "ref_lidar.npz"), scan=scan[:, :3].astype(np.float32), **lid)
    print("reference lidar golden written:", len(lid["points"]), "segmented points of", len(scan))
    ext = run_reference_extract(scan, LIDAR_CFG["horizon_scan"])
    out = {"scan": scan[:, :3].astype(np.float32), "ext_seg": ext["seg"], "ext_curvature": ext["curvature"], "ext_ground": ext["ground"], "ext_surf": ext["surf"]}
    for kind, mode in (("ground", 0), ("surf", 1)):
        sc = assoc_case(kind)
        e0 = synth.relative_rpyxyz(sc["map_pose"], sc["frame_pose"])
        head, tab = run_reference_scan2map(sc, mode, e0)
        out.update({kind + "_head": head, kind + "_table": tab, kind + "_e0": np.asarray(e0)})
        print("reference scan-to-map (%s): %d blocks of %d scan points, prior weight %.3f, huber %.2f" % (kind, int(head[0]), len(tab), head[2], head[3]))
    np.savez_compressed(os.path.join(HERE, "ref_assoc.npz"), **out)


if __name__ == "__main__":
    main()

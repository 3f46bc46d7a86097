#!/usr/bin/env python
"""bench.py -- headline measurement of the hot path (see DESIGN.md "Measurement").

A *step* is one Levenberg-Marquardt iteration of the sliding-window BA (fused evaluate + J^T J
assembly, Schur elimination, reduced Cholesky, back-substitution, candidate cost, LM decision) on the
synthetic KITTI-shaped window of BASELINE.json configs[1] (stereo+IMU, 10 keyframes, 4000 landmarks per
GPU).  `value` = Jacobian-carrying residual rows processed per second with the problem resident in HBM;
`e2e` = the same through the public host API (problem upload from host buffers, K iterations, result
download).  Extra objects: `icp` (configs[2]: 120k-point scan vs 1M-point map, points/s), `roofline`
(the TwoFrame Jacobian-eval kernel at configs[4] scale, algorithmic bytes / CUDA-event time / measured HBM
peak), `cpu_baseline` (the oracle restatement of the Ceres path timed on the host cores).

    python bench.py --gpus N --steps K --warmup W            # torchrun for N > 1
    python bench.py --impl reference ...                       # CPU oracle arm (rank 0 only)
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from lvio_fusion_b200 import backend, synth  # noqa: E402

N_KF, N_LM = 10, 4000                # configs[1]
ICP_K, ICP_P = 120000, 1000000       # configs[2]
EVAL_KF, EVAL_LM = 5000, 500000      # configs[4] scale for the eval-kernel roofline
# dram__bytes_read.sum + dram__bytes_write.sum of one ba_eval_two_frame_kernel launch at that scale, from the
# `ncu --set full` capture summarised in profiles/eval_r1_v3_summary.txt (75.19 MB + 293.14 MB)
EVAL_DRAM_BYTES_PER_LAUNCH = 368.33e6
BYTES_TWO_FRAME = 308                # 40 const + 12 idx + 16 r + 240 J


def cpu_threads():
    n = os.cpu_count() or 1
    return min(8, max(1, int(0.75 * n)))       # the reference's own rule, estimator.cpp:10


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f)["hbm_gbs"], "MEASURED_PEAKS.json"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    def __init__(self, device):
        super().__init__(daemon=True)
        self.device, self.rows, self.stop_flag = device, [], False

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.rows[0][1]) if self.rows[0][1].replace(".", "").isdigit() else None,
                "reasons": reasons, "samples": len(self.rows)}


def bench_options(api, iters, threads=1):
    # fixed iteration count: tolerances off so both arms do exactly the same number of LM iterations
    return backend.default_options(api, max_num_iterations=iters, function_tolerance=0.0, gradient_tolerance=0.0,
                                   parameter_tolerance=0.0, num_threads=threads)


def run_solves(problem, d, api, total_iters, per_solve, threads=1):
    """Run LM iterations in solves of `per_solve` iterations restarted from the initial guess until
    `total_iters` iterations have been executed; returns iterations actually executed."""
    done = 0
    while done < total_iters:
        k = min(per_solve, total_iters - done)
        problem.update_params(d["poses"], d["vec3"], d["rho"])
        s = problem.solve(bench_options(api, k, threads))
        done += max(1, s.num_iterations)
    return done


def reference_arm(args):
    """CPU arm: the oracle restatement of the reference's Ceres+PCL path on the host cores."""
    from oracle import binding
    orc = binding.load()
    T = cpu_threads()
    octx = backend.Context(orc)
    d = synth.make_ba_problem(N_KF, N_LM, with_imu=True, seed=synth.SEED)
    rows = synth.count_rows(d)
    p = backend.Problem.from_dict(octx, d)
    per = 10
    run_solves(p, d, orc, max(1, args.warmup), per, T)
    t0 = time.perf_counter()
    it = run_solves(p, d, orc, args.steps, per, T)
    dt = time.perf_counter() - t0
    val = rows * it / dt
    # ICP sample: 1/10 of the queries against the full map (kd-tree build included, as the reference rebuilds it per call)
    sc = synth.make_icp_problem(ICP_K // 10, ICP_P, seed=synth.SEED, kind="surf")
    fo = backend.FeatureAssociation(octx)
    orc.icp_set_threads(fo.h, T)
    e0 = synth.relative_rpyxyz(sc["map_pose"], sc["frame_pose"])
    t0 = time.perf_counter()
    fo.set_map(sc["map"], sc["cell_size"])
    fo.scan_to_map(sc["mode"], sc["scan"], sc["frame_pose"], sc["map_pose"], e0, sc["weight"], -1.0, sc["huber_a"], sc["thr"])
    icp_dt = time.perf_counter() - t0
    line = {
        "impl": "reference", "metric": "ba_residual_jacobian_rows_per_s", "value": val, "unit": "rows/s", "n_gpus": args.gpus,
        "steps": it, "warmup": args.warmup, "ms_per_step": 1e3 * dt / it, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "configs[1]: KITTI-shaped stereo+IMU 10-keyframe window, %d landmarks, %d residual rows; step = one LM iteration" % (N_LM, rows)},
        "cpu_baseline": {"value": val, "unit": "rows/s", "cores": T, "kind": "port",
                         "sample": "%d LM iterations of the full configs[1] window (oracle restatement of Ceres SPARSE_SCHUR, %d threads); the reference itself cannot be built here" % (it, T)},
        "icp": {"metric": "icp_points_per_s", "value": (ICP_K // 10) / icp_dt, "unit": "points/s",
                "sample": "%d queries vs %d-point map, kd-tree build + 3-NN + 4 LM iterations, %d threads" % (ICP_K // 10, ICP_P, T)},
        "e2e": {"value": val, "unit": "rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--skip-icp", action="store_true")
    ap.add_argument("--skip-roofline", action="store_true")
    ap.add_argument("--skip-global", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if rank == 0:
            reference_arm(args)
        return

    import torch
    import torch.distributed as dist
    from lvio_fusion_b200 import _capi
    lvb = _capi.load()
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        ctx = backend.Context(lvb, device=local_rank, stream=stream.cuda_stream)
        if world > 1:
            import ctypes
            uid = ctypes.create_string_buffer(128)
            if rank == 0:
                lvb.check(lvb.comm_unique_id(uid), "comm_unique_id")
            box = [uid.raw]
            dist.broadcast_object_list(box, src=0)
            ctx.comm_init(rank, world, box[0])

        def barrier():
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()

        def max_over_ranks(ms):
            if world == 1:
                return ms
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())

        # ---------------- BA: weak scaling, N_LM landmarks per GPU, sharded by landmark
        full = synth.make_ba_problem(N_KF, N_LM * world, with_imu=True, seed=synth.SEED)
        total_rows = synth.count_rows(full)
        d = synth.shard_ba_problem(full, rank, world) if world > 1 else full
        prob = backend.Problem.from_dict(ctx, d)
        per = 10
        run_solves(prob, d, lvb, max(3, args.warmup), per)
        sampler = ClockSampler(local_rank)
        sampler.start()
        barrier()
        l0 = ctx.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        iters = run_solves(prob, d, lvb, args.steps, per)
        e1.record(stream)
        barrier()
        ms = max_over_ranks(e0.elapsed_time(e1))
        launches = ctx.launch_count() - l0
        value = total_rows * iters / (ms * 1e-3)

        # ---------------- e2e: build + upload + solve + download through the public API, host buffers
        barrier()
        t0 = time.perf_counter()
        n_e2e, it_e2e, h2d, d2h = 0, 0, 0, 0
        while it_e2e < args.steps:
            p2 = backend.Problem.from_dict(ctx, d)
            s2 = p2.solve(bench_options(lvb, per))
            P = p2.poses(); p2.vec3(); p2.inv_depths()
            it_e2e += max(1, s2.num_iterations); n_e2e += 1
            h2d = sum(f[0].nbytes + f[1].nbytes for f in d["factors"].values()) + d["poses"].nbytes + d["vec3"].nbytes + d["rho"].nbytes + 22 * 8
            d2h = d["poses"].nbytes + d["vec3"].nbytes + d["rho"].nbytes
            p2.close()
        barrier()
        e2e_s = time.perf_counter() - t0
        e2e_ms = max_over_ranks(e2e_s * 1e3)
        e2e_value = total_rows * it_e2e / (e2e_ms * 1e-3)

        line = {
            "metric": "ba_residual_jacobian_rows_per_s", "value": value, "unit": "rows/s", "n_gpus": world, "steps": iters, "warmup": args.warmup,
            "ms_per_step": ms / iters, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "configs[1]: KITTI-shaped stereo+IMU 10-keyframe window, %d landmarks per GPU (sharded by landmark, one in-kernel all-reduce of the reduced system over NVLink peer memory per iteration), %d residual rows total; step = one LM iteration" % (N_LM, total_rows),
                       "blocks": synth.count_blocks(full), "iters_per_solve": per,
                       "l2": "BA working set (~3 MB) is L2-resident by design; the roofline leg streams 394 MB per launch (> 126 MB L2)"},
            "e2e": {"value": e2e_value, "unit": "rows/s", "h2d_bytes_per_step": int(h2d / per), "d2h_bytes_per_step": int(d2h / per),
                    "note": "Problem.from_dict (AoS->SoA, H2D) + solve(%d iterations) + D2H of all parameter blocks, per solve; bytes amortised per iteration" % per},
            "gpu_launches": int(launches),
        }

        # ---------------- roofline of the dominant streaming kernel: TwoFrame Jacobian eval at configs[4] scale
        # rank-0-only legs run on their own world-1 context: on the communicator-attached one Problem.from_dict is a collective
        # (lvb_ba_finalize agrees on the unknown order and the envelope over the ranks) and would pair with nothing
        ctx1 = backend.Context(lvb, device=local_rank, stream=stream.cuda_stream) if world > 1 else ctx
        if rank == 0 and not args.skip_roofline:
            big = synth.make_ba_problem(EVAL_KF, EVAL_LM, with_imu=False, seed=synth.SEED)
            big["factors"] = {0: big["factors"][0]}
            pb = backend.Problem.from_dict(ctx1, big)
            nb = len(big["factors"][0][0])
            for _ in range(3):
                pb.evaluate_device(0)
            torch.cuda.synchronize()
            reps = 10
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
            ev[0].record(stream)
            for i in range(reps):
                pb.evaluate_device(0)
                ev[i + 1].record(stream)
            torch.cuda.synchronize()
            times = [ev[i].elapsed_time(ev[i + 1]) for i in range(reps)]
            t_avg = sum(times) / reps
            peak, src = peaks()
            ach = nb * BYTES_TWO_FRAME / (t_avg * 1e-3) / 1e9
            line["roofline"] = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": EVAL_DRAM_BYTES_PER_LAUNCH,
                                "kernel": "ba_eval_two_frame_kernel", "blocks": nb, "bytes_per_block": BYTES_TWO_FRAME,
                                "us_per_launch": t_avg * 1e3, "peak_source": src, "workload": "configs[4]-scale: %d keyframes, %d landmarks" % (EVAL_KF, EVAL_LM)}
            pb.close()
            line["roofline"]["rows_per_s"] = 2.0 * nb / (t_avg * 1e-3)

        # ---------------- configs[4]: map-scale global BA (banded reduced system), landmarks sharded over the ranks
        if not args.skip_global:
            g_full = synth.make_ba_problem(EVAL_KF, EVAL_LM, with_imu=True, seed=synth.SEED + 1)
            g_rows = synth.count_rows(g_full)
            gd = synth.shard_ba_problem(g_full, rank, world) if world > 1 else g_full
            del g_full
            gp = backend.Problem.from_dict(ctx, gd)
            gp.solve(bench_options(lvb, 3))
            gp.update_params(gd["poses"], gd["vec3"], gd["rho"])
            barrier()
            e0.record(stream)
            gs = gp.solve(bench_options(lvb, 5))
            e1.record(stream)
            barrier()
            g_ms = max_over_ranks(e0.elapsed_time(e1)) / max(1, gs.num_iterations)
            line["global_ba"] = {"metric": "ba_residual_jacobian_rows_per_s", "value": g_rows / (g_ms * 1e-3), "unit": "rows/s", "ms_per_iteration": g_ms,
                                 "iterations": gs.num_iterations, "rows": g_rows, "camera_dims": gp.dims()[0], "scaling": "strong",
                                 "cost": [gs.initial_cost, gs.final_cost],
                                 "workload": "configs[4]-scale: %d keyframes + IMU, %d landmarks, sharded by landmark; banded reduced camera system, envelope Cholesky" % (EVAL_KF, EVAL_LM)}
            gp.close()
            del gd

        # ---------------- ICP (configs[2]): points/s through scan_to_map, map resident vs e2e with set_map
        if not args.skip_icp:
            sc = synth.make_icp_problem(ICP_K, ICP_P, seed=synth.SEED, kind="surf")
            lo, hi = rank * ICP_K // world, (rank + 1) * ICP_K // world     # query tiles sharded, map replicated
            scan = sc["scan"][lo:hi]
            fa = backend.FeatureAssociation(ctx)
            e_init = synth.relative_rpyxyz(sc["map_pose"], sc["frame_pose"])
            fa.set_map(sc["map"], sc["cell_size"])
            argsicp = (sc["mode"], scan, sc["frame_pose"], sc["map_pose"], e_init, sc["weight"], -1.0, sc["huber_a"], sc["thr"])
            for _ in range(3):
                fa.scan_to_map(*argsicp)
            barrier()
            reps = 10
            e0.record(stream)
            for _ in range(reps):
                fa.scan_to_map(*argsicp)
            e1.record(stream)
            barrier()
            icp_ms = max_over_ranks(e0.elapsed_time(e1)) / reps
            t0 = time.perf_counter()
            for _ in range(3):
                fa.set_map(sc["map"], sc["cell_size"])
                fa.scan_to_map(*argsicp)
            barrier()
            icp_e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3 / 3)
            line["icp"] = {"metric": "icp_points_per_s", "value": ICP_K / (icp_ms * 1e-3), "unit": "points/s", "ms_per_scan": icp_ms,
                           "e2e": {"value": ICP_K / (icp_e2e_ms * 1e-3), "unit": "points/s", "ms_per_scan": icp_e2e_ms,
                                   "note": "map upload + voxel build + scan upload + association + 4 LM iterations"},
                           "workload": "configs[2]: %d-point scan vs %d-point map, surf gate, Huber 0.1" % (ICP_K, ICP_P)}

        # ---------------- lidar feature pipeline: raw 64 x 1800 sweep -> ground / surf features, host buffers in and out
        if rank == 0 and not args.skip_icp:
            sweep = synth.make_lidar_scan(seed=synth.SEED)
            lf = backend.LidarFeatures(ctx1)
            for _ in range(3):
                gcl, scl = lf.extract(sweep)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(10):
                lf.extract(sweep)
            lf_ms = (time.perf_counter() - t0) * 1e3 / 10
            line["lidar_features"] = {"metric": "lidar_raw_points_per_s", "value": len(sweep) / (lf_ms * 1e-3), "unit": "points/s", "ms_per_sweep": lf_ms,
                                      "ground_points": int(len(gcl)), "surf_points": int(len(scl)),
                                      "workload": "FeatureAssociation::Process on a synthetic %d-point sweep (64 x 1800), e2e with H2D/D2H" % len(sweep)}

        sampler.stop_flag = True
        sampler.join(timeout=2)
        line["clocks"] = sampler.summary()

        # ---------------- CPU baseline beside it (rank 0, bounded sample)
        if rank == 0:
            from oracle import binding
            orc = binding.load()
            T = cpu_threads()
            octx = backend.Context(orc)
            dc = synth.make_ba_problem(N_KF, N_LM, with_imu=True, seed=synth.SEED)
            po = backend.Problem.from_dict(octx, dc)
            run_solves(po, dc, orc, 5, per, T)
            t0 = time.perf_counter()
            itc = run_solves(po, dc, orc, 40, per, T)
            dtc = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": synth.count_rows(dc) * itc / dtc, "unit": "rows/s", "cores": T, "kind": "port",
                                    "sample": "%d LM iterations of the configs[1] window on the oracle restatement (%d threads of %d cores)" % (itc, T, os.cpu_count() or 1)}
            # the roofline kernel's CPU counterpart: the reference's own TwoFrameReprojectionError under forward-mode duals (what Ceres'
            # AutoDiff does per residual block), prebuilt where the reference tree is mounted (oracle/ref_time_harness.cpp); one thread, ~2 s
            ref_time = os.path.join(ROOT, "oracle", "_ref", "ref_time")
            if os.path.exists(ref_time):
                try:
                    t = subprocess.run([ref_time, "20000", "2.0"], capture_output=True, text=True, timeout=60, check=True).stdout.split()
                    line["cpu_baseline"]["eval_kernel"] = {"value": 2.0 * float(t[2]), "unit": "rows/s", "cores": 1, "kind": "reference",
                                                           "sample": "%s TwoFrameReprojectionError blocks (visual_error.hpp:78-107 compiled in place, residual + 2x15 Jacobian by duals) in %s s" % (t[0], t[1])}
                except Exception as exc:                      # a baseline beside the number, never a reason to lose the line
                    line["cpu_baseline"]["eval_kernel"] = {"unavailable": str(exc)[:120]}
            if "lidar_features" in line:
                olf = backend.LidarFeatures(octx)
                t0 = time.perf_counter()
                for _ in range(5):
                    olf.extract(sweep)
                line["lidar_features"]["cpu_port_points_per_s"] = len(sweep) / ((time.perf_counter() - t0) / 5)
            print(json.dumps(line))
    if world > 1:
        dist.barrier()                  # rank 0 is still timing the CPU baseline: nobody tears the process group down under it
        torch.cuda.synchronize()
        ctx.close()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

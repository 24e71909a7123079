#!/usr/bin/env python
"""bench.py -- VO front-end frames/sec on B200 (BASELINE.json metric), one JSON line on stdout.

    python bench.py --gpus N --steps K --warmup W            # our arm (libmonovo_b200.so, CUDA)
    python bench.py --impl reference --gpus N --steps K ...   # the reference's own OpenCV CPU path (cv2)

Workload (config.workload): BASELINE.json configs[1] frames -- 1241x376 synthetic KITTI-shaped sequences,
2000 ORB features, the full front-end frame of SURVEY.md 8(d): ORB -> kNN+ratio -> LK (21x21, 4 levels) ->
H + F RANSAC -> E RANSAC -> recoverPose -> triangulate.  A step = one front-end frame for each of the S
independent camera streams of the stream group on this GPU (configs[4] sharding: S = 32 streams per GPU,
weak scaling, no collective on the data path).  The single-stream latency of configs[1] is reported in
`single_stream`.
  value : frames/s with the frames already resident in HBM when the timed region starts
  e2e   : frames/s through the C ABI with HOST (pinned) frames: H2D of every frame and D2H of every
          result record inside the timed region
Timing: CUDA events on the context's stream, barrier + synchronize on both sides, max over ranks.
Between steps every stream moves to its next frame; one step touches S * 10 MB > L2, no cache reuse.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W, H, NFEAT = 1241, 376, 2000
NFRAMES = 6                      # distinct frames per stream, visited ping-pong
METRIC = "VO front-end frames/sec (1241x376, 2000 ORB) per B200; ORB+match ms/frame"


def pingpong(t: int, n: int) -> int:
    period = 2 * (n - 1)
    k = t % period
    return k if k < n else period - k


def level_pixels(w, h):
    from oracle import orb_oracle as oo
    return [a * b for a, b in oo.level_sizes(w, h)]


# ------------------------------------------------------------------------------------------------
def cv2_front_end_frame(cv2, orb, bf, prev, img, K):
    """The reference's OpenCV calls for one front-end frame (call sites: feature_processor.cpp:22,29;
    tracker.cpp:68,243,248; initializer.cpp:228,236,125)."""
    kps, desc = orb.detectAndCompute(img, None)
    out = {"n_keypoints": len(kps)}
    if prev is not None:
        pimg, pkps, pdesc = prev
        knn = bf.knnMatch(pdesc, desc, 2)
        out["n_matches"] = sum(1 for m in knn if len(m) == 2 and m[0].distance < 0.7 * m[1].distance)
        pts = np.array([k.pt for k in pkps], np.float32)
        nxt, st, err = cv2.calcOpticalFlowPyrLK(pimg, img, pts, None)
        ok = (st.ravel() == 1) & (err.ravel() < 30.0)
        p1, p2 = pts[ok], nxt[ok]
        out["n_tracked"] = int(ok.sum())
        Hm, mh = cv2.findHomography(p1, p2, cv2.RANSAC, 1.0)
        F, mf = cv2.findFundamentalMat(p1, p2, cv2.FM_RANSAC, 1.0, 0.99)
        E, me = cv2.findEssentialMat(p1, p2, K, cv2.RANSAC, 0.99, 1.0)
        good, R, t, mp = cv2.recoverPose(E, p1, p2, K, mask=me.copy())
        X = cv2.triangulatePoints(K @ np.eye(3, 4), K @ np.column_stack([R, t]), p1.T.copy(), p2.T.copy())
        X3 = cv2.convertPointsFromHomogeneous(X.T).reshape(-1, 3)
        z2 = (R[2:3] @ X3.T.astype(np.float64)).ravel() + t[2]
        out.update(score_h=int(mh.sum()), score_f=int(mf.sum()), n_inliers_e=int(me.sum()), n_pose_good=int(good),
                   n_triangulated=int(((mp.ravel() != 0) & (X3[:, 2] > 0) & (z2 > 0)).sum()))
    return out, (img, kps, desc)


def cpu_reference_run(frames_by_stream, K, n_frames_budget, seconds_budget):
    """Time the cv2 front-end on the host cores.  Returns (frames/s, frames timed, stage dict, threads)."""
    import cv2
    threads = os.cpu_count() or 1
    cv2.setNumThreads(threads)
    orb = cv2.ORB_create(NFEAT)
    bf = cv2.BFMatcher(cv2.NORM_HAMMING)
    prevs = [None] * len(frames_by_stream)
    # warm-up: first frame of every stream (feature extraction only, like the first GPU step)
    for s, fr in enumerate(frames_by_stream):
        _, prevs[s] = cv2_front_end_frame(cv2, orb, bf, None, fr[0], K)
    done, t = 0, 1
    t0 = time.perf_counter()
    last = None
    while done < n_frames_budget and time.perf_counter() - t0 < seconds_budget:
        for s, fr in enumerate(frames_by_stream):
            last, prevs[s] = cv2_front_end_frame(cv2, orb, bf, prevs[s], fr[pingpong(t, len(fr))], K)
            done += 1
            if done >= n_frames_budget or time.perf_counter() - t0 >= seconds_budget:
                break
        t += 1
    dt = time.perf_counter() - t0
    return done / dt, done, last, threads, cv2.__version__


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clock / power / throttle reasons of one GPU through NVML from a thread (every ~5 ms), so that
    even a sub-second timed region gets samples; falls back to one nvidia-smi query if NVML is unavailable."""

    def __init__(self, index: int):
        self.index = index
        self.samples = []
        self.stop_flag = threading.Event()
        self.th = None
        self.nv = None
        self.h = None

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            self.nv = nv
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = self.index
            if vis:
                try:
                    idx = int(vis.split(",")[self.index])
                except Exception:
                    pass
            self.h = nv.nvmlDeviceGetHandleByIndex(idx)
            self.th = threading.Thread(target=self._loop, daemon=True)
            self.th.start()
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        while not self.stop_flag.is_set():
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                pw = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(
                    nv, "nvmlDeviceGetCurrentClocksEventReasons") else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append((sm, pw, rs))
            except Exception:
                pass
            time.sleep(0.005)

    def stop(self):
        if self.nv is None:
            return self._smi_once()
        nv = self.nv
        self.stop_flag.set()
        self.th.join(timeout=2)
        try:
            mx = nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)
        except Exception:
            mx = None
        bits = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4,
                "hw_power_brake_slowdown": 0x80}
        reasons = sorted(k for k, b in bits.items() if any(r & b for _, _, r in self.samples))
        sm = [x[0] for x in self.samples]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx,
                "power_w_max": max((x[1] for x in self.samples), default=None), "samples": len(sm),
                "reasons": reasons, "source": "nvml, 5 ms period, inside the timed region"}

    def _smi_once(self):
        try:
            q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
                 "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
                 "clocks_event_reasons.sw_power_cap")
            out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                 capture_output=True, text=True, timeout=10).stdout.strip().split(",")
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            return {"sm_mhz": float(out[0]), "sm_max_mhz": float(out[1]), "power_w_max": float(out[2]), "samples": 1,
                    "reasons": [n for n, v in zip(names, out[3:7]) if v.strip().lower().startswith("active")],
                    "source": "nvidia-smi, one query after the timed region"}
        except Exception:
            return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": ["nvml and nvidia-smi unavailable"]}


# ------------------------------------------------------------------------------------------------
def run_reference(args, rank):
    """--impl reference: the reference's own CPU implementation of the path (OpenCV through cv2) on the host."""
    if rank != 0:
        return
    from oracle import synth
    nstreams = 2
    seqs = [synth.synth_sequence(H, W, s, NFRAMES) for s in range(nstreams)]
    K = seqs[0][1]
    frames = [s[0] for s in seqs]
    try:
        import cv2  # noqa: F401
    except Exception as e:  # pragma: no cover
        print(json.dumps({"impl": "reference", "unavailable": f"cv2 not importable on this host: {e}"}))
        return
    per_step = 4                                     # a step = 4 front-end frames (bounded sample of the workload)
    cpu_reference_run(frames, K, args.warmup * per_step, 60.0)
    t0 = time.perf_counter()
    fps, done, last, threads, ver = cpu_reference_run(frames, K, args.steps * per_step, 240.0)
    dt = time.perf_counter() - t0
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(args.steps, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/i32/f32/f64", "data": "synthetic",
        "config": {"workload": "configs[1] frames: 1241x376 synthetic, 2000 ORB, full front-end frame "
                               "(ORB+kNN+LK+H/F/E RANSAC+recoverPose+triangulate); CPU sample of 4 frames per step",
                   "frames_per_step": per_step, "frames_timed": done},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "kind": "reference",
                         "sample": f"{done} front-end frames of 2 streams, cv2 {ver} (the OpenCV calls of the "
                                   f"reference's call sites), cv2.setNumThreads({threads})"},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--streams", type=int, default=32, help="camera streams per GPU (stream group size)")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    from oracle import synth
    from ros2_mono_vo_b200 import Context

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libmonovo_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    args.warmup = max(args.warmup, 3)
    S = args.streams

    # ---- synthetic sequences: stream ids are global so that ranks own disjoint streams -------------
    seqs = [synth.synth_sequence(H, W, rank * S + s, NFRAMES) for s in range(S)]
    K = seqs[0][1]
    host = torch.empty((NFRAMES, S, H, W), dtype=torch.uint8).pin_memory()
    for s in range(S):
        for f in range(NFRAMES):
            host[f, s] = torch.from_numpy(seqs[s][0][f])
    dev = host.cuda()
    stream = torch.cuda.Stream()
    ctx = Context(W, H, nfeatures=NFEAT, batch=S, device=local_rank, cuda_stream=stream.cuda_stream)
    host_np = host.numpy()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_dev(t):
        f = pingpong(t, NFRAMES)
        return ctx.group_step(None, K, device_ptr=dev[f].data_ptr(), shape=(H, W))

    def step_host(t):
        return ctx.group_step(host_np[pingpong(t, NFRAMES)], K)

    def submit_dev(t):
        ctx.group_submit(None, K, device_ptr=dev[pingpong(t, NFRAMES)].data_ptr(), shape=(H, W))

    def submit_host(t):
        ctx.group_submit(host_np[pingpong(t, NFRAMES)], K)

    def timed(t_start, steps, submit_fn=None, step_fn=None):
        """Times `steps` steps with CUDA events on the context's stream.  submit_fn: the pipelined public API
        (submit(t + 1) before collect(t): the upload / launch work of the next step overlaps the kernels of the current
        one; every step's frames and result records still cross PCIe inside the timed region).  step_fn: one
        synchronous mvo_group_step per step (also collects the per-stage event timings)."""
        stage_acc = {}
        barrier()
        sampler = ClockSampler(local_rank)
        sampler.start()
        l0 = ctx.launch_count
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream)
            if submit_fn is not None:
                submit_fn(t_start)
                for i in range(1, steps):
                    submit_fn(t_start + i)
                    res = ctx.group_collect()
                res = ctx.group_collect()
            else:
                for i in range(steps):
                    res = step_fn(t_start + i)
                    for k, v in ctx.stage_ms().items():
                        stage_acc[k] = stage_acc.get(k, 0.0) + v
            e1.record(stream)
        barrier()
        clocks = sampler.stop()
        ms = e0.elapsed_time(e1)
        if world > 1:
            tms = torch.tensor([ms], device="cuda")
            dist.all_reduce(tms, op=dist.ReduceOp.MAX)
            ms = float(tms.item())
        return ms, {k: v / steps for k, v in stage_acc.items()}, ctx.launch_count - l0, clocks, res

    # ---- value: frames resident in HBM ---------------------------------------------------------------
    ctx.group_reset()
    t = 0
    for i in range(args.warmup):
        step_dev(t + i)
    t += args.warmup
    ms_dev, _, launches, clocks, res = timed(t, args.steps, submit_fn=submit_dev)
    t += args.steps
    # per-stage device times (CUDA events inside the library) from synchronous steps; not part of `value`
    ms_dev_sync, stages, _, _, _ = timed(t, min(args.steps, 50), step_fn=step_dev)
    t += min(args.steps, 50)
    # ---- e2e: pinned host frames through the C ABI ----------------------------------------------------
    for i in range(3):
        step_host(t + i)
    t += 3
    ms_host, _, _, _, _ = timed(t, args.steps, submit_fn=submit_host)
    t += args.steps
    ms_host_sync, _, _, _, _ = timed(t, args.steps, step_fn=step_host)

    frames = S * world * args.steps
    value = frames / (ms_dev * 1e-3)
    e2e = frames / (ms_host * 1e-3)

    # ---- single-stream latency (configs[1] as worded: one stream on one B200) -------------------------
    single = None
    if rank == 0:
        c1 = Context(W, H, nfeatures=NFEAT, batch=1, device=local_rank)
        for i in range(5):
            c1.group_step(host_np[pingpong(i, NFRAMES), :1], K)
        t0 = time.perf_counter()
        n1 = 40
        st1 = {}
        for i in range(n1):
            c1.group_step(host_np[pingpong(5 + i, NFRAMES), :1], K)
            for k, v in c1.stage_ms().items():
                st1[k] = st1.get(k, 0.0) + v / n1
        dt = (time.perf_counter() - t0) / n1
        single = {"ms_per_frame_e2e": 1e3 * dt, "fps": 1.0 / dt, "launches_per_frame": None,
                  "stages_ms": {k: round(v, 4) for k, v in st1.items()},
                  "tracking_only_ms": round(st1.get("lk", 0) + st1.get("ransac_e", 0) + st1.get("pose", 0), 4)}
        l0 = c1.launch_count
        c1.group_step(host_np[0, :1], K)
        single["launches_per_frame"] = c1.launch_count - l0
        c1.close()

    # ---- roofline of the dominant kernel group: the fused per-level ORB kernel ---------------------------
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = json.load(open(peaks_path))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    P = level_pixels(W, H)
    alg_bytes_frame = 5 * sum(P) - P[-1] - P[0]              # pyramid r/w + FAST read + blur r/w
    dense_ms = stages.get("orb_dense", 0.0)
    achieved = (alg_bytes_frame * S) / (dense_ms * 1e-3) / 1e9 if dense_ms > 0 else None
    roofline = {"kernel": "orb_level_kernel (8 launches per step, one per pyramid level)", "bound": "hbm",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": (achieved / peak) if achieved else None,
                "traffic": None, "peak_source": peak_src,
                "algorithmic_bytes_per_frame": alg_bytes_frame, "frames_per_launch_group": S,
                "launch_group_ms": dense_ms}
    prof = os.path.join(ROOT, "profiles", "orb_level_traffic.json")
    if os.path.exists(prof):
        try:
            roofline["traffic"] = json.load(open(prof)).get("dram_bytes_per_launch_group")
        except Exception:
            pass

    # ---- the matching kernel is bound by the integer pipes, not by HBM: measured popc ceiling + live achieved rate.
    # Algorithmically a pair costs 8 xor + 8 popc32; the kernel executes 6 popc per pair (two carry-save adders on the
    # ALU pipe replace two of them).  `frac` is the utilisation of the popc pipe by the population counts actually
    # executed; `algorithmic_over_peak` is the algorithmic popc rate over the same ceiling -- a plain 8-popc kernel
    # cannot exceed 1.0 there.  The ALU pipe is the co-limiter (profiles/README.md). ----
    roofline_knn = None
    if rank == 0 and stages.get("knn", 0) > 0:
        try:
            popc_peak = c1b.measure_popc_peak() if (c1b := Context(64, 64, nfeatures=100, device=local_rank)) else 0.0
            c1b.close()
            nq = float(np.mean(res["n_keypoints"]))
            popc = nq * nq * 8 * S                       # 256-bit descriptors = 8 popc32 per pair
            executed = popc * 6 / 8
            ach = executed / (stages["knn"] * 1e-3)
            roofline_knn = {"kernel": "knn_top2_kernel + knn_finish_kernel",
                            "bound": "integer pipes: popc (quarter rate) and ALU (xor, carry-save adders, top-2)",
                            "achieved": ach / 1e12, "peak": popc_peak / 1e12, "unit": "Tpopc32/s", "frac": ach / popc_peak,
                            "executed_popc_per_pair": 6, "algorithmic_popc_per_pair": 8,
                            "algorithmic_over_peak": (popc / (stages["knn"] * 1e-3)) / popc_peak,
                            "peak_source": "measured live (mvo_measure_popc_peak: 8 independent xor+popc chains per thread)",
                            "algorithmic_popc_per_step": popc, "stage_ms": stages["knn"]}
        except Exception as e:  # pragma: no cover
            roofline_knn = {"error": str(e)}

    # ---- CPU baseline: the reference's OpenCV path on this box's host cores (rank 0, bounded sample) ----
    cpu = None
    if rank == 0 and not args.no_cpu:
        try:
            fps, done, last, threads, ver = cpu_reference_run([s[0] for s in seqs[:2]], K, 10 ** 9, args.cpu_seconds)
            cpu = {"value": fps, "unit": "frames/s", "cores": threads, "kind": "reference",
                   "sample": f"{done} front-end frames of streams 0-1 in {args.cpu_seconds:.0f} s, cv2 {ver} "
                             f"(the OpenCV functions at the reference's call sites), cv2.setNumThreads({threads})"}
        except ImportError:
            cpu = {"value": None, "unit": "frames/s", "cores": 0, "kind": "port", "sample": "cv2 not importable"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8/i32 (ORB, kNN), f32 (LK, H scoring), f64 (F/E solvers, pose)",
            "data": "synthetic",
            "api": "mvo_group_submit / mvo_group_collect (frames resident in HBM, two steps in flight)",
            "unpipelined_value": S * world * min(args.steps, 50) / (ms_dev_sync * 1e-3),
            "config": {"workload": "configs[1] frames (1241x376 KITTI-shaped synthetic sequences, 2000 ORB, full "
                                   "front-end frame: ORB+kNN/ratio+LK 21x21x4+H/F/E RANSAC+recoverPose+triangulate), "
                                   f"{S} independent streams per GPU in lock step (configs[4] sharding)",
                       "streams_per_gpu": S, "frames_per_step": S * world, "width": W, "height": H, "nfeatures": NFEAT,
                       "cache": "inputs larger than L2: each step reads a different frame set, "
                                f"{S} x 10 MB working set > 126 MB L2", "parallelism": f"replicas x{world}, no collective"},
            "e2e": {"value": e2e, "unit": "frames/s", "h2d_bytes_per_step": S * H * W,
                    "d2h_bytes_per_step": S * 128 + S * 4, "ms_per_step": ms_host / args.steps,
                    "api": "mvo_group_submit / mvo_group_collect (pinned host frames, two steps in flight)",
                    "unpipelined_value": frames / (ms_host_sync * 1e-3),
                    "unpipelined_api": "mvo_group_step (synchronous call per step)"},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": roofline,
            "roofline_matching": roofline_knn,
            "cpu_baseline": cpu,
            "stages_ms_per_step": {k: round(v, 4) for k, v in stages.items()},
            "orb_match_ms_per_frame": (stages.get("orb", 0) + stages.get("knn", 0)) / S,
            # configs[1] as worded (LK tracking + essential matrix + recoverPose only): sum of those stage times
            "tracking_only_ms_per_frame": (stages.get("lk", 0) + stages.get("ransac_e", 0) + stages.get("pose", 0)) / S,
            "single_stream": single,
            "last_result_stream0": {k: int(res[0][k]) for k in ("n_keypoints", "n_matches", "n_tracked", "score_h",
                                                                 "score_f", "n_inliers_e", "n_pose_good",
                                                                 "n_triangulated")},
        }
        print(json.dumps(line))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

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


def _cpu_worker(stream_id, nframes, seconds, barrier, q):
    """One single-threaded cv2 worker of the throughput-mode CPU leg: its own camera stream, front-end frames until
    `nframes` are done or `seconds` have passed.  Reports (frames, elapsed seconds)."""
    import cv2
    from oracle import synth
    cv2.setNumThreads(1)
    frames, K = synth.synth_sequence(H, W, stream_id, NFRAMES)
    orb = cv2.ORB_create(NFEAT)
    bf = cv2.BFMatcher(cv2.NORM_HAMMING)
    _, prev = cv2_front_end_frame(cv2, orb, bf, None, frames[0], K)
    _, prev = cv2_front_end_frame(cv2, orb, bf, prev, frames[1], K)      # warm-up frame with every stage
    barrier.wait()
    t0 = time.perf_counter()
    done, t = 0, 2
    while done < nframes and time.perf_counter() - t0 < seconds:
        _, prev = cv2_front_end_frame(cv2, orb, bf, prev, frames[pingpong(t, NFRAMES)], K)
        done += 1
        t += 1
    q.put((done, time.perf_counter() - t0))


def cpu_throughput_run(nframes_per_worker, seconds, workers=None):
    """Throughput mode of the CPU reference: one single-threaded cv2 process per host core, each on its own camera
    stream (ORB, the model searches and recoverPose are single-threaded inside OpenCV, so this -- not one process with
    cv2.setNumThreads(cores) -- is how the host's cores are all kept busy).  Returns (frames/s, frames, workers)."""
    import multiprocessing as mp
    workers = workers or (os.cpu_count() or 1)
    ctx = mp.get_context("spawn")
    barrier = ctx.Barrier(workers)
    q = ctx.Queue()
    procs = [ctx.Process(target=_cpu_worker, args=(s, nframes_per_worker, seconds, barrier, q)) for s in range(workers)]
    for p in procs:
        p.start()
    res = []
    deadline = time.perf_counter() + seconds + 120
    try:
        while len(res) < len(procs):
            try:
                res.append(q.get(timeout=1.0))
            except Exception:
                if any(p.exitcode not in (None, 0) for p in procs) or time.perf_counter() > deadline:
                    raise RuntimeError("a CPU baseline worker died or timed out")
    finally:
        for p in procs:
            p.join(timeout=5)
            if p.is_alive():
                p.kill()          # exact child processes started above
    frames = sum(r[0] for r in res)
    return frames / max(r[1] for r in res), frames, workers


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
    """--impl reference: the reference's own CPU implementation of the path (OpenCV through cv2) on the host, with
    all the host threads it can use: one single-threaded worker per core, each on its own camera stream (throughput
    mode, like the stream groups of our arm).  A step = one front-end frame per worker.  The single-process
    cv2.setNumThreads(cores) figure (latency mode) is reported beside it."""
    if rank != 0:
        return
    try:
        import cv2
    except Exception as e:  # pragma: no cover
        print(json.dumps({"impl": "reference", "unavailable": f"cv2 not importable on this host: {e}"}))
        return
    from oracle import synth
    cores = os.cpu_count() or 1
    t0 = time.perf_counter()
    fps, done, workers = cpu_throughput_run(max(args.steps, 1), 240.0, cores)
    dt_thr = done / fps
    # latency mode: one process, OpenCV's own thread pool
    seqs = [synth.synth_sequence(H, W, s, NFRAMES) for s in range(2)]
    K = seqs[0][1]
    frames = [s[0] for s in seqs]
    cpu_reference_run(frames, K, max(args.warmup, 1) * 2, 30.0)
    lat_fps, lat_done, last, threads, ver = cpu_reference_run(frames, K, 40, 20.0)
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt_thr / max(args.steps, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/i32/f32/f64", "data": "synthetic",
        "config": {"workload": "configs[1] frames: 1241x376 synthetic, 2000 ORB, full front-end frame "
                               "(ORB+kNN+LK+H/F/E RANSAC+recoverPose+triangulate); one camera stream per host core, "
                               "a step = one front-end frame per worker",
                   "frames_per_step": workers, "frames_timed": done, "streams": workers},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": workers, "kind": "reference",
                         "sample": f"{done} front-end frames: {workers} single-threaded cv2 {ver} processes (one per host "
                                   f"core), each on its own stream (the OpenCV calls of the reference's call sites)",
                         "latency_mode": {"value": lat_fps, "unit": "frames/s", "threads": threads,
                                          "sample": f"{lat_done} frames, one process, cv2.setNumThreads({threads})"}},
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
    ap.add_argument("--streams", type=int, default=32, help="camera streams per GPU (stream group size; weak scaling)")
    ap.add_argument("--total-streams", type=int, default=0,
                    help="strong scaling (configs[4] as worded): this many streams in total, sharded round-robin "
                         "over the ranks (256 -> 256 / N per GPU)")
    ap.add_argument("--min-seconds", type=float, default=1.0,
                    help="the timed region is repeated (whole multiples of --steps) until it lasts at least this long")
    ap.add_argument("--cpu-seconds", type=float, default=16.0)
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
    from ros2_mono_vo_b200 import Context, _lib, sharding

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libmonovo_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    args.warmup = max(args.warmup, 3)

    # ---- stream ownership: stream s lives on rank s mod world for its whole life ----------
    if args.total_streams > 0:
        mine = sharding.streams_for_rank(args.total_streams, rank, world)
        scaling, total_streams = "strong", args.total_streams
    else:
        mine = sharding.weak_scaling_streams(args.streams, rank, world)
        scaling, total_streams = "weak", args.streams * world
    S = len(mine)
    if S == 0:
        raise SystemExit("no streams for this rank")

    # ---- synthetic sequences of the streams this rank owns ---------------------------------------------
    seqs = [synth.synth_sequence(H, W, sid, NFRAMES) for sid in mine]
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

    def timed(t_start, steps, submit_fn=None, step_fn=None, consume=None):
        """Times `steps` steps with CUDA events on the context's stream.  submit_fn: the pipelined public API
        (submit(t + 1) before collect(t): the upload / launch work of the next step overlaps the kernels of the current
        one; every step's frames and outputs still cross PCIe inside the timed region).  step_fn: one synchronous
        mvo_group_step per step (also collects the per-stage event timings).  consume(res) is called on every
        collected step (the client reading the outputs)."""
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
                    if consume:
                        consume(res)
                res = ctx.group_collect()
                if consume:
                    consume(res)
            else:
                for i in range(steps):
                    res = step_fn(t_start + i)
                    for k, v in ctx.stage_ms().items():
                        stage_acc[k] = stage_acc.get(k, 0.0) + v
            e1.record(stream)
        barrier()
        clocks = sampler.stop()
        ms = sharding.max_over_ranks(e0.elapsed_time(e1), device="cuda")
        return ms, {k: v / steps for k, v in stage_acc.items()}, ctx.launch_count - l0, clocks, res

    def timed_floor(t_start, **kw):
        """K = --steps steps; if that was shorter than --min-seconds, the measurement is redone over the whole
        multiple of K that reaches it (decided from the max-over-ranks time, so every rank repeats alike)."""
        out = timed(t_start, args.steps, **kw)
        n = args.steps
        if out[0] < args.min_seconds * 1e3:
            reps = int(np.ceil(args.min_seconds * 1e3 / max(out[0], 1e-3) * 1.05))
            n = args.steps * reps
            out = timed(t_start + args.steps, n, **kw)
        return out, n

    # ---- value: frames resident in HBM, result records only ----------------------------------------------
    ctx.group_reset()
    t = 0
    for i in range(args.warmup):
        step_dev(t + i)
    t += args.warmup
    (ms_dev, _, launches, clocks, res), n_dev = timed_floor(t, submit_fn=submit_dev)
    t += args.steps + n_dev
    # per-stage device times (CUDA events inside the library) from synchronous steps; not part of `value`
    n_sync = min(args.steps, 50)
    ms_dev_sync, stages, _, _, _ = timed(t, n_sync, step_fn=step_dev)
    t += n_sync
    lk_track_ms = ctx.debug_time("lk_track", 20)
    lk_pyr_ms = ctx.debug_time("lk_pyramid", 20)
    knn_ms = ctx.debug_time("knn", 20)
    # the ORB stage overlaps the tracker inside a step (own streams), so its in-step event span is not its run time:
    # the level launches are re-run alone on the same frames (this forgets the previous frame; the next loop re-warms)
    orb_dense_ms = ctx.debug_time("orb_levels", 20)
    # ---- e2e, result records only (round-1 definition, kept for comparison) ------------------------------
    def warm_host(t0_):
        # both pipeline slots allocate their staging / pinned output blocks on first use: warm up through the same API
        step_host(t0_)
        timed(t0_ + 1, 4, submit_fn=submit_host)
        return t0_ + 5

    t = warm_host(t)
    (ms_host_rec, _, _, _, _), n_host_rec = timed_floor(t, submit_fn=submit_host)
    t += args.steps + n_host_rec
    # ---- e2e: pinned host frames in, the FULL per-stream outputs back (keypoints, descriptors, matches, LK tracks,
    # H / F / E + masks, triangulated points: everything the reference's data flow consumes) ---------------
    ctx.group_configure(channels=1, outputs=_lib.MVO_OUT_ALL)
    d2h_bytes = ctx.group_output_bytes()
    sink = {"n": 0}

    def consume(res_):
        # the client side of a step: views into the pinned output block of every stream (no copies)
        for s_ in range(0, S, max(S // 4, 1)):
            o = ctx.group_outputs(s_, copy=False)
            sink["n"] += int(o["n_keypoints"]) + int(o["n_tracked"])

    t = warm_host(t)
    (ms_host, _, launches_e2e, clocks_e2e, res_full), n_host = timed_floor(t, submit_fn=submit_host, consume=consume)
    t += args.steps + n_host
    ms_host_sync, _, _, _, _ = timed(t, n_sync, step_fn=step_host)
    t += n_sync
    ctx.group_configure(channels=1, outputs=0)

    value = S_total_frames(total_streams, n_dev) / (ms_dev * 1e-3)
    e2e = S_total_frames(total_streams, n_host) / (ms_host * 1e-3)

    # ---- the reference's real tracking frame in batched form (Tracker::update: LK on the tracked observations +
    # solvePnPRansac against their landmarks, mvo_group_track); secondary number, rank 0 only --------------
    tracking = None
    if rank == 0:
        try:
            tracking = tracking_frame_bench(ctx, synth, mine, host_np, K, S)
        except Exception as e:  # pragma: no cover
            tracking = {"error": str(e)}

    # ---- single-stream latency (configs[1] as worded: one stream on one B200) -------------------------
    single = None
    if rank == 0:
        c1 = Context(W, H, nfeatures=NFEAT, batch=1, device=local_rank)

        def run_single(n1):
            t0 = time.perf_counter()
            st = {}
            for i in range(n1):
                c1.group_step(host_np[pingpong(5 + i, NFRAMES), :1], K)
                for k, v in c1.stage_ms().items():
                    st[k] = st.get(k, 0.0) + v / n1
            return (time.perf_counter() - t0) / n1, st

        for i in range(6):
            c1.group_step(host_np[pingpong(i, NFRAMES), :1], K)
        # the synchronous step of a small group replays a captured CUDA graph (one launch instead of ~85)
        dt, _ = run_single(100)
        gstats = c1.graph_stats()
        l0 = c1.launch_count
        c1.group_step(host_np[0, :1], K)
        kernels_per_frame = c1.launch_count - l0
        # the same loop with the graph form off: plain launches, per-stage CUDA-event spans available
        c1.debug_set("graph", 0)
        for i in range(3):
            c1.group_step(host_np[pingpong(i, NFRAMES), :1], K)
        dt_plain, st1 = run_single(40)
        single = {"ms_per_frame_e2e": 1e3 * dt, "fps": 1.0 / dt, "launches_per_frame": kernels_per_frame,
                  "api": "mvo_group_step, batch 1, host frame in, result record out, synchronous; CUDA-graph replay",
                  "graph": gstats, "ms_per_frame_e2e_plain_launches": 1e3 * dt_plain,
                  "stages_ms": {k: round(v, 4) for k, v in st1.items()},
                  "stages_note": "spans from the plain-launch loop; ORB runs beside the tracker and the model searches",
                  "tracking_only_ms": round(st1.get("lk", 0) + st1.get("ransac_e", 0) + st1.get("pose", 0), 4)}
        c1.close()

    sm_clk_hz = 1e6 * float(clocks.get("sm_mhz") or 1965.0)
    issue_slots_per_s = 148 * 4 * sm_clk_hz                 # 148 SMs x 4 schedulers, one warp instruction per cycle each

    # ---- roofline of the HBM-streaming kernel group: the fused per-level ORB kernel ---------------------------
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peaks = json.load(open(peaks_path)) if os.path.exists(peaks_path) else {}
    if "hbm_gbs" in peaks:
        peak, peak_src = peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    P = level_pixels(W, H)
    alg_bytes_frame = 5 * sum(P) - P[-1] - P[0]              # pyramid r/w + FAST read + blur r/w
    dense_ms = orb_dense_ms
    achieved = (alg_bytes_frame * S) / (dense_ms * 1e-3) / 1e9 if dense_ms > 0 else None
    roofline = {"kernel": "orb_level_kernel (8 launches per step, one per pyramid level)", "bound": "hbm",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": (achieved / peak) if achieved else None,
                "traffic": None, "peak_source": peak_src,
                "algorithmic_bytes_per_frame": alg_bytes_frame, "frames_per_launch_group": S,
                "launch_group_ms": dense_ms}
    prof = _load_profile("orb_level_traffic.json")
    if prof and S == prof.get("frames_per_launch_group", 32):
        roofline["traffic"] = prof.get("dram_bytes_per_launch_group")
        wi = prof.get("warp_instructions_per_launch_group")
        if wi and dense_ms > 0:
            # the kernel is issue-bound, not HBM-bound: share of the GPU's warp-instruction issue slots it uses
            roofline["issue_frac"] = wi / (issue_slots_per_s * dense_ms * 1e-3)
            roofline["warp_instructions_per_launch_group"] = wi
            roofline["thread_instructions_per_pyramid_pixel"] = 32.0 * wi / (sum(P) * S)

    # ---- the dominant kernel by time: lk_track2_kernel.  Its data (a 32 x 32 window region per point and level) lives
    # in shared memory; DRAM < 1 %.  The bound is the issue rate of integer instructions: reported as the share of issue
    # slots used (warp instructions from the committed ncu capture of this workload / slots in the live kernel time).
    roofline_lk = {"kernel": "lk_track2_kernel (1 launch per step: all streams, all points, 4 levels)",
                   "bound": "issue slots (integer pipes: IDP / IMAD, ALU, LSU on shared memory)",
                   "launch_ms": lk_track_ms, "pyramid_ms": lk_pyr_ms, "stage_ms": stages.get("lk"),
                   "share_of_step": lk_track_ms / (ms_dev / n_dev) if ms_dev > 0 else None}
    lkp = _load_profile("lk_track_issue.json")
    if lkp and S == lkp.get("streams", 32) and lk_track_ms > 0:
        wi = lkp["warp_instructions_per_launch"]
        roofline_lk.update({
            "warp_instructions_per_launch": wi, "issue_frac": wi / (issue_slots_per_s * lk_track_ms * 1e-3),
            "thread_instructions_per_point_level": 32.0 * wi / (lkp.get("points_per_launch", 64000) * 4),
            "window_samples_per_point_level": 441,
            "source": lkp.get("source")})
        for k in ("mean_iterations_per_point_level", "thread_instructions_per_window_sample_iteration",
                  "pipe_utilisation_pct"):
            if k in lkp:
                roofline_lk[k] = lkp[k]

    # ---- matching: all-pairs Hamming distances as an int8 GEMM on the tensor cores (descriptor bits -> +-8,
    # q.t = 64 (256 - 2 d)); tcgen05.mma kind::i8, accumulators in TMEM, top-2 selection in the TMEM-read epilogue.
    roofline_knn = None
    if stages.get("knn", 0) > 0:
        nq = float(np.mean(res["n_keypoints"]))
        ops = 2.0 * nq * nq * 256 * S                       # 2 x MACs of the S Nq x Nt x 256 products of a step
        i8_peak = 2.0 * peaks.get("bf16_tflops", 2250.0 / 2 * 2)   # kind::i8 issues at twice the bf16 rate
        roofline_knn = {"kernel": "knn_mma_kernel + knn_finish_kernel", "bound": "tensor",
                        "achieved": ops / (knn_ms * 1e-3) / 1e12, "peak": i8_peak, "unit": "TOP/s (int8)",
                        "frac": ops / (knn_ms * 1e-3) / 1e12 / i8_peak,
                        "peak_source": "2 x MEASURED_PEAKS.json bf16_tflops (kind::i8 runs at twice the dense bf16 rate)"
                        if "bf16_tflops" in peaks else "nominal 4.5 POP/s",
                        "algorithmic_ops_per_step": ops, "stage_ms": knn_ms,
                        "note": "the MMAs of a 128 x 128 x 256 tile take ~0.06 us of tensor time; the kernel is bound "
                                "by the in-kernel operand expansion (32 B -> 256 B per row) and the top-2 epilogue "
                                "(tcgen05.ld + packed 16-bit min/max), see profiles/README.md"}

    # ---- CPU baseline: the reference's OpenCV path on this box's host cores (rank 0, bounded sample) ----
    cpu = None
    if rank == 0 and not args.no_cpu:
        try:
            half = args.cpu_seconds / 2
            thr_fps, thr_done, workers = cpu_throughput_run(10 ** 9, half)
            fps, done, last, threads, ver = cpu_reference_run([s[0] for s in seqs[:2]], K, 10 ** 9, half)
            cpu = {"value": thr_fps, "unit": "frames/s", "cores": workers, "kind": "reference",
                   "sample": f"{thr_done} front-end frames in {half:.0f} s: {workers} single-threaded cv2 {ver} processes "
                             f"(one per host core), each on its own stream (the OpenCV functions at the reference's "
                             f"call sites) -- throughput mode, like the stream groups on the GPU",
                   "latency_mode": {"value": fps, "unit": "frames/s", "threads": threads,
                                    "sample": f"{done} frames of streams 0-1 in {half:.0f} s, one process, "
                                              f"cv2.setNumThreads({threads})"}}
        except ImportError:
            cpu = {"value": None, "unit": "frames/s", "cores": 0, "kind": "port", "sample": "cv2 not importable"}
        except Exception as e:  # pragma: no cover
            cpu = {"value": None, "unit": "frames/s", "cores": 0, "kind": "reference", "sample": f"failed: {e}"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "timed_steps": n_dev, "warmup": args.warmup, "ms_per_step": ms_dev / n_dev, "higher_is_better": True,
            "scaling": scaling, "vs_baseline": None,
            "dtype": "u8/i32 (ORB, LK), s8 tensor-core products (kNN), f32 (LK solve, H scoring), f64 (F/E solvers, pose)",
            "data": "synthetic",
            "api": "mvo_group_submit / mvo_group_collect (frames resident in HBM, two steps in flight)",
            "unpipelined_value": S_total_frames(total_streams, n_sync) / (ms_dev_sync * 1e-3),
            "config": {"workload": "configs[1] frames (1241x376 KITTI-shaped synthetic sequences, 2000 ORB, full "
                                   "front-end frame: ORB+kNN/ratio+LK 21x21x4+H/F/E RANSAC+recoverPose+triangulate), "
                                   f"{S} independent streams per GPU in lock step (configs[4] sharding)",
                       "streams_per_gpu": S, "total_streams": total_streams, "frames_per_step": total_streams,
                       "width": W, "height": H, "nfeatures": NFEAT,
                       "stream_ownership": "round-robin, stream s on rank s mod N (ros2_mono_vo_b200/sharding.py)",
                       "cache": "inputs larger than L2: each step reads a different frame set, "
                                f"{S} x 10 MB working set > 126 MB L2", "parallelism": f"replicas x{world}, no collective",
                       "min_seconds": args.min_seconds},
            "e2e": {"value": e2e, "unit": "frames/s", "h2d_bytes_per_step": S * H * W,
                    "d2h_bytes_per_step": int(d2h_bytes), "ms_per_step": ms_host / n_host, "timed_steps": n_host,
                    "outputs": "full per-stream outputs (MVO_OUT_ALL: keypoints, descriptors, matches, LK tracks / status / "
                               "err, H / F / E + inlier masks, recoverPose mask, triangulated points) + result records, "
                               "into pinned host memory every step",
                    "api": "mvo_group_configure(MVO_OUT_ALL) + mvo_group_submit / mvo_group_collect / mvo_group_outputs "
                           "(pinned host frames, two steps in flight)",
                    "unpipelined_value": S_total_frames(total_streams, n_sync) / (ms_host_sync * 1e-3),
                    "unpipelined_api": "mvo_group_step (synchronous call per step)",
                    "records_only": {"value": S_total_frames(total_streams, n_host_rec) / (ms_host_rec * 1e-3),
                                     "d2h_bytes_per_step": S * 128 + S * 4},
                    "gpu_launches": launches_e2e, "clocks": clocks_e2e},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": roofline,
            "roofline_lk": roofline_lk,
            "roofline_matching": roofline_knn,
            "cpu_baseline": cpu,
            "stages_ms_per_step": {k: round(v, 4) for k, v in stages.items()},
            "stages_note": "CUDA-event spans inside synchronous steps; ORB, the tracker, kNN and the three model searches "
                           "run on their own streams and overlap, so the spans do not add up to the step",
            "kernels_alone_ms": {"orb_dense": round(orb_dense_ms, 4), "lk_track": round(lk_track_ms, 4),
                                 "lk_pyramid": round(lk_pyr_ms, 4), "knn": round(knn_ms, 4)},
            "orb_match_ms_per_frame": (stages.get("orb", 0) + stages.get("knn", 0)) / S,   # in-step spans (overlapped)
            # configs[1] as worded (LK tracking + essential matrix + recoverPose only): sum of those stage times
            "tracking_only_ms_per_frame": (stages.get("lk", 0) + stages.get("ransac_e", 0) + stages.get("pose", 0)) / S,
            "tracking_frame": tracking,
            "single_stream": single,
            "last_result_stream0": {k: int(res[0][k]) for k in ("n_keypoints", "n_matches", "n_tracked", "score_h",
                                                                 "score_f", "n_inliers_e", "n_pose_good",
                                                                 "n_triangulated")},
        }
        print(json.dumps(line))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def S_total_frames(total_streams, steps):
    return total_streams * steps


def _load_profile(name):
    try:
        return json.load(open(os.path.join(ROOT, "profiles", name)))
    except Exception:
        return None


def tracking_frame_bench(ctx, synth, mine, host_np, K, S, reps=30):
    """mvo_group_track over the stream group: every stream tracks the ORB keypoints of its first frame (back-projected
    with the scene depth as landmarks) through the sequence: LK + status / err filter + solvePnPRansac per frame."""
    Kinv = np.linalg.inv(K)
    ctx.group_configure(channels=1, outputs=_lib_out_keypoints())
    ctx.group_step(host_np[0], K)
    obs = []
    for s in range(S):
        o = ctx.group_outputs(s)
        xy = np.stack([o["keypoints"]["x"], o["keypoints"]["y"]], 1).astype(np.float32)
        depth = synth.sequence_depth(H, W, mine[s])
        d = depth[np.clip(np.rint(xy[:, 1]).astype(int), 0, H - 1), np.clip(np.rint(xy[:, 0]).astype(int), 0, W - 1)]
        xyz = ((Kinv @ np.column_stack([xy, np.ones(len(xy))]).T).T * d[:, None]).astype(np.float32)
        obs.append((xy, xyz))
    ctx.group_configure(channels=1, outputs=0)
    times = []
    last = None
    for r in range(reps):
        ctx.group_track(host_np[0], K)                       # (re)start: store frame 0
        for s in range(S):
            ctx.group_set_tracks(s, *obs[s])
        t0 = time.perf_counter()
        last = ctx.group_track(host_np[1], K)
        times.append(time.perf_counter() - t0)
    ms = 1e3 * float(np.median(times))
    return {"api": "mvo_group_track (Tracker::update per-frame path: LK on the tracked observations + solvePnPRansac)",
            "ms_per_step": ms, "frames_per_s": S / (ms * 1e-3), "observations_per_stream": int(np.mean([len(o[0]) for o in obs])),
            "n_tracked_stream0": int(last[0]["n_tracked"]), "n_pnp_inliers_stream0": int(last[0]["n_pnp_inliers"]),
            "pnp_ok_streams": int(np.sum(last["pnp_ok"])), "timing": "host wall clock around the synchronous call, "
            "pinned host frames, median of %d" % reps}


def _lib_out_keypoints():
    from ros2_mono_vo_b200 import _lib
    return _lib.MVO_OUT_KEYPOINTS


if __name__ == "__main__":
    main()

"""Small driver for ncu: mvo_group_track over 32 streams (LK on tracked observations + batched solvePnPRansac)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import synth
from ros2_mono_vo_b200 import Context, _lib
S = int(sys.argv[1]) if len(sys.argv) > 1 else 32
H, W, N = 376, 1241, 2000
seqs = [synth.synth_sequence(H, W, s, 3) for s in range(S)]
K = seqs[0][1]
Kinv = np.linalg.inv(K)
ctx = Context(W, H, nfeatures=N, batch=S)
ctx.group_configure(channels=1, outputs=_lib.MVO_OUT_KEYPOINTS)
f = [np.stack([seqs[s][0][t] for s in range(S)]) for t in range(3)]
ctx.group_step(f[0], K)
obs = []
for s in range(S):
    o = ctx.group_outputs(s)
    xy = np.stack([o["keypoints"]["x"], o["keypoints"]["y"]], 1).astype(np.float32)
    depth = synth.sequence_depth(H, W, s)
    d = depth[np.clip(np.rint(xy[:, 1]).astype(int), 0, H - 1), np.clip(np.rint(xy[:, 0]).astype(int), 0, W - 1)]
    obs.append((xy, ((Kinv @ np.column_stack([xy, np.ones(len(xy))]).T).T * d[:, None]).astype(np.float32)))
ctx.group_configure(channels=1, outputs=0)
for rep in range(4):
    ctx.group_track(f[0], K)
    for s in range(S):
        ctx.group_set_tracks(s, *obs[s])
    t0 = time.perf_counter()
    r = ctx.group_track(f[1], K)
    t1 = time.perf_counter()
    r2 = ctx.group_track(f[2], K)
    t2 = time.perf_counter()
    print("track ms", (t1 - t0) * 1e3, (t2 - t1) * 1e3, int(r[0]["n_tracked"]), int(r[0]["n_pnp_inliers"]), int(r2[0]["n_pnp_inliers"]))
ctx.close()

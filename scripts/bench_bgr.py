"""Secondary figure: the same 32-stream front-end step on BGR8 frames (what the node feeds, src/mono_vo.cpp:94): ORB on
the fused BGR->gray conversion, LK on the three colour planes (lk_track2_kernel<true, 3>: teams of three warps) -- and on
BGR8 frames whose three planes are identical (a gray camera behind the BGR8 conversion: the tracker's gray path with
tripled sums).  Pipelined API, frames resident in HBM and from pinned host memory."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import synth
from ros2_mono_vo_b200 import Context
S, H, W, N = 32, 376, 1241, 2000
seqs = [synth.synth_sequence(H, W, s, 6) for s in range(S)]
K = seqs[0][1]
host = torch.empty((6, S, H, W, 3), dtype=torch.uint8).pin_memory()
for s in range(S):
    for f in range(6):
        g = torch.from_numpy(seqs[s][0][f])
        host[f, s, :, :, 0] = g
        host[f, s, :, :, 1] = torch.roll(g, 1, 1) // 2 + g // 2
        host[f, s, :, :, 2] = 255 - g
order = [0, 1, 2, 3, 4, 5, 4, 3, 2, 1]
ctx = Context(W, H, nfeatures=N, batch=S)
ctx.group_configure(channels=3, outputs=0)
out = {}


def run(content, mode, dev, hn):
    def submit(t):
        if mode == "device":
            ctx.group_submit(None, K, device_ptr=dev[order[t % 10]].data_ptr(), shape=(H, W))
        else:
            ctx.group_submit(hn[order[t % 10]], K)
    for t in range(4):
        submit(t); ctx.group_collect()
    torch.cuda.synchronize()
    n = 300
    t0 = time.perf_counter()
    submit(4)
    for t in range(5, 5 + n - 1):
        submit(t); res = ctx.group_collect()
    res = ctx.group_collect()
    torch.cuda.synchronize()
    out[content + "_" + mode + "_frames_per_s"] = S * n / (time.perf_counter() - t0)
    out[content + "_lk_track_ms"] = ctx.debug_time("lk_track", 10)
    out[content + "_last"] = {k: int(res[0][k]) for k in ("n_keypoints", "n_tracked", "n_inliers_e", "n_pose_good")}


for content in ("colour", "gray_replicated"):
    if content == "gray_replicated":   # three identical planes
        host[:, :, :, :, 1] = host[:, :, :, :, 0]
        host[:, :, :, :, 2] = host[:, :, :, :, 0]
    dev = host.cuda()
    hn = host.numpy()
    for mode in ("device", "host"):
        run(content, mode, dev, hn)
print(json.dumps(out))
ctx.close()

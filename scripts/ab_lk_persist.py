"""A/B of lk_track2_kernel: persistent warps (lk_impl 2) vs one point per warp (lk_impl 3) vs first generation (1)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import synth
from ros2_mono_vo_b200 import Context, _lib
S, H, W, N = 32, 376, 1241, 2000
seqs = [synth.synth_sequence(H, W, s, 3) for s in range(S)]
K = seqs[0][1]
dev = torch.from_numpy(np.stack([np.stack([seqs[s][0][f] for s in range(S)]) for f in range(3)])).cuda()
res = {}
for impl in (3, 2, 1):
    ctx = Context(W, H, nfeatures=N, batch=S)
    ctx.group_configure(channels=1, outputs=_lib.MVO_OUT_TRACKS)
    ctx.debug_set("lk_impl", impl)
    for t in range(3):
        ctx.group_step(None, K, device_ptr=dev[t].data_ptr(), shape=(H, W))
    outs = [ctx.group_outputs(s) for s in range(S)]
    res[impl] = outs
    ms = min(ctx.debug_time("lk_track", 20) for _ in range(3))
    print("lk_impl", impl, "lk_track ms", round(ms, 4))
    ctx.close()
for impl in (2, 1):
    same = all(np.array_equal(a[k], b[k]) for a, b in zip(res[3], res[impl]) for k in ("track_xy", "track_status", "track_err"))
    print("impl", impl, "identical to impl 3:", same)

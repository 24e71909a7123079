"""Pipelined 32-stream throughput against the number of resident CTAs per SM of the persistent LK kernel."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import synth
from ros2_mono_vo_b200 import Context
S, H, W, N = 32, 376, 1241, 2000
seqs = [synth.synth_sequence(H, W, s, 6) for s in range(S)]
K = seqs[0][1]
dev = torch.from_numpy(np.stack([np.stack([seqs[s][0][f] for s in range(S)]) for f in range(6)])).cuda()
order = [0, 1, 2, 3, 4, 5, 4, 3, 2, 1]
for per_sm in (5, 4, 3, 5, 4):
    ctx = Context(W, H, nfeatures=N, batch=S)
    ctx.debug_set("lk_ctas_per_sm", per_sm)
    for t in range(4):
        ctx.group_step(None, K, device_ptr=dev[order[t % 10]].data_ptr(), shape=(H, W))
    torch.cuda.synchronize()
    n = 400
    t0 = time.perf_counter()
    ctx.group_submit(None, K, device_ptr=dev[order[4]].data_ptr(), shape=(H, W))
    for t in range(5, 5 + n - 1):
        ctx.group_submit(None, K, device_ptr=dev[order[t % 10]].data_ptr(), shape=(H, W))
        ctx.group_collect()
    ctx.group_collect()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("lk CTAs per SM", per_sm, "frames/s", round(S * n / dt), "lk_track alone ms", round(ctx.debug_time("lk_track", 10), 4))
    ctx.close()

"""Schedule of one pipelined group step under steady state: start / end of every stage relative to the step start."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import synth
from ros2_mono_vo_b200 import Context
S = int(sys.argv[1]) if len(sys.argv) > 1 else 32
H, W, N = 376, 1241, 2000
seqs = [synth.synth_sequence(H, W, s, 6) for s in range(S)]
K = seqs[0][1]
dev = torch.from_numpy(np.stack([np.stack([seqs[s][0][f] for s in range(S)]) for f in range(6)])).cuda()
ctx = Context(W, H, nfeatures=N, batch=S)
order = [0, 1, 2, 3, 4, 5, 4, 3, 2, 1]
for mode in ("sync", "pipelined"):
    ctx.group_reset()
    if mode == "sync":
        for t in range(12):
            ctx.group_step(None, K, device_ptr=dev[order[t % 10]].data_ptr(), shape=(H, W))
    else:
        ctx.group_submit(None, K, device_ptr=dev[0].data_ptr(), shape=(H, W))
        for t in range(1, 40):
            ctx.group_submit(None, K, device_ptr=dev[order[t % 10]].data_ptr(), shape=(H, W))
            ctx.group_collect()
        ctx.group_collect()
    torch.cuda.synchronize()
    sp = ctx.stage_spans_ms()
    print(mode)
    for k, (a, b) in sorted(sp.items(), key=lambda kv: kv[1][0]):
        print(f"  {k:12s} {a:7.3f} -> {b:7.3f}  ({b - a:6.3f} ms)")
ctx.close()

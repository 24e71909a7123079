"""Host-side cost of one mvo_group_submit (enqueue only) against the GPU time of the step."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import synth
from ros2_mono_vo_b200 import Context
S = int(sys.argv[1]) if len(sys.argv) > 1 else 32
H, W, N = 376, 1241, 2000
seqs = [synth.synth_sequence(H, W, s, 4) for s in range(S)]
K = seqs[0][1]
host = [torch.from_numpy(np.stack([seqs[s][0][f] for s in range(S)])).pin_memory() for f in range(4)]
dev = [h.cuda() for h in host]
ctx = Context(W, H, nfeatures=N, batch=S)
for mode in ("resident", "host"):
    ctx.group_reset()
    sub, col = [], []
    def submit(t):
        f = [0, 1, 2, 3, 2, 1][t % 6]
        t0 = time.perf_counter()
        if mode == "resident":
            ctx.group_submit(None, K, device_ptr=dev[f].data_ptr(), shape=(H, W))
        else:
            ctx.group_submit(host[f].numpy(), K)
        sub.append(time.perf_counter() - t0)
    submit(0)
    T0 = time.perf_counter()
    for t in range(1, 201):
        submit(t)
        t0 = time.perf_counter()
        ctx.group_collect()
        col.append(time.perf_counter() - t0)
    ctx.group_collect()
    wall = (time.perf_counter() - T0) / 200
    print(f"{mode}: wall {1e3*wall:.3f} ms/step, submit mean {1e3*np.mean(sub[5:]):.3f} ms (p95 {1e3*np.percentile(sub[5:],95):.3f}), collect wait mean {1e3*np.mean(col[5:]):.3f} ms")
ctx.close()

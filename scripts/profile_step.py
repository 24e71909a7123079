"""Small driver for ncu: a stream group of 32 C2 streams, a few warm-up steps and N profiled steps (device-resident
frames, synchronous mvo_group_step so that launches of consecutive steps do not interleave in the launch list)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from oracle import synth
from ros2_mono_vo_b200 import Context

S = int(sys.argv[1]) if len(sys.argv) > 1 else 32
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
H, W, N = 376, 1241, 2000
seqs = [synth.synth_sequence(H, W, s, 6) for s in range(S)]
K = seqs[0][1]
dev = torch.from_numpy(np.stack([np.stack([seqs[s][0][f] for s in range(S)]) for f in range(6)])).cuda()
ctx = Context(W, H, nfeatures=N, batch=S)
for t in range(3 + steps):
    f = [0, 1, 2, 3, 4, 5, 4, 3, 2, 1][t % 10]
    if t == 3:
        torch.cuda.synchronize()
        print("LAUNCHES_BEFORE", ctx.launch_count)
    res = ctx.group_step(None, K, device_ptr=dev[f].data_ptr(), shape=(H, W))
print("LAUNCHES_AFTER", ctx.launch_count, "per step", None)
print(ctx.stage_ms())
ctx.close()

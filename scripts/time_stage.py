"""Device time of one re-run pipeline stage (mvo_debug_time) for 32 C2 streams, per kernel choice.
    python scripts/time_stage.py knn      -> tensor-core vs popc matching kernel
    python scripts/time_stage.py lk_track -> second- vs first-generation LK kernel"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from oracle import synth
from ros2_mono_vo_b200 import Context

what = sys.argv[1] if len(sys.argv) > 1 else "knn"
S, H, W, N = 32, 376, 1241, 2000
seqs = [synth.synth_sequence(H, W, s, 3) for s in range(S)]
K = seqs[0][1]
dev = torch.from_numpy(np.stack([np.stack([seqs[s][0][f] for s in range(S)]) for f in range(3)])).cuda()
ctx = Context(W, H, nfeatures=N, batch=S)
for t in range(3):
    res = ctx.group_step(None, K, device_ptr=dev[t].data_ptr(), shape=(H, W))
key = {"knn": "knn_impl", "lk_track": "lk_impl"}.get(what)
for impl in ((0, 1) if what == "knn" else (2, 1)) if key else (0,):
    if key:
        ctx.debug_set(key, impl)
    ms = min(ctx.debug_time(what, 20) for _ in range(3))
    print(f"{what} impl {impl}: {ms:.4f} ms per 32-stream step (n_matches stream 0: {res[0]['n_matches']})")
ctx.close()

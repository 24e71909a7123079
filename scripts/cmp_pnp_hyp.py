"""Per-hypothesis agreement of the EPnP kernel (mvo_pnp_get_hypotheses) with the oracle on the same subsets."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import pnp_oracle as po, synth
from ros2_mono_vo_b200 import Context
ctx = Context(1241, 376, nfeatures=2000)
for seed in (1, 3):
    obj, img, K, _, _ = synth.pnp_scene(2000, seed, 0.5, 0.05)
    ctx.solve_pnp_ransac(obj, img, K)
    sub, mdl, cnt = ctx.pnp_hypotheses(100)
    ref = po.sample_subsets(2000, 100)
    print("subsets equal", np.array_equal(sub, ref))
    dRs, dts, dc = [], [], []
    xn_all = po.normalize(img, K).astype(np.float32).astype(np.float64)
    for i in range(100):
        R, t = po.epnp(obj[sub[i]].astype(np.float64), xn_all[sub[i]])
        dRs.append(np.abs(mdl[i, :9].reshape(3, 3) - R).max()); dts.append(np.abs(mdl[i, 9:] - t).max())
    dRs, dts = np.array(dRs), np.array(dts)
    print("seed", seed, "tight(1e-9)", int((dRs < 1e-9).sum()), "<1e-6", int((dRs < 1e-6).sum()), "<5e-3", int((dRs < 5e-3).sum()), "max dR", dRs.max(), "max dt", dts.max())
    print(" worst:", np.argsort(-dRs)[:8], np.sort(dRs)[::-1][:8], "counts", cnt[np.argsort(-dRs)[:8]])
ctx.close()

"""What the tolerant assertions of round 1 actually see (status flips LK vs oracle, recoverPose count, PnP inlier sets)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from conftest import load_golden
from oracle import lk_oracle as lo, ransac_oracle as ro, synth
from ros2_mono_vo_b200 import Context

ctx = Context(1920, 1080, nfeatures=5000, max_points=5200)
for (h, w, seed, n) in [(480, 640, 41, 1000), (1080, 1920, 43, 5000), (100, 47, 44, 50)]:
    f0, f1 = synth.synth_pair(h, w, seed)
    rng = np.random.default_rng(seed)
    pts = np.stack([rng.uniform(-25, w + 25, n), rng.uniform(-25, h + 25, n)], 1).astype(np.float32)
    nxt, st, err = ctx.lk_track(f0, f1, pts)
    onxt, ost, oerr = lo.lk_track(f0, f1, pts)
    d = np.nonzero(st != ost)[0]
    print("LK", (h, w, seed, n), "status flips", d.tolist(), [(int(st[i]), int(ost[i]), pts[i].tolist(), nxt[i].tolist(), onxt[i].tolist()) for i in d])
g = load_golden("ransac.npz")
K = g["K"]
for tag in "abcd":
    n, seed, planar = g[f"{tag}_args"].tolist()
    p1, p2, Rgt, tgt, inl = synth.scene_correspondences(int(n), int(seed), planar=bool(planar), outlier_frac=0.25)
    E, me, ne = ctx.find_essential(p1, p2, K, 0.99, 1.0)
    R, t, mp, good = ctx.recover_pose(E, p1, p2, K, mask=me)
    print("pose golden", tag, "good", good, "cv2", int(g[f"{tag}_good"]), "mask equal", np.array_equal(mp != 0, g[f"{tag}_mask_pose"] != 0))
for seed in range(60, 70):
    for planar in (False, True):
        p1, p2, Rgt, tgt, inl = synth.scene_correspondences(1000, seed, planar=planar, outlier_frac=0.25)
        try:
            E, me, ne = ctx.find_essential(p1, p2, K, 0.99, 1.0)
        except Exception as e:
            print("E failed", seed, planar, e); continue
        R, t, mp, good = ctx.recover_pose(E, p1, p2, K, mask=me)
        Ro, to, mo, goodo = ro.recover_pose(E, p1, p2, K, mask=me)
        if good != goodo or not np.array_equal(mp != 0, np.asarray(mo).ravel() != 0):
            print("pose oracle diff", seed, planar, good, goodo, np.nonzero((mp != 0) != (np.asarray(mo).ravel() != 0))[0])
print("pose oracle sweep done")
gp = load_golden("pnp.npz")
for tag in ["easy", "outl30", "outl50", "noisy", "few", "c3"]:
    n, seed, noise, outl, planar = gp[f"{tag}_args"].tolist()
    obj, img, Kp, _, _ = synth.pnp_scene(int(n), int(seed), noise, outl, bool(planar))
    ok, r, t, inl = ctx.solve_pnp_ransac(obj, img, Kp)
    want = gp[f"{tag}_inliers"]
    print("pnp", tag, "sym diff", sorted(set(inl.tolist()) ^ set(want.tolist())), "dr", np.abs(r - gp[f"{tag}_rvec"]).max(), "dt", np.abs(t - gp[f"{tag}_tvec"]).max())
ctx.close()

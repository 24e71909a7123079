"""Live parity sweep on a GPU box that has cv2: every stage of the path against OpenCV itself on seeded inputs that are
not among the committed goldens (other seeds, sizes, feature counts).  Prints one summary line per stage.
  ORB   keypoint sets / responses / angles / descriptors (cv2.ORB_create(n).detectAndCompute)
  kNN   ratio-test matches (BFMatcher(NORM_HAMMING).knnMatch + 0.7 ratio)
  LK    status, positions, errors (calcOpticalFlowPyrLK defaults), gray and BGR8
  H/F/E RANSAC masks and models on scene correspondences; recoverPose masks"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, cv2
from oracle import synth
from ros2_mono_vo_b200 import Context
cv2.setNumThreads(1)
ctx = Context(1920, 1080, nfeatures=5000, max_points=5000)

# ---- ORB + kNN ----
tot = bad_set = bad_desc = 0; n_kp = 0; knn_bad = 0; knn_tot = 0
for (h, w, n) in ((376, 1241, 2000), (480, 640, 1000), (240, 320, 500), (1080, 1920, 5000), (300, 401, 700)):
    octx = Context(w, h, nfeatures=n)
    for seed in (201, 202, 203):
        f0, f1 = synth.synth_pair(h, w, seed)
        orb = cv2.ORB_create(n)
        descs = []
        for f in (f0, f1):
            kc, dc = orb.detectAndCompute(f, None)
            kg, dg = octx.orb_detect_and_compute(f)
            tot += 1
            key_c = {(k.octave, k.pt[0], k.pt[1]): (k.response, k.angle, i) for i, k in enumerate(kc)}
            key_g = {(int(k["octave"]), float(k["x"]), float(k["y"])): (float(k["response"]), float(k["angle"]), i) for i, k in enumerate(kg)}
            n_kp += len(kc)
            if set(key_c) != set(key_g) or any(key_c[k][:2] != key_g[k][:2] for k in key_c):
                bad_set += 1
                continue
            bad_desc += sum(1 for k in key_c if not np.array_equal(dc[key_c[k][2]], dg[key_g[k][2]]))
            descs.append((dc, dg, kc, kg))
        if len(descs) == 2:
            (d0c, d0g, k0c, k0g), (d1c, d1g, k1c, k1g) = descs
            mc = {(k0c[m[0].queryIdx].pt, k1c[m[0].trainIdx].pt) for m in cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(d0c, d1c, 2) if len(m) == 2 and m[0].distance < 0.7 * m[1].distance}
            mg = octx.knn_ratio(d0g, d1g, 0.7)
            mg = {((float(k0g["x"][a]), float(k0g["y"][a])), (float(k1g["x"][b]), float(k1g["y"][b]))) for a, b in zip(mg["query_idx"], mg["train_idx"])}
            knn_tot += len(mc); knn_bad += len(mc ^ mg)
    octx.close()
print(f"ORB: {tot} frames, {n_kp} keypoints: frames with a different keypoint set / response / angle: {bad_set}; descriptors differing: {bad_desc}")
print(f"kNN: {knn_tot} ratio-test matches, symmetric difference to cv2: {knn_bad}  (ties between equal distances may pick another of two equal keypoints)")

# ---- LK ----
for bgr in (False, True):
    n_pts = st_bad = 0; dmax = emax = 0.0; over = [0, 0]
    for (h, w, n) in ((376, 1241, 2000), (480, 640, 1000), (121, 203, 300), (1080, 1920, 3000)):
        for seed in (301, 302, 303):
            f0, f1 = (synth.synth_pair_bgr if bgr else synth.synth_pair)(h, w, seed)
            rng = np.random.default_rng(seed)
            pts = rng.uniform([-2, -2], [w + 1, h + 1], (n, 2)).astype(np.float32)
            nc, sc, ec = cv2.calcOpticalFlowPyrLK(f0, f1, pts, None)
            ng, sg, eg = ctx.lk_track(f0, f1, pts)
            sc = sc.ravel(); ec = ec.ravel()
            n_pts += n; st_bad += int((sc != sg).sum())
            m = (sc == 1) & (sg == 1)
            d = np.abs(nc[m] - ng[m]).max(1)
            over[0] += int((d > 1e-3).sum()); over[1] += int((d > 1e-2).sum())
            dmax = max(dmax, float(d.max())); emax = max(emax, float(np.abs(ec[m] - eg[m]).max()))
    print(f"LK {'BGR8' if bgr else 'gray'}: {n_pts} points (incl. border and outside points): status differs on {st_bad}; max |dpos| {dmax:.2e} px "
          f"({over[0]} points above 1e-3, {over[1]} above 1e-2), max |derr| {emax:.2e}")

# ---- H / F / E / recoverPose ----
K = synth.KITTI_K
res = {k: [0, 0] for k in ("H", "F", "E", "pose")}
worst = {k: 0.0 for k in ("H", "F", "E")}
for n in (200, 1000, 2000):
    for seed in range(400, 410):
        for outl, planar in ((0.1, False), (0.3, False), (0.5, False), (0.3, True)):
            p1, p2, R, t, inl = synth.scene_correspondences(n, seed, outl, 0.3, planar)
            cv2.setRNGSeed(0); Hc, mh = cv2.findHomography(p1, p2, cv2.RANSAC, 1.0)
            cv2.setRNGSeed(0); Fc, mf = cv2.findFundamentalMat(p1, p2, cv2.FM_RANSAC, 1.0, 0.99)
            cv2.setRNGSeed(0); Ec, me = cv2.findEssentialMat(p1, p2, K, cv2.RANSAC, 0.99, 1.0)
            Hg, mhg, _ = ctx.find_homography(p1, p2, 1.0)
            Fg, mfg, _ = ctx.find_fundamental(p1, p2, 1.0, 0.99)
            Eg, meg, _ = ctx.find_essential(p1, p2, K, 0.99, 1.0)
            for name, mc, mg, Mc, Mg in (("H", mh, mhg, Hc, Hg), ("F", mf, mfg, Fc, Fg), ("E", me, meg, Ec, Eg)):
                res[name][0] += 1
                same = mc is not None and np.array_equal(mc.ravel(), mg)
                res[name][1] += int(same)
                if same and Mc is not None and Mc.shape == (3, 3):
                    a = Mc / np.linalg.norm(Mc); b = Mg / np.linalg.norm(Mg)
                    worst[name] = max(worst[name], float(min(np.abs(a - b).max(), np.abs(a + b).max())))
            if Ec is not None and Ec.shape == (3, 3) and np.array_equal(me.ravel(), meg):
                gc, Rc, tc, mp = cv2.recoverPose(Ec, p1, p2, K, distanceThresh=50.0, mask=me.copy())[:4]
                Rg, tg, mpg, gg = ctx.recover_pose(Eg, p1, p2, K, meg)
                res["pose"][0] += 1
                res["pose"][1] += int(np.array_equal(mp.ravel() != 0, mpg != 0))
for k in ("H", "F", "E"):
    print(f"{k}: mask identical to cv2 on {res[k][1]} / {res[k][0]} scenes; model difference on those <= {worst[k]:.1e} (unit norm)")
print(f"recoverPose (50-baseline threshold): mask identical on {res['pose'][1]} / {res['pose'][0]} scenes")
ctx.close()

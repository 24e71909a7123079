"""Generate golden vectors from cv2 (the OpenCV the reference links, via its Python binding).

Run in the build container:  python tests/golden/gen_golden.py
Every fixture records cv2.__version__.  Inputs come from oracle/synth.py seeds (numpy-only, so they can
be regenerated anywhere); each fixture stores the sha256 of its input so a drifting generator is caught.
The calls mirror the reference's call sites:
  ORB   /root/reference/src/feature_processor.cpp:6,22   cv::ORB::create(n)->detectAndCompute
  kNN   /root/reference/src/feature_processor.cpp:7,29   BFMatcher(NORM_HAMMING).knnMatch(q, t, 2)
"""
import hashlib
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import synth  # noqa: E402


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def kp_arrays(kps):
    return dict(x=np.array([k.pt[0] for k in kps], np.float32), y=np.array([k.pt[1] for k in kps], np.float32),
                size=np.array([k.size for k in kps], np.float32), angle=np.array([k.angle for k in kps], np.float32),
                response=np.array([k.response for k in kps], np.float32),
                octave=np.array([k.octave for k in kps], np.int32))


def gen_orb(name, h, w, seed, n, store_image):
    img = synth.synth_frame(h, w, seed)
    orb = cv2.ORB_create(n)
    kps, desc = orb.detectAndCompute(img, None)
    d = kp_arrays(kps)
    d.update(desc=desc, cv2_version=cv2.__version__, h=h, w=w, seed=seed, nfeatures=n, image_sha=sha(img))
    # pyramid + per-level FAST from cv2 primitives (sub-stage oracles)
    scales = [np.float32(float(np.float32(1.2)) ** l) for l in range(8)]
    lvl = img
    fast = cv2.FastFeatureDetector_create(20, True)
    k32 = cv2.getGaussianKernel(7, 2, cv2.CV_32F)
    for l in range(8):
        if l > 0:
            inv = np.float32(1.0) / scales[l]
            sz = (int(np.rint(np.float32(w) * inv)), int(np.rint(np.float32(h) * inv)))
            lvl = cv2.resize(lvl, sz, interpolation=cv2.INTER_LINEAR_EXACT)
        d[f"level{l}_sha"] = sha(lvl)
        d[f"blur{l}_sha"] = sha(cv2.sepFilter2D(lvl, -1, k32, k32, borderType=cv2.BORDER_REFLECT_101))
        fk = fast.detect(lvl)
        d[f"fast{l}"] = np.array([(int(k.pt[0]), int(k.pt[1]), int(k.response)) for k in fk], np.int32).reshape(-1, 3)
    if store_image:
        d["image"] = img
    np.savez_compressed(os.path.join(HERE, name), **d)
    print(name, len(kps))


def gen_knn():
    rng = np.random.default_rng(5)
    # low-entropy descriptors -> many first/second ties (tie rule: lowest train index)
    q = rng.integers(0, 4, (300, 32)).astype(np.uint8)
    t = rng.integers(0, 4, (400, 32)).astype(np.uint8)
    m = cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(q, t, 2)
    idx = np.array([[a.trainIdx, b.trainIdx] for a, b in m], np.int32)
    dist = np.array([[a.distance, b.distance] for a, b in m], np.float32)
    # realistic descriptors: ORB of a synthetic pair
    f0, f1 = synth.synth_pair(240, 320, 11)
    orb = cv2.ORB_create(300)
    _, d0 = orb.detectAndCompute(f0, None)
    _, d1 = orb.detectAndCompute(f1, None)
    m2 = cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(d0, d1, 2)
    idx2 = np.array([[a.trainIdx, b.trainIdx] for a, b in m2], np.int32)
    dist2 = np.array([[a.distance, b.distance] for a, b in m2], np.float32)
    good = np.array([(a.queryIdx, a.trainIdx, a.distance) for a, b in m2 if a.distance < 0.7 * b.distance], np.float32)
    np.savez_compressed(os.path.join(HERE, "knn.npz"), q=q, t=t, idx=idx, dist=dist, d0=d0, d1=d1, idx2=idx2,
                        dist2=dist2, good=good, cv2_version=cv2.__version__)
    print("knn", len(m), len(good))


def lk_points(h, w, f0, n, seed):
    """ORB keypoints of f0 + border / outside / flat-region points (edge cases of src/tracker.cpp:61-69)."""
    orb = cv2.ORB_create(n)
    kps = orb.detect(f0, None)
    pts = np.array([k.pt for k in kps], np.float32)
    rng = np.random.default_rng(seed)
    extra = np.array([[0, 0], [w - 1, h - 1], [2.5, 3.5], [w - 3.2, 10.1], [-5, -5], [w + 30, h + 30],
                      [w / 2, h - 1.5], [-40, 20], [w + 25.5, h / 2]], np.float32)
    rnd = np.stack([rng.uniform(-15, w + 15, 64), rng.uniform(-15, h + 15, 64)], 1).astype(np.float32)
    return np.concatenate([pts, extra, rnd])


def gen_lk():
    out = {}
    for tag, h, w, seed, n in (("small", 240, 320, 11, 300), ("c2", 376, 1241, 2, 2000)):
        f0, f1 = synth.synth_pair(h, w, seed)
        pts = lk_points(h, w, f0, n, seed)
        nxt, st, err = cv2.calcOpticalFlowPyrLK(f0, f1, pts, None)
        out.update({f"{tag}_pts": pts, f"{tag}_next": nxt, f"{tag}_status": st.ravel(), f"{tag}_err": err.ravel(),
                    f"{tag}_sha0": sha(f0), f"{tag}_sha1": sha(f1), f"{tag}_hw_seed": np.array([h, w, seed])})
        print("lk", tag, len(pts), int(st.sum()))
    # a flat next image: every point must fail the min-eigenvalue / go nowhere consistently
    f0 = synth.synth_frame(120, 160, 4)
    flat = np.full_like(f0, 100)
    pts = lk_points(120, 160, f0, 100, 1)
    nxt, st, err = cv2.calcOpticalFlowPyrLK(flat, f0, pts, None)
    out.update(flat_pts=pts, flat_next=nxt, flat_status=st.ravel(), flat_err=err.ravel())
    out["cv2_version"] = cv2.__version__
    np.savez_compressed(os.path.join(HERE, "lk.npz"), **out)


def gen_lk_bgr():
    """cn = 3 windows: the node tracks on BGR8 images (src/mono_vo.cpp:94 -> src/tracker.cpp:68)."""
    out = {}
    for tag, h, w, seed, n in (("small", 240, 320, 12, 300), ("c2", 376, 1241, 3, 2000)):
        f0, f1 = synth.synth_pair_bgr(h, w, seed)
        pts = lk_points(h, w, cv2.cvtColor(f0, cv2.COLOR_BGR2GRAY), n, seed)
        nxt, st, err = cv2.calcOpticalFlowPyrLK(f0, f1, pts, None)
        out.update({f"{tag}_pts": pts, f"{tag}_next": nxt, f"{tag}_status": st.ravel(), f"{tag}_err": err.ravel(),
                    f"{tag}_sha0": sha(f0), f"{tag}_sha1": sha(f1), f"{tag}_hw_seed": np.array([h, w, seed])})
        print("lk_bgr", tag, len(pts), int(st.sum()))
    out["cv2_version"] = cv2.__version__
    np.savez_compressed(os.path.join(HERE, "lk_bgr.npz"), **out)


PNP_CASES = (("easy", 2000, 1, 0.5, 0.05, False), ("outl30", 2000, 2, 1.0, 0.30, False), ("outl50", 500, 3, 0.3, 0.50, False),
             ("noisy", 100, 4, 2.0, 0.20, False), ("few", 30, 5, 0.5, 0.10, False), ("c3", 5000, 6, 0.7, 0.15, False))


def gen_pnp():
    """cv::solvePnPRansac with the reference's arguments (src/tracker.cpp:309): 100 iterations, 8 px, 0.99."""
    out = {}
    for tag, n, seed, noise, outl, planar in PNP_CASES:
        obj, img, K, _, _ = synth.pnp_scene(n, seed, noise, outl, planar)
        ok, r, t, inl = cv2.solvePnPRansac(obj, img, K, None, None, None, False, 100, 8.0, 0.99)
        out.update({f"{tag}_args": np.array([n, seed, noise, outl, int(planar)], np.float64), f"{tag}_sha": sha(np.concatenate([obj.ravel(), img.ravel()])),
                    f"{tag}_ok": ok, f"{tag}_rvec": r.ravel(), f"{tag}_tvec": t.ravel(), f"{tag}_inliers": inl.ravel().astype(np.int32)})
        print("pnp", tag, ok, len(inl))
    # minimal solver + refinement vectors
    obj, img, K, _, _ = synth.pnp_scene(400, 9, 0.5, 0.0)
    rng = np.random.default_rng(0)
    i6 = np.array([rng.choice(400, 6, replace=False) for _ in range(8)])
    out["hyp_i6"] = i6
    out["hyp_epnp"] = np.array([np.concatenate([x.ravel() for x in cv2.solvePnP(obj[i], img[i], K, None, flags=cv2.SOLVEPNP_EPNP)[1:]])
                                for i in i6])
    ok, r, t = cv2.solvePnP(obj, img, K, None, flags=cv2.SOLVEPNP_ITERATIVE)
    out["iter_rt"] = np.concatenate([r.ravel(), t.ravel()])
    out["cv2_version"] = cv2.__version__
    np.savez_compressed(os.path.join(HERE, "pnp.npz"), **out)


from gen_golden_common import distort_pixels  # noqa: E402


PNP_DIST = np.array([-0.28, 0.07, 0.0008, -0.0005, -0.01])


def gen_pnp_dist():
    """cv::solvePnPRansac with CameraInfo-style distortion coefficients (the reference passes d_, src/tracker.cpp:309)."""
    out = {"dist": PNP_DIST, "cv2_version": cv2.__version__}
    for tag, n, seed, noise, outl in (("d1", 1500, 11, 0.5, 0.2), ("d2", 300, 12, 0.3, 0.1)):
        obj, img, K, rv, tv = synth.pnp_scene(n, seed, noise, outl)
        imgd = distort_pixels(img, K, PNP_DIST)
        ok, r, t, inl = cv2.solvePnPRansac(obj, imgd, K, PNP_DIST, None, None, False, 100, 8.0, 0.99)
        out.update({f"{tag}_args": np.array([n, seed, noise, outl]), f"{tag}_ok": ok, f"{tag}_rvec": r.ravel(), f"{tag}_tvec": t.ravel(),
                    f"{tag}_inliers": inl.ravel().astype(np.int32), f"{tag}_rv_gt": rv, f"{tag}_tv_gt": tv})
        print("pnp_dist", tag, ok, len(inl), np.abs(r.ravel() - rv).max(), np.abs(t.ravel() - tv).max())
    np.savez_compressed(os.path.join(HERE, "pnp_dist.npz"), **out)


def gen_ransac():
    """cv2 outputs at the reference's RANSAC call sites (src/initializer.cpp:82,87,228,236,125)."""
    K = synth.KITTI_K
    out = {"K": K, "cv2_version": cv2.__version__}
    for tag, seed, planar, n in (("a", 5, True, 2000), ("b", 6, False, 2000), ("c", 7, True, 500), ("d", 8, False, 800)):
        p1, p2, R, t, inl = synth.scene_correspondences(n, seed, planar=planar, outlier_frac=0.25)
        H, mh = cv2.findHomography(p1, p2, cv2.RANSAC, 1.0)
        F, mf = cv2.findFundamentalMat(p1, p2, cv2.FM_RANSAC, 1.0, 0.99)
        E, me = cv2.findEssentialMat(p1, p2, K, cv2.RANSAC, 0.99, 1.0)
        mp = me.copy()
        good, Rr, tr, mp = cv2.recoverPose(E, p1, p2, K, mask=mp)
        P0 = K @ np.eye(3, 4)
        P1 = K @ np.column_stack([Rr, tr])
        X = cv2.triangulatePoints(P0, P1, p1.T.copy(), p2.T.copy())
        out.update({f"{tag}_args": np.array([n, seed, int(planar)]), f"{tag}_sha": sha(np.concatenate([p1, p2])),
                    f"{tag}_H": H, f"{tag}_mask_h": mh.ravel(), f"{tag}_F": F, f"{tag}_mask_f": mf.ravel(),
                    f"{tag}_E": E, f"{tag}_mask_e": me.ravel(), f"{tag}_R": Rr, f"{tag}_t": tr.ravel(),
                    f"{tag}_mask_pose": mp.ravel(), f"{tag}_good": good, f"{tag}_X": X, f"{tag}_Rgt": R, f"{tag}_tgt": t})
        print("ransac", tag, int(mh.sum()), int(mf.sum()), int(me.sum()), good)
    # per-hypothesis solver vectors
    p1, p2, R, t, inl = synth.scene_correspondences(400, 9, outlier_frac=0.0)
    rng = np.random.default_rng(0)
    i4 = np.array([rng.choice(400, 4, replace=False) for _ in range(8)])
    i7 = np.array([rng.choice(400, 7, replace=False) for _ in range(8)])
    i5 = np.array([rng.choice(400, 5, replace=False) for _ in range(8)])
    out["hyp_p1"], out["hyp_p2"], out["hyp_i4"], out["hyp_i7"], out["hyp_i5"] = p1, p2, i4, i7, i5
    out["hyp_H"] = np.array([cv2.findHomography(p1[i], p2[i], 0)[0] for i in i4])
    q1 = np.stack([(p1[:, 0] - K[0, 2]) / K[0, 0], (p1[:, 1] - K[1, 2]) / K[1, 1]], 1).astype(np.float64)
    q2 = np.stack([(p2[:, 0] - K[0, 2]) / K[0, 0], (p2[:, 1] - K[1, 2]) / K[1, 1]], 1).astype(np.float64)
    for j, i in enumerate(i7):
        out[f"hyp_F{j}"] = cv2.findFundamentalMat(p1[i], p2[i], cv2.FM_7POINT)[0].reshape(-1, 3, 3)
    for j, i in enumerate(i5):
        out[f"hyp_E{j}"] = cv2.findEssentialMat(q1[i], q2[i], np.eye(3))[0].reshape(-1, 3, 3)
    np.savez_compressed(os.path.join(HERE, "ransac.npz"), **out)


def gen_f_small():
    """cv::findFundamentalMat(FM_RANSAC) below 15 correspondences (the Tracker's has_parallax can get there with
    min_tracked_points = 10, src/tracker.cpp:239-248): N == 7 direct, 8 <= N < 15 LMedS."""
    out = {"cv2_version": cv2.__version__}
    cases = [(n, 100 + s) for s in range(6) for n in (7, 8, 10, 13, 14)]
    out["cases"] = np.array(cases, np.int32)
    for n, seed in cases:
        p1, p2, R, t, inl = synth.scene_correspondences(n, seed, outlier_frac=0.2, noise_px=0.3)
        F, mask = cv2.findFundamentalMat(p1, p2, cv2.FM_RANSAC, 1.0, 0.99)
        out[f"n{n}_s{seed}_sha"] = sha(np.concatenate([p1, p2]))
        out[f"n{n}_s{seed}_F"] = np.zeros((0, 3)) if F is None else F
        out[f"n{n}_s{seed}_mask"] = mask.ravel()
        print("f_small", n, seed, None if F is None else F.shape, int(mask.sum()))
    np.savez_compressed(os.path.join(HERE, "f_small.npz"), **out)


def gen_c3():
    """BASELINE configs[2] size against cv2 itself: 1920x1080, 5000 ORB features over 8 levels, LK on the keypoints,
    keyframe matching 5000 x 5000 with Lowe ratio 0.7 (round-1 parity at this size was against the oracle only)."""
    gen_orb("orb_c3.npz", 1080, 1920, 3, 5000, False)
    f0, f1 = synth.synth_pair(1080, 1920, 3)
    pts = lk_points(1080, 1920, f0, 5000, 3)
    nxt, st, err = cv2.calcOpticalFlowPyrLK(f0, f1, pts, None)
    np.savez_compressed(os.path.join(HERE, "lk_c3.npz"), c3_pts=pts, c3_next=nxt, c3_status=st.ravel(), c3_err=err.ravel(),
                        c3_sha0=sha(f0), c3_sha1=sha(f1), c3_hw_seed=np.array([1080, 1920, 3]), cv2_version=cv2.__version__)
    print("lk c3", len(pts), int(st.sum()))
    orb = cv2.ORB_create(5000)
    _, d0 = orb.detectAndCompute(f0, None)
    _, d1 = orb.detectAndCompute(f1, None)
    m = cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(d0, d1, 2)
    idx = np.array([[a.trainIdx, b.trainIdx] for a, b in m], np.int32)
    dist = np.array([[a.distance, b.distance] for a, b in m], np.float32)
    good = np.array([(a.queryIdx, a.trainIdx, a.distance) for a, b in m if a.distance < 0.7 * b.distance], np.float32)
    np.savez_compressed(os.path.join(HERE, "knn_c3.npz"), d0=d0, d1=d1, idx=idx, dist=dist, good=good, cv2_version=cv2.__version__)
    print("knn c3", len(m), len(good))


if __name__ == "__main__":
    only = sys.argv[1:]
    if only:            # e.g. `python gen_golden.py gen_lk_bgr`: regenerate single fixtures
        for name in only:
            globals()[name]()
        sys.exit(0)
    gen_ransac()
    gen_f_small()
    gen_lk()
    gen_lk_bgr()
    gen_pnp()
    gen_orb("orb_small.npz", 240, 320, 3, 300, True)
    gen_orb("orb_c1.npz", 480, 640, 1, 1000, False)
    gen_orb("orb_c2.npz", 376, 1241, 2, 2000, False)
    gen_knn()
    gen_c3()
    gen_pnp_dist()

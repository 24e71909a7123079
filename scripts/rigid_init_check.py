"""The reference's initialisation chain (ORB -> kNN ratio 0.7 -> findEssentialMat(0.99, 1.0) -> recoverPose with the 50-baseline
threshold: /root/reference/src/initializer.cpp:150-280) on frames (0, j) of the rigid synthetic sequence: the library's single
calls beside cv2's, both against the ground truth."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, cv2
from oracle import synth
from ros2_mono_vo_b200 import Context
cv2.setRNGSeed(0)
fr, K, R_wc, C = synth.synth_rigid_sequence(376, 1241, int(sys.argv[1]) if len(sys.argv) > 1 else 0, 4)
ctx = Context(1241, 376, nfeatures=1000)
def err(Rr, t, j):
    Rgt = R_wc[j].T; tgt = -R_wc[j].T @ C[j]; tgt /= np.linalg.norm(tgt)
    return (np.degrees(np.arccos(np.clip((np.trace(Rr @ Rgt.T) - 1) / 2, -1, 1))), np.degrees(np.arccos(np.clip(np.ravel(t) @ tgt / np.linalg.norm(t), -1, 1))))
orb = cv2.ORB_create(1000)
k0c, d0c = orb.detectAndCompute(fr[0], None)
k0, d0 = ctx.orb_detect_and_compute(fr[0])
xy = np.array([k.pt for k in k0c], np.float32)
print("keypoints equal", len(k0c) == len(k0) and np.array_equal(xy[:, 0], k0["x"]) and np.array_equal(xy[:, 1], k0["y"]), "descriptor rows differing", int((d0c != d0).any(1).sum()), "of", len(d0))
for j in (1, 2, 3):
    k1c, d1c = orb.detectAndCompute(fr[j], None)
    good = [m[0] for m in cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(d0c, d1c, 2) if len(m) == 2 and m[0].distance < 0.7 * m[1].distance]
    p0 = np.array([k0c[m.queryIdx].pt for m in good], np.float32); p1 = np.array([k1c[m.trainIdx].pt for m in good], np.float32)
    E, mask = cv2.findEssentialMat(p0, p1, K, cv2.RANSAC, 0.99, 1.0)
    n, Rr, t, mp, X = cv2.recoverPose(E, p0, p1, K, distanceThresh=50.0, mask=mask.copy())
    print(j, "cv2 ", len(good), int(mask.sum()), n, "rot / t err deg %.3f %.3f" % err(Rr, t, j))
    k1, d1 = ctx.orb_detect_and_compute(fr[j])
    mm = ctx.knn_ratio(d0, d1, 0.7)
    q0 = np.stack([k0["x"][mm["query_idx"]], k0["y"][mm["query_idx"]]], 1).astype(np.float32)
    q1 = np.stack([k1["x"][mm["train_idx"]], k1["y"][mm["train_idx"]]], 1).astype(np.float32)
    print("   same matches", np.array_equal(q0, p0) and np.array_equal(q1, p1))
    Eg, mg, ng = ctx.find_essential(q0, q1, K, 0.99, 1.0)
    Rg, tg, mg2, ngood = ctx.recover_pose(Eg, q0, q1, K, mg)
    print(j, "b200", len(mm), ng, ngood, "rot / t err deg %.3f %.3f" % err(Rg, tg, j), "E equal", np.abs(Eg / np.linalg.norm(Eg) - E / np.linalg.norm(E)).max())
ctx.close()

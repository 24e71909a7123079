"""The reference's Initializer / Tracker (oracle/_ref/mono_vo_host on the GPU library) on an exactly rigid sequence:
estimated camera poses against the ground truth, per frame."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from oracle import synth
import test_ref_host as trh
import pathlib, tempfile
n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
stream = int(sys.argv[2]) if len(sys.argv) > 2 else 0
frames, K, R_wc, C = synth.synth_rigid_sequence(376, 1241, stream, n)
d = pathlib.Path(tempfile.mkdtemp())
trh.write_sequence(d / "seq.bin", frames, K)
params, recs = trh._run(d / "seq.bin")
init = [r["frame"] for r in recs if r["init_event"]]
print("init events", init)
k0 = init[0]
pose = {r["frame"]: np.array(r["pose_wc"]).reshape(4, 4) for r in recs[k0:] if r["pose_wc"] is not None}
scale = None
for f in sorted(pose):
    P = pose[f]
    Rerr = np.degrees(np.arccos(np.clip((np.trace(P[:3, :3] @ R_wc[f].T) - 1) / 2, -1, 1)))
    if scale is None:
        scale = np.linalg.norm(C[f]) / max(np.linalg.norm(P[:3, 3]), 1e-9)
    rv = lambda R: np.degrees(np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]]) / 2)
    print(f, "est", np.round(P[:3, 3] * scale, 3), "gt", np.round(C[f], 3), "rot err %.3f deg" % Rerr, "rvec est", np.round(rv(P[:3, :3]), 3),
          "gt", np.round(rv(R_wc[f]), 3), "landmarks", recs[f]["landmarks"])

"""GPU: the batched product surface.

* mvo_group_outputs: every per-stream output of a group step (keypoints, descriptors, matches, LK tracks, H / F / E and
  their masks, the recoverPose mask, triangulated points) equals the chain of single-call ABI functions **bit for bit**, at
  the benchmarked batch (32 streams) and at configs[4]'s 256 streams, in the synchronous and the pipelined form.
* BGR8 frames in the group step == the BGR8 single calls (ORB on the fused gray conversion, 3-channel LK).
* mvo_group_track (Tracker::update's per-frame path: LK on the tracked observations + solvePnPRansac against their
  landmarks, /root/reference/src/tracker.cpp:274-316) == mvo_lk_track + the status / err filter + mvo_solve_pnp_ransac.
The single calls are each parity-tested against cv2 goldens / the oracle elsewhere.
"""
import numpy as np
import pytest

from oracle import synth

pytestmark = pytest.mark.gpu


def _streams(h, w, batch, nframes, nbase=8, bgr=False):
    """`batch` distinct frame sequences: `nbase` rendered rigid-scene sequences, the others are cyclic shifts of them
    (a shifted sequence is still a rigid scene seen by the same camera motion; it keeps the generator cost bounded)."""
    base = [synth.synth_sequence(h, w, s, nframes) for s in range(min(nbase, batch))]
    K = base[0][1]
    out = []
    for b in range(batch):
        fr = base[b % len(base)][0]
        k = b // len(base)
        fr = [np.roll(f, (7 * k, 13 * k), (0, 1)) for f in fr] if k else fr
        if bgr:
            fr = [np.stack([f, np.roll(f, 1, 1) // 2 + f // 2, 255 - f], 2) for f in fr]
        out.append(fr)
    return out, K


def _chain(ctx1, prev, img, K):
    """One front-end frame through the single-call ABI; returns everything the group step outputs."""
    kps, desc = ctx1.orb_detect_and_compute(img)
    o = {"keypoints": kps, "descriptors": desc}
    if prev is not None:
        pimg, pkps, pdesc = prev
        o["matches"] = ctx1.knn_ratio(pdesc, desc, 0.7)
        pts = np.stack([pkps["x"], pkps["y"]], 1)
        nxt, st, err = ctx1.lk_track(pimg, img, pts)
        o.update(track_xy=nxt, track_status=st, track_err=err)
        ok = (st == 1) & (err < 30.0)
        p1, p2 = pts[ok], nxt[ok]
        o["H"], o["mask_h"], _ = ctx1.find_homography(p1, p2, 1.0)
        o["F"], o["mask_f"], _ = ctx1.find_fundamental(p1, p2, 1.0, 0.99)
        o["E"], o["mask_e"], _ = ctx1.find_essential(p1, p2, K, 0.99, 1.0)
        R, t, o["mask_pose"], _ = ctx1.recover_pose(o["E"], p1, p2, K, mask=o["mask_e"])
        o["X4"] = ctx1.triangulate(K @ np.eye(3, 4), K @ np.column_stack([R, t]), p1, p2)
        o["R"], o["t"] = R, t
    return o, (img, kps, desc)


def _assert_outputs_equal(got, exp, res, where):
    for key in ("keypoints", "descriptors", "matches", "track_xy", "track_status", "track_err"):
        if key in exp:
            a, b = got[key], exp[key]
            assert a is not None and a.shape == b.shape and a.tobytes() == b.tobytes(), (where, key)
    if "H" in exp:
        for key in ("mask_h", "mask_f", "mask_e", "mask_pose"):
            assert np.array_equal(got[key] != 0, np.asarray(exp[key]).ravel() != 0), (where, key)
        for key in ("H", "F", "E"):
            assert np.array_equal(got[key], np.asarray(exp[key]).reshape(3, 3)), (where, key)
        assert got["X4"].tobytes() == np.ascontiguousarray(exp["X4"], np.float32).tobytes(), (where, "X4")
        assert np.array_equal(res["R"].reshape(3, 3), exp["R"]) and np.array_equal(res["t"], np.ravel(exp["t"])), (where, "pose")
        assert got["n_tracked"] == len(exp["mask_h"]) and got["n_prev"] == len(exp["track_status"]), where


@pytest.mark.parametrize("h,w,n,batch,pipelined", [(376, 1241, 2000, 32, False), (376, 1241, 2000, 256, True),
                                                   (240, 320, 300, 3, True)])
def test_group_outputs_equal_single_calls(h, w, n, batch, pipelined):
    import torch
    from ros2_mono_vo_b200 import Context, _lib
    nframes = 3
    seqs, K = _streams(h, w, batch, nframes)
    frames = torch.empty((nframes, batch, h, w), dtype=torch.uint8).pin_memory()
    fn = frames.numpy()
    for t in range(nframes):
        for s in range(batch):
            fn[t, s] = seqs[s][t]
    grp = Context(w, h, nfeatures=n, batch=batch)
    grp.group_configure(channels=1, outputs=_lib.MVO_OUT_ALL)
    assert grp.group_output_bytes() > batch * n * 100
    single = Context(w, h, nfeatures=n, batch=1, max_points=n + n // 4 + 64)
    prev = [None] * batch

    def check(t, res):
        for s in range(batch):
            exp, prev[s] = _chain(single, prev[s], fn[t, s], K)
            _assert_outputs_equal(grp.group_outputs(s), exp, res[s], (t, s))

    if pipelined:
        grp.group_submit(fn[0], K)
        for t in range(1, nframes):
            grp.group_submit(fn[t], K)
            check(t - 1, grp.group_collect())
        check(nframes - 1, grp.group_collect())
    else:
        for t in range(nframes):
            check(t, grp.group_step(fn[t], K))
    grp.close()
    single.close()


def test_group_step_bgr_equals_single_calls():
    from ros2_mono_vo_b200 import Context, _lib
    h, w, n, batch, nframes = 240, 320, 300, 3, 3
    seqs, K = _streams(h, w, batch, nframes, bgr=True)
    grp = Context(w, h, nfeatures=n, batch=batch)
    grp.group_configure(channels=3, outputs=_lib.MVO_OUT_ALL)
    single = Context(w, h, nfeatures=n, batch=1, max_points=n + n // 4 + 64)
    prev = [None] * batch
    for t in range(nframes):
        imgs = np.stack([seqs[s][t] for s in range(batch)])
        res = grp.group_step(imgs, K)
        for s in range(batch):
            exp, prev[s] = _chain(single, prev[s], imgs[s], K)
            _assert_outputs_equal(grp.group_outputs(s), exp, res[s], (t, s))
        assert (res["n_keypoints"] > 100).all()
        if t:
            assert (res["n_tracked"] > 100).all()
    # a gray group context on the gray conversion of the same frames sees the same keypoints (ORB runs on gray)
    grp.close()
    single.close()


@pytest.mark.parametrize("h,w,n,batch,bgr", [(376, 1241, 2000, 32, False), (240, 320, 300, 256, False),
                                             (240, 320, 300, 4, True)])
def test_group_track_equals_single_calls(h, w, n, batch, bgr):
    """Tracker::update's per-frame path over four frames; the observation list stays on the device between frames."""
    from ros2_mono_vo_b200 import Context
    nframes = 4
    seqs, K = _streams(h, w, batch, nframes, bgr=bgr)
    grp = Context(w, h, nfeatures=n, batch=batch)
    grp.group_configure(channels=3 if bgr else 1, outputs=0)
    single = Context(w, h, nfeatures=n, batch=1, max_points=n + n // 4 + 64)
    Kinv = np.linalg.inv(K)
    state = []
    res0 = grp.group_track(np.stack([seqs[s][0] for s in range(batch)]), K)
    assert (res0["n_tracked"] == 0).all() and (res0["pnp_ok"] == 0).all()
    for s in range(batch):
        # observations with landmarks: ORB keypoints of frame 0 back-projected with the scene depth
        kps, _ = single.orb_detect_and_compute(seqs[s][0])
        m = n // 2 + 17 * (s % 5)
        xy = np.stack([kps["x"], kps["y"]], 1)[:m].astype(np.float32)
        depth = synth.sequence_depth(h, w, s % 8)
        d = depth[np.clip(np.rint(xy[:, 1]).astype(int), 0, h - 1), np.clip(np.rint(xy[:, 0]).astype(int), 0, w - 1)]
        xyz = ((Kinv @ np.column_stack([xy, np.ones(len(xy))]).T).T * d[:, None]).astype(np.float32)
        if s == 1:
            xy, xyz = xy[:4], xyz[:4]          # too few points for PnP: pnp_ok 0, tracks still carried
        grp.group_set_tracks(s, xy, xyz)
        state.append((xy, xyz))
    for t in range(1, nframes):
        imgs = np.stack([seqs[s][t] for s in range(batch)])
        res = grp.group_track(imgs, K)
        for s in range(batch):
            xy, xyz = state[s]
            nxt, st, err = single.lk_track(seqs[s][t - 1], seqs[s][t], xy)
            ok = (st == 1) & (err < 30.0)
            kept = np.nonzero(ok)[0].astype(np.int32)
            gxy, gsrc, ginl = grp.group_get_tracks(s)
            assert int(res[s]["n_prev"]) == len(xy) and int(res[s]["n_tracked"]) == len(kept), (t, s)
            assert np.array_equal(gsrc, kept) and gxy.tobytes() == nxt[ok].tobytes(), (t, s)
            state[s] = (nxt[ok], xyz[ok])
            if len(kept) >= 6:
                okp, rvec, tvec, inl = single.solve_pnp_ransac(xyz[ok], nxt[ok], K)
                assert bool(res[s]["pnp_ok"]) == bool(okp), (t, s)
                if okp:
                    assert np.array_equal(ginl, inl), (t, s)
                    assert np.array_equal(res[s]["rvec"], np.ravel(rvec)) and np.array_equal(res[s]["tvec"], np.ravel(tvec)), (t, s)
            else:
                assert int(res[s]["pnp_ok"]) == 0 and len(ginl) == 0, (t, s)
        if not bgr:
            big = res["n_tracked"] >= 50
            assert big.sum() >= batch - 1 and (res["pnp_ok"][big] == 1).all()
    grp.close()
    single.close()


def test_lk_single_call_between_group_steps_resets_previous_frame():
    """A synchronous mvo_lk_track rebuilds both LK pyramids of a batch-1 context: the next group step must not track
    against them (it restarts from feature extraction, like after mvo_group_reset)."""
    from ros2_mono_vo_b200 import Context
    h, w, n = 240, 320, 300
    (frames, K) = synth.synth_sequence(h, w, 3, 3)
    c = Context(w, h, nfeatures=n, batch=1)
    c.group_step(frames[0][None], K)
    r1 = c.group_step(frames[1][None], K)
    assert r1["n_tracked"][0] > 50
    c.lk_track(frames[2], frames[0], np.array([[50.0, 50.0]], np.float32))
    r2 = c.group_step(frames[2][None], K)
    assert r2["n_tracked"][0] == 0 and r2["n_matches"][0] == 0 and r2["n_keypoints"][0] > 100
    r3 = c.group_step(frames[1][None], K)
    assert r3["n_tracked"][0] > 50
    c.close()

"""The C++17 drop-in (mono_vo::FeatureProcessor + mono_vo::gpu::*) builds warning-free against the cv shim
and, on a GPU box, returns the same results as the Python mirror of the same ABI."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT
from oracle import synth

CPP = os.path.join(ROOT, "ros2_mono_vo_b200", "cpp")
EXE = os.path.join(ROOT, "tests", "cpp", "test_adapter")


def _build():
    from ros2_mono_vo_b200 import build
    build.build()
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Wpedantic", "-Werror", "-DMVO_FORCE_CV_SHIM",
           f"-I{CPP}/include", f"-I{ROOT}/include", f"{CPP}/src/feature_processor.cpp",
           f"{ROOT}/tests/cpp/test_adapter.cpp", f"-L{ROOT}/ros2_mono_vo_b200", "-lmonovo_b200",
           f"-Wl,-rpath,{ROOT}/ros2_mono_vo_b200", "-o", EXE]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return EXE


def test_adapter_compiles_with_reference_signatures():
    _build()
    hdr = open(os.path.join(CPP, "include", "mono_vo", "feature_processor.hpp")).read()
    # the four public signatures of the reference class (include/mono_vo/feature_processor.hpp:14-31)
    for sig in ("int num_features = 1000, rclcpp::Logger logger = rclcpp::get_logger(\"FeatureProcessor\")",
                "std::vector<cv::KeyPoint> detect(const cv::Mat & image) const;",
                "const cv::Mat & image, std::vector<cv::KeyPoint> & keypoints, cv::Mat & descriptors) const;",
                "const cv::Mat & descriptors1, const cv::Mat & descriptors2, double lowes_distance_ratio) const;"):
        assert sig in hdr


@pytest.mark.gpu
def test_adapter_matches_python_mirror(tmp_path):
    exe = _build()
    h, w, n = 240, 320, 300
    (frames, K) = synth.synth_sequence(h, w, 5, 2)
    f0, f1 = frames
    p0, p1 = tmp_path / "f0.raw", tmp_path / "f1.raw"
    f0.tofile(p0)
    f1.tofile(p1)
    obj, img, Kp, _, _ = synth.pnp_scene(500, 8, 0.5, 0.2)
    po_, pi_ = tmp_path / "obj.raw", tmp_path / "img.raw"
    obj.tofile(po_)
    img.tofile(pi_)
    r = subprocess.run([exe, str(p0), str(p1), str(w), str(h), str(n), str(po_), str(pi_), str(len(obj))],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    tok = r.stdout.split()
    assert tok[0] == "pnp"
    pnp_inl, pnp_rt = int(tok[1]), np.array(tok[2:8], np.float64)
    tok = tok[8:]
    got = {tok[i]: tok[i + 1] for i in range(0, len(tok) - 1) if tok[i] in ("detect", "matches", "tracked", "score_h",
                                                                             "score_f", "good")}
    from ros2_mono_vo_b200 import Context
    ctx = Context(w, h, nfeatures=n, max_points=n + n // 4 + 64)
    k0, d0 = ctx.orb_detect_and_compute(f0)
    k1, d1 = ctx.orb_detect_and_compute(f1)
    m = ctx.knn_ratio(d0, d1, 0.7)
    pts = np.stack([k0["x"], k0["y"]], 1)
    nxt, st, err = ctx.lk_track(f0, f1, pts)
    ok = (st == 1) & (err < 30.0)
    H, mh, nh = ctx.find_homography(pts[ok], nxt[ok], 1.0)
    F, mf, nf = ctx.find_fundamental(pts[ok], nxt[ok], 1.0, 0.99)
    E, me, ne = ctx.find_essential(pts[ok], nxt[ok], K, 0.99, 1.0)
    R, t, mp_, good = ctx.recover_pose(E, pts[ok], nxt[ok], K, mask=me)
    assert int(tok[1]) == len(k0) and int(tok[2]) == len(k1)
    assert int(got["detect"]) == len(k0) and int(got["matches"]) == len(m) and int(got["tracked"]) == int(ok.sum())
    assert int(got["score_h"]) == nh and int(got["score_f"]) == nf and int(got["good"]) == good
    okp, rv, tv, inl = ctx.solve_pnp_ransac(obj, img, Kp)
    assert okp and pnp_inl == len(inl) and np.abs(pnp_rt - np.concatenate([rv, tv])).max() < 1e-9
    ctx.close()

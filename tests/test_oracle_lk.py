"""Pin the LK oracle (oracle/lk_oracle.py) against cv2.calcOpticalFlowPyrLK golden vectors."""
import numpy as np
import pytest

from conftest import load_golden, sha
from oracle import lk_oracle as lo
from oracle import synth


def _pair(g, tag):
    h, w, seed = g[f"{tag}_hw_seed"].tolist()
    f0, f1 = synth.synth_pair(h, w, seed)
    assert sha(f0) == str(g[f"{tag}_sha0"]) and sha(f1) == str(g[f"{tag}_sha1"]), "synthetic generator drifted"
    return f0, f1


@pytest.mark.parametrize("tag", ["small", "c2", "c3"])
def test_lk_oracle_vs_cv2(tag):
    g = load_golden("lk_c3.npz" if tag == "c3" else "lk.npz")
    f0, f1 = _pair(g, tag)
    nxt, st, err = lo.lk_track(f0, f1, g[f"{tag}_pts"])
    assert np.array_equal(st, g[f"{tag}_status"])
    m = st == 1
    assert m.sum() > 0.9 * len(st) - 80
    assert np.abs(nxt[m] - g[f"{tag}_next"][m]).max() < 1e-3          # px
    assert np.abs(err[m] - g[f"{tag}_err"][m]).max() < 5e-3
    # lost points keep the last position estimate
    assert np.abs(nxt[~m] - g[f"{tag}_next"][~m]).max() < 1e-3


def test_lk_oracle_flat_image():
    g = load_golden("lk.npz")
    f0 = synth.synth_frame(120, 160, 4)
    nxt, st, err = lo.lk_track(np.full_like(f0, 100), f0, g["flat_pts"])
    assert np.array_equal(st, g["flat_status"])
    assert st.sum() == 0                                              # min-eigenvalue rejects every point
    assert np.allclose(nxt, g["flat_next"], atol=1e-4)


def test_pyrdown_and_levels():
    img = synth.synth_frame(100, 47, 1)
    assert len(lo.build_pyramid(img)) == 2          # 47 -> 24 (> 21) -> 12 (stop)
    assert len(lo.build_pyramid(synth.synth_frame(376, 1241, 1))) == 4
    p = lo.pyr_down(np.full((9, 11), 77, np.uint8))
    assert p.shape == (5, 6) and (p == 77).all()


@pytest.mark.parametrize("tag", ["small", "c2"])
def test_lk_oracle_bgr_vs_cv2(tag):
    """cn = 3 windows (the node tracks BGR8 images): every sum spans the three channels."""
    g = load_golden("lk_bgr.npz")
    h, w, seed = g[f"{tag}_hw_seed"].tolist()
    f0, f1 = synth.synth_pair_bgr(h, w, seed)
    assert sha(f0) == str(g[f"{tag}_sha0"]) and sha(f1) == str(g[f"{tag}_sha1"]), "synthetic generator drifted"
    nxt, st, err = lo.lk_track(f0, f1, g[f"{tag}_pts"])
    assert np.array_equal(st, g[f"{tag}_status"])
    m = st == 1
    assert np.abs(nxt[m] - g[f"{tag}_next"][m]).max() < 1e-3
    assert np.abs(err[m] - g[f"{tag}_err"][m]).max() < 5e-3
    assert np.abs(nxt[~m] - g[f"{tag}_next"][~m]).max() < 1e-3

import hashlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


def sha(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def golden_image(g):
    """Regenerate (or load) the input image of an ORB fixture and check it against the recorded sha."""
    from oracle import synth
    if "image" in g.files:
        img = g["image"]
    else:
        img = synth.synth_frame(int(g["h"]), int(g["w"]), int(g["seed"]))
    assert sha(img) == str(g["image_sha"]), "synthetic generator drifted: regenerate tests/golden"
    return img


def kp_key_set(octave, x, y):
    return set(zip(np.asarray(octave).tolist(), np.round(np.asarray(x, np.float64), 3).tolist(),
                   np.round(np.asarray(y, np.float64), 3).tolist()))


def kp_dict(octave, x, y, *vals):
    keys = zip(np.asarray(octave).tolist(), np.round(np.asarray(x, np.float64), 3).tolist(),
               np.round(np.asarray(y, np.float64), 3).tolist())
    return {k: i for i, k in enumerate(keys)}


@pytest.fixture(scope="session")
def lib():
    from ros2_mono_vo_b200 import _lib
    return _lib.load()

"""ros2_mono_vo_b200 -- B200-native (sm_100a) front-end hot path of Tatsuya-2/ros2_mono_vo.

The product is libmonovo_b200.so (hand-written CUDA behind the C ABI of include/monovo_b200.h) plus
host-side mirrors of the reference's FeatureProcessor interface.  Python is used for tests and the
benchmark harness only; the reference-facing host code is C++ (ros2_mono_vo_b200/cpp).
"""
from .api import Context, FeatureProcessor, MvoError  # noqa: F401

__all__ = ["Context", "FeatureProcessor", "MvoError"]

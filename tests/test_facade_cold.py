"""Host-only helpers of the OpenCV-API facade (cv::Mat / cv::Affine3d algebra, convertPointsFromHomogeneous, ...): built
with the reference's warning level and run on the CPU (no GPU call is made)."""
import os
import subprocess

from conftest import ROOT

FACADE = os.path.join(ROOT, "ros2_mono_vo_b200", "cpp", "facade")
EXE = os.path.join(ROOT, "tests", "cpp", "test_facade_cold")


def test_facade_cold_helpers():
    from ros2_mono_vo_b200 import build
    build.build()
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Wpedantic", "-Werror", f"-I{FACADE}/include", f"-I{ROOT}/include",
           f"{FACADE}/src/opencv_b200.cpp", f"{ROOT}/tests/cpp/test_facade_cold.cpp", f"-L{ROOT}/ros2_mono_vo_b200",
           "-lmonovo_b200", f"-Wl,-rpath,{ROOT}/ros2_mono_vo_b200", "-o", EXE]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([EXE], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stdout + r.stderr

"""Build recipe: the reference's OWN host sources, unchanged, on top of libmonovo_b200.so.

    python -m oracle.build_ref_host            ->  oracle/_ref/mono_vo_host (+ params.txt)

Compiles /root/reference/src/{feature_processor,frame,keyframe,landmark,map,match_data,initializer,tracker}.cpp
where they lie (nothing is copied into the repo) with the reference's own include directory, against the
OpenCV-API facade in ros2_mono_vo_b200/cpp/facade (opencv2/*, rclcpp/* stand-ins whose hot cv:: functions call
the C ABI of include/monovo_b200.h), plus the small driver facade/src/vo_host_main.cpp that replays
MonoVO::image_callback (/root/reference/src/mono_vo.cpp:83-153).  The node itself (mono_vo.cpp, utils.cpp) needs
ROS 2 message packages and is not built.  The parameter YAML (/root/reference/config/params.yaml) is flattened to
oracle/_ref/params.txt so that the run on the GPU box (where /root/reference does not exist) uses the reference's
own parameter values.

Outputs go to oracle/_ref/ only (git-ignored, travels with gpurun).  This is TEST infrastructure: it demonstrates the
drop-in boundary (the reference's Initializer / Tracker state machines running on the CUDA library) and is used by
tests/test_ref_host.py only.
"""
from __future__ import annotations

import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("MVO_REFERENCE", "/root/reference")
OUT = os.path.join(ROOT, "oracle", "_ref")
FACADE = os.path.join(ROOT, "ros2_mono_vo_b200", "cpp", "facade")
LIBDIR = os.path.join(ROOT, "ros2_mono_vo_b200")
REF_SOURCES = ["feature_processor", "frame", "keyframe", "landmark", "map", "match_data", "initializer", "tracker"]
BIN = os.path.join(OUT, "mono_vo_host")
PARAMS = os.path.join(OUT, "params.txt")


def available() -> bool:
    return os.path.isdir(os.path.join(REF, "src"))


def flatten_params() -> None:
    """config/params.yaml (two-level 'group: / name: value' under ros__parameters) -> 'group.name value' lines."""
    import yaml
    with open(os.path.join(REF, "config", "params.yaml")) as f:
        doc = yaml.safe_load(f)
    params = doc["mono_vo"]["ros__parameters"]
    with open(PARAMS, "w") as f:
        for group, entries in params.items():
            if isinstance(entries, dict):
                for name, value in entries.items():
                    f.write(f"{group}.{name} {float(value)!r}\n")


def build(force: bool = False) -> str:
    if not available():
        raise RuntimeError(f"{REF} is not present: the reference host build only runs where the reference is mounted")
    os.makedirs(OUT, exist_ok=True)
    flatten_params()
    srcs = [os.path.join(REF, "src", s + ".cpp") for s in REF_SOURCES]
    mine = [os.path.join(FACADE, "src", "opencv_b200.cpp"), os.path.join(FACADE, "src", "vo_host_main.cpp")]
    deps = srcs + mine + [os.path.join(LIBDIR, "libmonovo_b200.so")]
    for d, _, files in os.walk(os.path.join(FACADE, "include")):
        deps += [os.path.join(d, f) for f in files]
    if not force and os.path.exists(BIN) and all(os.path.getmtime(BIN) >= os.path.getmtime(p) for p in deps):
        return BIN
    # the reference's own warning level (CMakeLists.txt:4-6) minus -Werror: its sources trip -Wparentheses /
    # -Wunused-variable under gcc 13 and they are compiled as they are
    cmd = ["g++", "-std=c++17", "-O2", "-Wall", "-Wextra", "-Wno-parentheses", "-Wno-unused-variable", "-Wno-unused-parameter",
           "-I", os.path.join(FACADE, "include"), "-I", os.path.join(REF, "include"), "-I", os.path.join(ROOT, "include"),
           *srcs, *mine, "-L", LIBDIR, "-lmonovo_b200", "-Wl,-rpath,$ORIGIN/../../ros2_mono_vo_b200", "-o", BIN]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        raise RuntimeError("reference host build failed")
    if r.stderr.strip():
        sys.stderr.write(r.stderr)
    return BIN


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))

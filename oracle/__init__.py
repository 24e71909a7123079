"""CPU oracles for the ros2_mono_vo front-end hot path.

TEST INFRASTRUCTURE.  Nothing in ros2_mono_vo_b200/ imports this package; only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do.
"""

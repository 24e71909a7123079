"""A/B timing of kernel build variants in one GPU call.

    python scripts/lk_variants.py build     (here, no GPU)  -> ros2_mono_vo_b200/variants/libmonovo_<name>.so
    python scripts/lk_variants.py run       (GPU box)       -> one line per variant: LK track ms for 32 x C2 streams

A variant = extra nvcc -D flags for one source file; everything else is linked from the regular build's objects.  `run`
loads every variant in its own process (MVO_B200_LIB), runs three group steps of 32 C2 streams, times the re-run of the
LK stage on that state (mvo_debug_time) and checks the tracks against the first-generation kernel (bit-identical)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = os.path.join(ROOT, "ros2_mono_vo_b200")
VDIR = os.path.join(PKG, "variants")

VARIANTS = {
    # name: (source, [flags])
    "ti1b5": ("lk.cu", ["-DMVO_LK2_TI_SMEM=1", "-DMVO_LK2_MINB=5"]),
    "ti1b6": ("lk.cu", ["-DMVO_LK2_TI_SMEM=1", "-DMVO_LK2_MINB=6"]),
    "ti0b5": ("lk.cu", ["-DMVO_LK2_TI_SMEM=0", "-DMVO_LK2_MINB=5"]),
    "ti0b6": ("lk.cu", ["-DMVO_LK2_TI_SMEM=0", "-DMVO_LK2_MINB=6"]),
    "ti1b4": ("lk.cu", ["-DMVO_LK2_TI_SMEM=1", "-DMVO_LK2_MINB=4"]),
    "w2b12": ("lk.cu", ["-DMVO_LK2_TI_SMEM=1", "-DMVO_LK2_WARPS=2", "-DMVO_LK2_MINB=12"]),
}


def build():
    from ros2_mono_vo_b200 import build as b
    b.build()
    os.makedirs(VDIR, exist_ok=True)
    for name, (src, flags) in VARIANTS.items():
        obj = os.path.join(VDIR, f"{name}_{src[:-3]}.o")
        cmd = [b.NVCC, *b.FLAGS, *flags, "-Xptxas=-v", "-c", os.path.join(b.CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.exit(r.stderr)
        info = [l for l in r.stderr.splitlines() if "track2" in l or "Used" in l or "spill" in l]
        k = [i for i, l in enumerate(info) if "track2" in l and "Compiling" in l]
        print(name, " | ".join(s.strip() for s in info[k[0] + 1:k[0] + 4]) if k else "")
        objs = [os.path.join(b.OBJ, s[:-3] + ".o") if s != src else obj for s in b._sources()]
        lib = os.path.join(VDIR, f"libmonovo_{name}.so")
        subprocess.run([b.NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "--shared", "-o", lib, *objs, "-lcudart"],
                       check=True)


def one(reps=20):
    import numpy as np
    import torch

    from oracle import synth
    from ros2_mono_vo_b200 import Context
    S, H, W, N = 32, 376, 1241, 2000
    seqs = [synth.synth_sequence(H, W, s, 3) for s in range(S)]
    K = seqs[0][1]
    dev = torch.from_numpy(np.stack([np.stack([seqs[s][0][f] for s in range(S)]) for f in range(3)])).cuda()
    ctx = Context(W, H, nfeatures=N, batch=S)
    for t in range(3):
        ctx.group_step(None, K, device_ptr=dev[t].data_ptr(), shape=(H, W))
    out = {}
    for impl in (1, 2):
        ctx.debug_set("lk_impl", impl)
        out[f"impl{impl}_ms"] = round(min(ctx.debug_time("lk_track", reps) for _ in range(3)), 4)
    ctx.close()
    # parity of the two kernels through the single-call ABI (interior, border and out-of-image points)
    f0, f1 = seqs[0][0][0], seqs[0][0][1]
    c1 = Context(W, H, nfeatures=N)
    kp, _ = c1.orb_detect_and_compute(f0)
    rng = np.random.default_rng(5)
    pts = np.concatenate([np.stack([kp["x"], kp["y"]], 1),
                          rng.uniform([-5, -5], [W + 5, H + 5], (600, 2)).astype(np.float32),
                          np.array([[0, 0], [W - 1, H - 1], [3.5, 200.25], [W - 2.5, 7.75]], np.float32)])
    res = {}
    for impl in (1, 2):
        c1.debug_set("lk_impl", impl)
        res[impl] = c1.lk_track(f0, f1, pts)
    c1.close()
    same = all(np.array_equal(a, b, equal_nan=True) for a, b in zip(res[1], res[2]))
    out["identical"] = bool(same)
    if not same:
        d = np.abs(res[1][0] - res[2][0]).max(axis=1)
        out["n_pos_diff"] = int((d > 0).sum())
        out["max_pos_diff"] = float(np.nanmax(d))
        out["n_status_diff"] = int((res[1][1] != res[2][1]).sum())
        out["n_err_diff"] = int((res[1][2] != res[2][2]).sum())
    print("RESULT", json.dumps(out))


def run():
    names = sys.argv[2:] or sorted(VARIANTS)
    for name in ["default"] + names:
        env = dict(os.environ)
        if name != "default":
            env["MVO_B200_LIB"] = os.path.join(VDIR, f"libmonovo_{name}.so")
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "one"], env=env, capture_output=True, text=True)
        line = [l for l in r.stdout.splitlines() if l.startswith("RESULT")]
        print(name, line[0][7:] if line else "FAILED\n" + r.stdout[-2000:] + r.stderr[-3000:], flush=True)


if __name__ == "__main__":
    {"build": build, "run": run, "one": one}[sys.argv[1]]()

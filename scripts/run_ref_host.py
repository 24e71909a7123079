"""Scratch driver: write a synthetic sequence file and run oracle/_ref/mono_vo_host on it (GPU box)."""
import json, os, struct, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import synth

def write_seq(path, frames, K):
    f0 = frames[0]
    cn = 1 if f0.ndim == 2 else f0.shape[2]
    with open(path, "wb") as f:
        f.write(b"MVOSEQ1\0")
        f.write(struct.pack("<4i", len(frames), f0.shape[0], f0.shape[1], cn))
        f.write(np.asarray(K, dtype=np.float64).tobytes())
        for fr in frames:
            f.write(np.ascontiguousarray(fr).tobytes())

if __name__ == "__main__":
    h, w, n = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    nfeat = sys.argv[4] if len(sys.argv) > 4 else "1000"
    frames, K, T = synth.synth_sequence(h, w, 0, n, return_poses=True)
    write_seq("/tmp/seq.bin", frames, K)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([os.path.join(root, "oracle/_ref/mono_vo_host"), "--seq", "/tmp/seq.bin", "--params",
                        os.path.join(root, "oracle/_ref/params.txt"), "--nfeatures", nfeat], capture_output=True, text=True)
    print(r.stdout[-6000:])
    print(r.stderr[-6000:], file=sys.stderr)
    print("T", T.tolist())

import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import orb_oracle as oo, synth
from ros2_mono_vo_b200 import Context
h, w, n, seed = 240, 320, 300, 3
img = synth.synth_frame(h, w, seed)
ctx = Context(w, h, nfeatures=n)
ctx.orb_detect_and_compute(img)
sm = ctx.orb_level(0, True).astype(int)
om = oo.fast_score_map(img)
om2 = om.copy(); om2[:30,:]=0; om2[-30:,:]=0; om2[:,:30]=0; om2[:,-30:]=0
d = sm - om2
print('nonzero gpu', np.count_nonzero(sm), 'oracle', np.count_nonzero(om2), 'mismatch', np.count_nonzero(d))
ys, xs = np.nonzero(d)
for y, x in list(zip(ys, xs))[:12]:
    print((x, y), 'gpu', sm[y, x], 'oracle', om2[y, x], 'tile', (x // 64, y // 32), 'local', (x % 64, y % 32))
print('mismatch by local y', np.bincount(ys % 32, minlength=32).tolist())
print('mismatch by local x', np.bincount(xs % 64, minlength=64).tolist())
print('mismatch by global y', np.bincount(ys, minlength=240).tolist())

import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import orb_oracle as oo, synth
from ros2_mono_vo_b200 import Context
h, w, n, seed = 240, 320, 300, 3
img = synth.synth_frame(h, w, seed)
ctx = Context(w, h, nfeatures=n)
kps, desc = ctx.orb_detect_and_compute(img)
pyr = oo.build_pyramid(img)
quotas = oo.level_quotas(n)
print('n', len(kps), 'quotas', quotas)
for l in range(8):
    lvl = ctx.orb_level(l)
    d = (lvl.astype(int) - pyr[l].astype(int))
    print('level', l, lvl.shape, 'pyr mismatches', np.count_nonzero(d), 'first', np.argwhere(d != 0)[:3].tolist())
    bl = ctx.orb_level(l, True); ob = oo.blur_orb(pyr[l])
    db = bl.astype(int) - ob.astype(int)
    print('   blur mismatches', np.count_nonzero(db), np.argwhere(db != 0)[:3].tolist())
    xs, ys, sc = ctx.orb_fast(l)
    oxs, oys, osc = oo.fast_nms(oo.fast_score_map(pyr[l]))
    hh, ww = pyr[l].shape
    m = (oxs >= 31) & (oxs < ww - 31) & (oys >= 31) & (oys < hh - 31)
    a = set(zip(xs.tolist(), ys.tolist(), sc.tolist())); b = set(zip(oxs[m].tolist(), oys[m].tolist(), osc[m].tolist()))
    print('   fast gpu', len(a), 'oracle', len(b), 'common', len(a & b), 'only gpu', sorted(a - b)[:4], 'only oracle', sorted(b - a)[:4])
okp, odesc = oo.orb_detect_and_compute(img, n)
for l in range(8):
    g = kps[kps['octave'] == l]; o = okp[okp['octave'] == l]
    ga = set(zip(g['x'].tolist(), g['y'].tolist())); oa = set(zip(o['x'].tolist(), o['y'].tolist()))
    print('level', l, 'kps gpu', len(g), 'oracle', len(o), 'common', len(ga & oa))
    if len(g) and len(o):
        print('   gpu first', g[:2].tolist()); print('   ora first', o[:2].tolist())

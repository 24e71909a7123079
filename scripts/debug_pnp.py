import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import pnp_oracle as po, synth
from ros2_mono_vo_b200 import Context
obj, img, K, rv, tv = synth.pnp_scene(2000, 1, 0.5, 0.05)
ctx = Context(1241, 376, nfeatures=2000)
ok, r, t, inl = ctx.solve_pnp_ransac(obj, img, K)
sub, mdl, cnt = ctx.pnp_hypotheses(100)
osub = po.sample_subsets(2000, 100)
print('ok', ok, 'inliers', len(inl), 'subsets equal', np.array_equal(sub, osub))
for it in range(8):
    s = osub[it]
    xn = po.normalize(img[s], K).astype(np.float32).astype(np.float64)
    R, tt = po.epnp(obj[s].astype(np.float64), xn)
    rvec = po.matrix_to_rodrigues(R)
    c = int((po.reproj_errors(obj, img, rvec, tt, K) <= np.float32(64)).sum())
    print(it, 'gpu count', cnt[it], 'oracle count', c, 'dR', np.abs(mdl[it, :9].reshape(3, 3) - R).max(), 'dt', np.abs(mdl[it, 9:] - tt).max())
    if it < 2:
        print('   gpu R', mdl[it, :9], 't', mdl[it, 9:]); print('   ora R', R.ravel(), 't', tt)

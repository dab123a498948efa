"""findEssentialMat timing: engine vs cv2 on the two-view bootstrap (sfm.py:307)."""
import sys, time
sys.path.insert(0, '.')
import cv2, numpy as np
import sfm_mvs_b200 as sfm
from sfm_mvs_b200 import synth
ctx = sfm.default_context()
K = synth.K_GUSTAV
for n, outl in ((1000, 0.2), (3000, 0.2), (3000, 0.5), (10000, 0.3)):
    p0, p1, _, _ = synth.two_view_pair(n, seed=n, outliers=outl)
    for _ in range(3):
        sfm.findEssentialMat(p0, p1, K, method=cv2.RANSAC, prob=0.999, threshold=0.4, ctx=ctx)
    t0 = time.perf_counter()
    for _ in range(10):
        E, m = sfm.findEssentialMat(p0, p1, K, method=cv2.RANSAC, prob=0.999, threshold=0.4, ctx=ctx)
    tg = (time.perf_counter() - t0) / 10
    ctx.set_profiling(True); before = ctx.profile()
    for _ in range(5):
        sfm.findEssentialMat(p0, p1, K, method=cv2.RANSAC, prob=0.999, threshold=0.4, ctx=ctx)
    after = ctx.profile(); ctx.set_profiling(False)
    prof = {k: round((after[k]['ms'] - before.get(k, dict(ms=0))['ms']) / 5, 4) for k in after}
    t0 = time.perf_counter()
    for _ in range(3):
        Ec, mc = cv2.findEssentialMat(p0, p1, K, method=cv2.RANSAC, prob=0.999, threshold=0.4)
    tc = (time.perf_counter() - t0) / 3
    print(f"n={n} outl={outl}: engine {tg*1e3:.3f} ms  cv2 {tc*1e3:.1f} ms  x{tc/tg:.0f}  mask equal {np.array_equal(m, mc)} "
          f"info {sfm.findEssentialMat.last_info}")
    print("   ", {k: v for k, v in prof.items() if 'essential' in k or 'misc' in k})

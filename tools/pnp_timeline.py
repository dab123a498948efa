"""Phase timings (SM cycles) of the EPnP kernel (hypothesis 0) and of the PnP tail kernel.
usage: SFM_PNP_TIMELINE=1 python tools/pnp_timeline.py [n]     (a clean problem, as in the bench's registration loop)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sfm_mvs_b200 as sfm
from sfm_mvs_b200 import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
K = synth.K_GUSTAV
rng = np.random.default_rng(0)
X = np.column_stack([rng.uniform(-3, 3, n), rng.uniform(-2, 2, n), rng.uniform(5, 11, n)])
R = sfm.rodrigues_to_matrix(np.array([0.02, 0.3, -0.01])); t = np.array([0.4, -0.1, 0.3])
Y = X @ R.T + t
p = np.column_stack([K[0, 0] * Y[:, 0] / Y[:, 2] + K[0, 2], K[1, 1] * Y[:, 1] / Y[:, 2] + K[1, 2]]) + rng.normal(0, 0.4, (n, 2))
ctx = sfm.Context(0)
for _ in range(3):
    ok, r, tt, inl, info = ctx.pnp_ransac(X.astype(np.float32), p.astype(np.float32), K)
print(info["iters_run"], info["refine_iters"], len(inl))

# the registration loop's fused tail (cluster of 8 CTAs): a short chain at the bench's descriptor count
from sfm_mvs_b200 import pipeline
scene = synth.orbit_scene(5, int(sys.argv[2]) if len(sys.argv) > 2 else 5000, seed=1)
pipeline.register_chain(scene, ctx)

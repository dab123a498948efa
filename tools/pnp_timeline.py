import sys; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np
import sfm_mvs_b200 as sfm
from test_gpu_pnp import _problem, K
ctx=sfm.Context(0)
X,p=_problem(1,n=1000)
for _ in range(3):
    ok,r,t,inl,info=ctx.pnp_ransac(X,p,K)
print(info)

import sys; sys.path.insert(0,'.')
import numpy as np, cv2, time
import sfm_mvs_b200 as sfm
from sfm_mvs_b200 import synth
sys.path.insert(0,'tests')
from test_gpu_pnp import _problem, K, D0
ctx=sfm.Context(0)
for seed in range(40):
    X,p=_problem(seed)
    okr,rr,tr,ir=cv2.solvePnPRansac(X,p,K,D0)
    ok,r,t,inl,info=ctx.pnp_ransac(X,p,K)
    a,b=set(inl.tolist()),set(ir[:,0].tolist()) if okr else set()
    pr,_=cv2.projectPoints(X[ir[:,0]],rr,tr,K,None); pm,_=cv2.projectPoints(X[ir[:,0]],r,t,K,None)
    print(seed,len(X),okr,ok,'cv',len(b),'mine',len(a),'jac %.3f'%(len(a&b)/max(1,len(a|b))),'dproj %.3f'%np.abs(pr-pm).max(), 'iters',info['iters_run'],'refine',info['refine_iters'])
ctx.set_profiling(True)
X,p=_problem(1,n=1000)
for _ in range(20): ctx.pnp_ransac(X,p,K)
print(ctx.profile())
t0=time.time()
for _ in range(50): ctx.pnp_ransac(X,p,K)
print('pnp_ransac ms/call',(time.time()-t0)/50*1e3)
t0=time.time()
for _ in range(50): cv2.solvePnPRansac(X,p,K,D0)
print('cv2 ms/call',(time.time()-t0)/50*1e3)

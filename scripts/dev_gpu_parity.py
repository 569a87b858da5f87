import numpy as np, sys, time
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bonmin_b200 import instances as I
from bonmin_b200.engine import Problem, Engine
from oracle import ipm
def check(m, XL, XU, x0, name):
    eng = Engine(Problem.from_model(m))
    t=time.time(); r = eng.solve_batch(XL, XU, x0); dt=time.time()-t
    o = ipm.solve_batch(m, XL, XU, x0, threads=8)
    same = (r['status']==o['status'])
    feas = o['status']==1
    rel = np.abs(r['obj']-o['obj'])/np.maximum(1,np.abs(o['obj']))
    print(name, 'B',len(r['obj']),'time',dt,'stats',eng.stats(),'status match',same.mean(),'max rel obj (feasible)', rel[feas&same].max() if (feas&same).any() else None,
          'iters gpu/oracle', r['iters'].mean(), o['iters'].mean(), 'xdiff', np.abs(r['x']-o['x'])[feas&same].max() if (feas&same).any() else None)
    if not same.all():
        bad=np.flatnonzero(~same)[:10]; print('  mismatch', bad, r['status'][bad], o['status'][bad], r['obj'][bad], o['obj'][bad])
    return r,o
m=I.mytoy()
XL=np.tile(m.xl,(3,1)); XU=np.tile(m.xu,(3,1)); XL[1,2]=1; XU[2,2]=0
r,o=check(m,XL,XU,m.x0,'mytoy'); print(r['obj'], r['status'], r['iters'], r['x'][0], r['duals'][0])
m=I.config_c2()
root=ipm.solve_batch(m,m.xl,m.xu,m.x0); xs=root['x'][0]
XL,XU=I.c3_children(m,xs,256)
check(m,XL,XU,xs,'c2')
m=I.config_c3()
root=ipm.solve_batch(m,m.xl,m.xu,m.x0); xs=root['x'][0]
XL,XU=I.c3_children(m,xs,592)
check(m,XL,XU,xs,'c3')

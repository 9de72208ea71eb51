import sys, time
sys.path.insert(0,'/root/repo')
import numpy as np
from inovesa_b200.params import kara_params
kw=dict(grid_size=512, steps_per_Ts=2000, bunch_current=300e-6)
p=kara_params(**kw)
mode=sys.argv[1]
if mode=='oracle':
    from oracle.pyoracle import Oracle
    s=Oracle(False).sim(p)
    obs=lambda: s.observe()[0]
else:
    from inovesa_b200 import capi
    capi.load()
    ctx=capi.Context(0, capi.ARITH_STRICT)
    s=capi.Simulation(ctx,p,1); s.set_radiation()
    obs=lambda: s.observe()[0]
t=time.time()
for k in range(8):
    s.step(1000)
    o=obs()
    print(mode, (k+1)*1000, "charge %.6f sigma_q %.4f sigma_p %.4f"%(o[0],o[3],o[6]), "%.0fs"%(time.time()-t), flush=True)

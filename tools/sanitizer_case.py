import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lfit_python_b200 import _cabi, workloads
eng=_cabi.Engine(0)
wl=workloads.config(2, ecl_per_band=2, n_ph=300)
wl.make_data(lambda p,x,w: eng.calc_flux(p,x,w))
wl.apply(eng)
th=wl.walkers(96, scatter=0.03)
th[5,0]=-1; th[7,1]=0.2
for what in (0,1,2):
    r=eng.log_prob(th, what=what, return_chisq=True)
print('ok', np.isfinite(r[0]).sum())
tot,comp=eng.calc_flux(wl.cv_pars(wl.p0,0), np.linspace(-0.5,0.5,700), np.full(700,0.0007), components=True)
print(tot[:3])
o,ok=eng.roche(_cabi.ROCHE_BSPOT,[0.1,0.2],[0.3,0.25]); print(o[0])
eng.close()

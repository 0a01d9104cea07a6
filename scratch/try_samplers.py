import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ppcseq_b200 as P
from ppcseq_b200 import inference, ppc, api
z = np.load('tests/golden/bundled_test53.npz')
m = P.NBModel(z['counts'], z['X'], z['exposure_rate'], int(z['K']))
t=time.time()
fit = inference.advi(m, output_samples=1000, iter=50000, tol_rel_obj=0.005, seed=3)
print('advi', time.time()-t, fit.info(8))
lay = m.layout
print('hyper means', fit.param_mean(0,3), fit.param_mean(lay.o_tail,3))
print('intercept', fit.param_mean(lay.o_intercept, 5), np.log(z['counts'][:5].mean(axis=1)))
print('slope', fit.slope())
lo, up, mean, sd = fit.ppc_summary(0.05, exact=True, seed=1)
fl = ppc.flags(m, lo, up, mean, fit.slope())
print('pass1 failed', fl['ppc_samples_failed'], fl['tot_deleterious_outliers'])
t=time.time()
fit2 = inference.sample_nuts(m, chains=3, iter=334+150, warmup=150, seed=5)
print('nuts', time.time()-t, fit2.info(8))
print('hyper means', fit2.param_mean(0,3), fit2.param_mean(lay.o_tail,3))
print('slope', fit2.slope())
lo, up, mean, sd = fit2.ppc_summary(0.05, exact=True, seed=1)
fl = ppc.flags(m, lo, up, mean, fit2.slope())
print('pass1 failed', fl['ppc_samples_failed'], fl['tot_deleterious_outliers'])

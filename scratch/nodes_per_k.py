import sys
sys.path.insert(0,'/root/repo')
import numpy as np, ctypes as C
from tests.host_emul import emul_lib
E = emul_lib.load()
tab = np.ones(E.M)
for name, ks, qmax in (('k_interp', np.logspace(np.log10(5e-6), 0, 250)[::25], 1e5), ('ml', np.logspace(np.log10(5e-6), 0, 250)[::25], 10.), ('1loop', np.exp(np.log(1e-5)+np.arange(1000)*np.log(1e6)/999)[::100], 1e5)):
    print(name)
    for k in ks:
        out, nn = E.ik(0, np.array([k]), tab, tab, 0, qmax=qmax)
        print('   k=%.3e nodes %d'%(k, nn))

"""First GPU contact: parity of the CUDA engine vs the oracle on a small Mackey-Glass ensemble, then a timing."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, sympy
from jitcdde_b200._symbols import *
from jitcdde_b200._codegen import generate
from jitcdde_b200 import _capi, _build
from oracle import oracle

beta,tau=sympy.symbols("beta tau")
p=generate([beta*y(0,t-tau)/(1+y(0,t-tau)**10)-0.1*y(0)],control_pars=[beta,tau])
past=(np.array([-1.0,0.0]), np.ones((2,1)), np.zeros((2,1)))

def run(N, mode, offs, cap=128, check=False):
    i=np.arange(N)
    pars=np.stack([0.15+0.2*((i*7919)%2500)/2499.0, 10+15*((i*104729)%4000)/3999.0],axis=1)
    E=_capi.Engine(1,1,p.n_sites,2,N,ring_capacity=cap)
    E.load_image(_build.build_image(p,mode=mode))
    E.set_past(*past); E.set_parameters(pars)
    E.set_integration_parameters(25.0, **oracle.DEFAULTS)
    out,st=E.step_on_discontinuities(pars[:,1:2].copy())
    t_after=E.get_t()
    res=[]
    E.synchronize(); t0=time.time()
    a0=E.sum_counters()[0]
    for o in offs:
        if check:
            out,st=E.integrate(t_after+o); res.append(out.copy())
        else:
            E.integrate(t_after+o, fetch=False)
    E.synchronize(); el=time.time()-t0
    a,r,f=E.sum_counters()
    print(f"N={N} mode={mode} steps={a-a0} time={el:.4f}s rate={(a-a0)/el:.3e} steps/s  status={np.unique(E.get_status())}", flush=True)
    if check:
        lib=oracle.build(p)
        ref=oracle.run_ensemble(lib,past,pars,25.0,1,pars[:,1:2].copy(),offs)
        res=np.array(res); ca,cr,cf=E.get_counters()
        print("  bit-exact states:",np.array_equal(res,ref['out']), "counts equal:",np.array_equal(ca,ref['counters'][:,0]), np.array_equal(cr,ref['counters'][:,1]),
              " max rel diff", np.max(np.abs(res-ref['out'])/np.abs(ref['out'])), flush=True)
    E.close()

offs=np.arange(0,200,10.0)
run(512,"verify",offs,check=True)
run(512,"fast",offs,check=True)
for N in (1<<16, 1<<20):
    run(N,"fast",np.arange(0,100,10.0))
    run(N,"verify",np.arange(0,100,10.0))

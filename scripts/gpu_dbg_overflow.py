import os, sys
sys.path.insert(0, "/root/repo")
import numpy as np
from jitcdde_b200 import _build, _capi
from tests import systems
from oracle import oracle
N = 1250000
prog = systems.program(systems.mackey_glass()); pars = systems.mg_grid(N)
E = _capi.Engine(1, 1, prog.n_sites, 2, N, ring_capacity=128); E.load_image(_build.build_image(prog))
E.set_past(np.array([-1.0, 0.0]), np.ones((2, 1)), np.zeros((2, 1))); E.set_parameters(pars)
E.set_integration_parameters(25.0, **oracle.DEFAULTS)
E.step_on_discontinuities(pars[:, 1:2].copy())
k = 0
for n in [int(x) for x in sys.argv[1].split(",")]:
    E.integrate_samples(30.0+10.0*np.arange(k+1, k+n+1), fetch=False); k += n
    st = E.get_status()
    print("after", k, "samples: failed", int((st != 0).sum()), np.unique(st), flush=True)
    if st.any():
        bad = np.nonzero(st)[0][:5]
        print("members", bad, pars[bad], "t", E.get_t()[bad], "dt", E.get_dt()[bad])
        for b in bad[:2]:
            times, states, diffs = E.get_state(int(b))
            print(len(times), times[:3], times[-3:])
        break

import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, warnings
warnings.simplefilter("ignore")
from jitcdde_b200 import jitcdde_lyap
from tests import scenarios, systems
N = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
past, k, offsets = scenarios.c3_inputs(N)
system = systems.two_roessler(control=True)
DDE = jitcdde_lyap(system["f"], n_lyap=4, control_pars=system["control_pars"], n_members=N, verbose=True, ring_capacity=1024)
DDE.add_past_points(past); DDE.set_integration_parameters(**systems.LYAP_TEST_PARAMETERS); DDE.set_parameters(k)
DDE.step_on_discontinuities()
t0 = DDE.t
DDE.integrate(t0+20.0)
E = DDE._engine
print("ring", E.ring_capacity)
base = np.asarray(E.get_t())
for i in range(3):
    w0 = time.perf_counter()
    state, lyaps, dt, status = E.integrate_lyap(base+10.0*(i+1))
    print(i, "status", np.unique(status, return_counts=True), f"{1e3*(time.perf_counter()-w0):.1f} ms", "accepted", E.sum_counters()[0])
    if status.any():
        bad = np.nonzero(status)[0][:3]
        for b in bad:
            t, s_, d = E.get_state(int(b)); print("member", b, k[b], "anchors", len(t), t[0], t[-1], "t", E.get_t()[b])
        break

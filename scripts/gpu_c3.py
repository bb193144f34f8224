"""C3 (two Roesslers + 4 exponents, n = 30) with the launch shape from JITCDDE_B200_LAUNCH; prints member-steps/s of the sampling loop."""
import os, sys, time, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.simplefilter("ignore")
import numpy as np
from jitcdde_b200 import jitcdde_lyap
from tests import scenarios, systems
N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
past, k, offsets = scenarios.c3_inputs(N)
system = systems.two_roessler(control=True)
DDE = jitcdde_lyap(system["f"], n_lyap=4, control_pars=system["control_pars"], n_members=N, verbose=False, ring_capacity=int(os.environ.get("CAP", 256)))
DDE.add_past_points(past); DDE.set_integration_parameters(**systems.LYAP_TEST_PARAMETERS); DDE.set_parameters(k)
DDE.step_on_discontinuities(); t0 = DDE.t
DDE.integrate(t0+10)
a0 = DDE.counters()[0].sum(); w0 = time.perf_counter()
for o in (20, 30, 40, 50): DDE.integrate(t0+o)
el = time.perf_counter()-w0; a1 = DDE.counters()[0].sum()
print(f"LAUNCH={os.environ.get('JITCDDE_B200_LAUNCH','default')} N={N}: {a1-a0} accepted steps in {el:.3f}s → {(a1-a0)/el:.3e} member-steps/s", flush=True)

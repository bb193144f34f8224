"""
Lyapunov exponents of two delay-coupled Rössler oscillators (the reference's examples/two_Roesslers.py and tests/test_lyaps.py) for
a whole range of coupling strengths at once: an ensemble of jitcdde_lyap systems, one coupling per member.

    python examples/two_roesslers_lyap.py [members]
"""
import sys

import numpy as np
import sympy

from jitcdde_b200 import jitcdde_lyap, t, y

members = int(sys.argv[1]) if len(sys.argv) > 1 else 64
omega, delay, k = np.array([0.88167179, 0.87768425]), 4.5, sympy.Symbol("k")
f = [
	omega[0]*(-y(1)-y(2)),
	omega[0]*(y(0)+0.165*y(1)),
	omega[0]*(0.2+y(2)*(y(0)-10.0)),
	omega[1]*(-y(4)-y(5)) + k*(y(0, t-delay)-y(3)),
	omega[1]*(y(3)+0.165*y(4)),
	omega[1]*(0.2+y(5)*(y(3)-10.0)),
]
n_lyap = 4
DDE = jitcdde_lyap(f, n_lyap=n_lyap, control_pars=[k], n_members=members, verbose=False)
DDE.constant_past(np.array([0.1, 0.2, 0.3, 0.4, 0.5, 0.6]))
DDE.set_integration_parameters(rtol=1e-7, atol=1e-7, first_step=0.1)
DDE.set_parameters(np.linspace(0.0, 0.5, members)[:, None])
DDE.step_on_discontinuities()

weights, local_exponents = [], []
start = float(np.max(DDE.t))                           # (an ensemble has one time per member; after the discontinuity steps they agree)
for time in start+np.arange(10, 2010, 10):
	state, lyaps, weight = DDE.integrate(time)           # per member: state (members, 6), local exponents (members, 4), weights (members,)
	local_exponents.append(lyaps)
	weights.append(weight)
weights, local_exponents = np.array(weights)[100:], np.array(local_exponents)[100:]
exponents = (local_exponents*weights[:, :, None]).sum(axis=0)/weights.sum(axis=0)[:, None]
for coupling, spectrum in zip(np.linspace(0.0, 0.5, members)[::max(1, members//8)], exponents[::max(1, members//8)]):
	print(f"k = {coupling:.3f}   λ = " + "  ".join(f"{x:+.4f}" for x in spectrum))

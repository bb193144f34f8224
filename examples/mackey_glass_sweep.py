"""
A (β, τ) parameter sweep of the Mackey–Glass equation as ONE ensemble: every member is the system of the reference's
examples/mackey_glass.py with its own control parameters, all members are integrated by one kernel launch per batch of samples.

    python examples/mackey_glass_sweep.py [members]

Needs a B200 (the engine has no CPU fallback).
"""
import sys

import numpy as np
import sympy

from jitcdde_b200 import jitcdde, t, y

members = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
beta, tau = sympy.symbols("beta tau")
f = [beta*y(0, t-tau)/(1+y(0, t-tau)**10) - 0.1*y(0)]

rng = np.random.default_rng(0)
parameters = np.stack([rng.uniform(0.15, 0.35, members), rng.uniform(10, 25, members)], axis=1)      # one (β, τ) per member

DDE = jitcdde(f, control_pars=[beta, tau], max_delay=25.0, n_members=members, verbose=False)
DDE.constant_past([1.0])
DDE.set_parameters(parameters)
DDE.delays = parameters[:, 1:2]          # per-member numeric delays: step_on_discontinuities steps to each member's own τ
DDE.step_on_discontinuities()

# 100 samples, 10 time units apart, in one launch: (samples, members, n) — the same steps and states as a loop of integrate() calls
series = DDE.integrate_samples(30.0+10.0*np.arange(1, 101))
amplitude = series[50:, :, 0].max(axis=0)-series[50:, :, 0].min(axis=0)
accepted, rejected, _ = DDE.counters()
print(f"{members} members, {int(accepted.sum())} accepted steps; oscillation amplitude from {amplitude.min():.3f} to {amplitude.max():.3f}")

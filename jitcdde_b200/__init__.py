"""
jitcdde_b200 — B200-native integration engine for delay differential equations with
the user-facing API of neurophysik/jitcdde (reference export list:
/root/reference/jitcdde/__init__.py:1).

    from jitcdde_b200 import jitcdde, y, t
    DDE = jitcdde([0.25*y(0,t-15)/(1+y(0,t-15)**10) - 0.1*y(0)])
    DDE.constant_past([1.0]); DDE.step_on_discontinuities(); DDE.integrate(100.)

The integrator runs as CUDA kernels for sm_100a behind a C ABI
(include/jitcdde_b200.h); there is no CPU fallback. Ensembles of trajectories
(`n_members=N`) are the data-parallel axis. Scope and design: DESIGN.md.
"""

from ._symbols import anchors, current_y, dy, input, past_dy, past_y, quadrature, t, y  # noqa: A004
from ._jitcdde import UnsuccessfulIntegration, jitcdde, jitcdde_input, jitcdde_lyap, test, _find_max_delay, _propagate_delays
from ._codegen import get_delays as _get_delays
from ._capi import pinned_empty


from ._dist import gather_members, shard
from ._network import jitcdde_network


from ._transversal import GroupHandler, jitcdde_restricted_lyap, jitcdde_transversal_lyap

__all__ = [
	"UnsuccessfulIntegration", "anchors", "current_y", "dy", "input", "jitcdde", "jitcdde_input", "jitcdde_lyap", "jitcdde_restricted_lyap", "jitcdde_transversal_lyap",
	"past_dy", "past_y", "quadrature", "t", "test", "y", "jitcdde_network", "shard", "gather_members", "pinned_empty",
]

"""
Host-side logic of the code generator and the build policy (no GPU, no compiler): blockwise Lyapunov programs, the
restated GroupHandler, and which access path / stepper an image is built with.
"""

import numpy as np
import sympy

from jitcdde_b200 import _build, _codegen, t, y
from jitcdde_b200._transversal import GroupHandler
from tests import systems


def test_lyapunov_program_is_generated_blockwise():
	"""tangent blocks are printed once and instantiated per block (one-thread engines) or per lane (lane groups)"""
	program, n_basic = systems.lyap_program(systems.two_roessler(control=True), 4)
	assert (program.n, n_basic, program.group_lanes, program.n_sites) == (30, 6, 5, 1)
	body = program.body
	assert body.count("JDDE_ANCHORS(") == 1                                  # one lookup, made by everyone, ahead of the branches
	assert body.index("JDDE_ANCHORS(") < body.index("#ifndef JDDE_GROUP", 10)
	for offset in (6, 12, 18, 24):
		assert f"#define JDDE_TOFF {offset}\n" in body
	one_thread, group = body.split("#else\n")
	assert group.count("JDDE_SET_TDY(0,") == 1 and one_thread.count("JDDE_SET_TDY(0,") == 4
	assert "if (JDDE_GROUP_IS_BASIC)" in group
	# the lane-group stepper is chosen for it, and neither shared-memory window of the one-thread engine
	assert _build.use_group(program) == 5 and _build.use_window(program) == 0 and _build.group_window(program) == 1


def test_blockwise_generation_falls_back_when_blocks_differ():
	"""zero padding (jitcdde_restricted_lyap) or a hand-made extended system: blocks are not copies of each other"""
	f = [-y(0) + y(1, t-1), -y(1), -2*y(2), -3*y(3), -4*y(4), -5*y(5)]
	program = _codegen.generate(f, blocks=(2, 2))
	assert program.group_lanes == 0 and "JDDE_TOFF" not in program.body
	f = [-y(0) + y(1, t-1), -y(1), -2*y(2), -3*y(3) + y(4), -2*y(4), -3*y(5)]       # a block that reads another block
	assert _codegen.generate(f, blocks=(2, 2)).group_lanes == 0
	_, program, _ = systems.projected_program("restricted")
	assert program.group_lanes == 0 and program.n == 60


def test_build_policy_access_paths():
	mg = systems.program(systems.mackey_glass())
	assert _build.use_window(mg) == 2 and _build.launch_shape(mg) == (32, 28)          # one 32-byte anchor per sector: per-thread window, by value; 72 registers
	neutral = systems.program(systems.neutral())
	assert _build.use_window(neutral) == 3 and _build.launch_shape(neutral) == (64, 6)  # 13 doubles per anchor: pointer window in shared memory
	kuramoto = systems.program(systems.kuramoto())
	assert _build.use_window(kuramoto) == 0                                            # 30 lookup sites: too much shared memory per thread
	ode = systems.program(systems.fitzhugh_nagumo_ode())
	assert ode.n_sites == 0 and _build.use_window(ode) == 0


def test_group_handler_transformation():
	"""differences of consecutive members + vanishing projection onto the manifold determine the tangent vector"""
	G = GroupHandler(([0, 2, 4], [1, 3], [5]))
	assert G.main_indices == [0, 1, 5] and G.tangent_indices == [2, 3, 4]
	assert [G.map_to_main(i) for i in range(6)] == [0, 1, 0, 1, 0, 5]
	z = sympy.symbols("z0:6")
	dy = G.back_transform(list(z))
	for group in G.groups:
		assert sympy.simplify(sum(dy[i] for i in group)) == 0                          # no component along the manifold
		for previous, member in zip(group, group[1:]):
			assert sympy.simplify(dy[previous]-dy[member]-z[member]) == 0                # z = difference of consecutive members
	entries = list(G.iterate("abcdef"))
	assert entries == [0, 1, ("a", "c"), ("b", "d"), ("c", "e"), 2]


def test_transversal_system_has_no_coupling_in_the_synchronised_dynamics():
	system = systems.synchronised_fhn()
	DDE, program, projector = systems.projected_program("transversal")
	k = system["control_pars"][0]
	f = DDE.f_sym()
	assert not f[0].has(k) and not f[3].has(k)                                          # k·(y1−y0) vanishes on the manifold
	assert f[1].has(k) and f[2].has(k)                                                 # but drives the differences
	assert projector["indices"] == [1, 2, 4, 5] and DDE.main_indices == [0, 3]
	state = np.array([0.3, 0.1])
	DDE.constant_past(state, 0.0)
	assert DDE.past[0].state.shape == (6,) and np.allclose(DDE.past[0].state[[0, 3]], state)
	assert np.isclose(np.linalg.norm(DDE.past[0].state[[1, 2, 4, 5]]), 1.0)

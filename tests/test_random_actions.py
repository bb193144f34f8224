"""
Restatement of the reference's tests/compare_C_and_Python_core.py: the reference subjects its C core and its Python core to
the same random series of core calls and compares the results. Here the two sides are the REFERENCE'S C CORE (compiled
unmodified, oracle/build_ref.py) and THIS ENGINE's functions (csrc/jdde_engine.cuh, compiled for the host behind the
per-step test entry points of tests/host_twin/twin.cpp). Both compute in IEEE double without contraction, so unlike the
reference's own comparison (rtol 1e-3) this one demands identical bits.

The protocol is the reference's (:41-212): two Rösslers plus two idle components (n = 8, n_basic = 2), a random past,
two steps, then ten random actions out of get_next_step, get_t, get_recent_state, get_current_state, get_full_state,
get_p, accept_step, forget, check_new_y_diff, past_within_step, orthonormalise, remove_projections (or
remove_state/diff_component), normalise_indices, truncate, apply_jump.
"""

import ctypes
import random

import numpy as np
import pytest

from jitcdde_b200 import _capi, _codegen, t, y
from oracle import build_ref, oracle
from tests import twin as twin_module

N, N_BASIC, DELAY = 8, 2, 4.5
TANGENT_INDICES = [1, 4, 6]
REF_NAME = "ref_random_actions"


def system():
	omega = np.array([0.88167179, 0.87768425])
	k = 0.25
	return [
		omega[0] * (-y(1) - y(2)),
		omega[0] * (y(0) + 0.165 * y(1)),
		omega[0] * (0.2 + y(2) * (y(0) - 10.0)),
		omega[1] * (-y(4) - y(5)) + k * (y(0, t-DELAY) - y(3)),
		omega[1] * (y(3) + 0.165 * y(4)),
		omega[1] * (0.2 + y(5) * (y(3) - 10.0)),
		0,
		0,
	]


def reference_core(program, past, name=REF_NAME, n_basic=N_BASIC, tangent_indices=TANGENT_INDICES):
	if not build_ref.have_ref(name):
		if not build_ref.reference_available():
			pytest.skip("reference core not built and /root/reference not present")
		build_ref.build_ref(program, name, n_basic=n_basic, tangent_indices=tangent_indices)
	return build_ref.load_ref(name).dde_integrator([(float(a[0]), a[1].copy(), a[2].copy()) for a in past])


class EngineCore:
	"""the engine's functions behind the method names of the reference's dde_integrator"""

	def __init__(self, program, past, n=N, n_basic=N_BASIC, max_delay=DELAY):
		dp = ctypes.POINTER(ctypes.c_double)
		self.n = n
		self.lib = twin_module.build_twin(program, n_basic=n_basic)
		for name, restype, argtypes in (
				("twin_step_begin", ctypes.c_int, [ctypes.c_void_p]),
				("twin_get_next_step", None, [ctypes.c_void_p, ctypes.c_double]),
				("twin_get_t", ctypes.c_double, [ctypes.c_void_p]),
				("twin_past_within_step", ctypes.c_double, [ctypes.c_void_p]),
				("twin_get_p", ctypes.c_double, [ctypes.c_void_p, ctypes.c_double, ctypes.c_double]),
				("twin_check_new_y_diff", ctypes.c_int, [ctypes.c_void_p, ctypes.c_double, ctypes.c_double]),
				("twin_accept_step", None, [ctypes.c_void_p]),
				("twin_forget", None, [ctypes.c_void_p, ctypes.c_double]),
				("twin_get_recent_state", None, [ctypes.c_void_p, ctypes.c_double, dp]),
				("twin_get_current_state", None, [ctypes.c_void_p, dp]),
				("twin_get_full_state", ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, dp, dp, dp]),
				("twin_truncate", None, [ctypes.c_void_p, ctypes.c_double]),
				("twin_apply_jump", None, [ctypes.c_void_p, dp, ctypes.c_double, ctypes.c_double, dp, dp]),
				("twin_orthonormalise", None, [ctypes.c_void_p, ctypes.c_int, ctypes.c_double, dp]),
				("twin_remove_projections", ctypes.c_double, [ctypes.c_void_p, ctypes.c_double, dp]),
				("twin_remove_component", None, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]),
				("twin_normalise_indices", ctypes.c_double, [ctypes.c_void_p, ctypes.c_double, ctypes.c_int, ctypes.POINTER(ctypes.c_int)]),
			):
			if hasattr(self.lib, name):      # the Lyapunov entries exist only for n_basic < n
				function = getattr(self.lib, name)
				function.restype, function.argtypes = restype, argtypes
		saved = _capi.load_library, _capi.Engine.load_image
		_capi.load_library = lambda path=None: self.lib
		_capi.Engine.load_image = lambda engine, path: None
		try:
			self.engine = _capi.Engine(n, n_basic, program.n_sites, 0, 1, ring_capacity=256)
		finally:
			_capi.load_library, _capi.Engine.load_image = saved
		self.engine.set_past(np.array([a[0] for a in past]), np.array([a[1] for a in past]), np.array([a[2] for a in past]))
		self.engine.set_integration_parameters(max_delay, **oracle.DEFAULTS)
		self.h = self.engine._h
		assert self.lib.twin_step_begin(self.h) == 0

	@staticmethod
	def _p(array):
		return array.ctypes.data_as(ctypes.POINTER(ctypes.c_double))

	def get_next_step(self, dt): self.lib.twin_get_next_step(self.h, dt)
	def get_t(self): return self.lib.twin_get_t(self.h)
	def get_p(self, atol, rtol): return self.lib.twin_get_p(self.h, atol, rtol)
	def check_new_y_diff(self, atol, rtol): return bool(self.lib.twin_check_new_y_diff(self.h, atol, rtol))
	def accept_step(self): self.lib.twin_accept_step(self.h)
	def forget(self, delay): self.lib.twin_forget(self.h, delay)
	def truncate(self, time): self.lib.twin_truncate(self.h, time)

	@property
	def past_within_step(self): return self.lib.twin_past_within_step(self.h)

	def get_recent_state(self, time):
		out = np.empty(self.n)
		self.lib.twin_get_recent_state(self.h, time, self._p(out))
		return out

	def get_current_state(self):
		out = np.empty(self.n)
		self.lib.twin_get_current_state(self.h, self._p(out))
		return out

	def get_full_state(self):
		times, states, diffs = np.empty(256), np.empty((256, self.n)), np.empty((256, self.n))
		count = self.lib.twin_get_full_state(self.h, 256, self._p(times), self._p(states), self._p(diffs))
		return [(times[k], states[k], diffs[k]) for k in range(count)]

	def apply_jump(self, change, time, width):
		change = np.ascontiguousarray(change, dtype=float)
		minima, maxima = np.empty(self.n), np.empty(self.n)
		self.lib.twin_apply_jump(self.h, self._p(change), time, width, self._p(minima), self._p(maxima))
		return minima, maxima

	def orthonormalise(self, n_lyap, delay):
		norms = np.empty(n_lyap)
		self.lib.twin_orthonormalise(self.h, n_lyap, delay, self._p(norms))
		return norms

	def remove_projections(self, delay, vectors):
		(state, diff), = vectors
		flat = np.ascontiguousarray(np.concatenate([state, diff]), dtype=float)
		return self.lib.twin_remove_projections(self.h, delay, self._p(flat))

	def remove_state_component(self, index): self.lib.twin_remove_component(self.h, 0, int(index))
	def remove_diff_component(self, index): self.lib.twin_remove_component(self.h, 1, int(index))

	def normalise_indices(self, delay):
		indices = (ctypes.c_int*len(TANGENT_INDICES))(*TANGENT_INDICES)
		return self.lib.twin_normalise_indices(self.h, delay, len(TANGENT_INDICES), indices)


def same(a, b, what):
	a, b = np.asarray(a), np.asarray(b)
	assert a.shape == b.shape and np.array_equal(a, b), f"{what}: {a} != {b}"


def run_realisation(program, rng, pyrng, log):
	past_rng = np.random.default_rng(rng.integers(1000000))
	times = np.sort(past_rng.uniform(-10, 0, past_rng.integers(2, 10)))
	past = [(time, past_rng.random(N), 0.1*past_rng.random(N)) for time in times]
	R, E = reference_core(program, past), EngineCore(program, past)

	def both(name, *args):
		return getattr(R, name)(*args), getattr(E, name)(*args)

	def get_next_step():
		both("get_next_step", rng.uniform(1e-5, 1e-3))

	def get_t():
		same(*both("get_t"), "get_t")

	def get_recent_state():
		same(*both("get_recent_state", R.get_t()+rng.uniform(-0.1, 0.1)), "get_recent_state")

	def get_current_state():
		same(*both("get_current_state"), "get_current_state")

	def get_full_state():
		A, B = both("get_full_state")
		assert len(A) == len(B)
		for a, b in zip(A, B):
			for k in range(3):
				same(a[k], b[k], "get_full_state")

	def get_p():
		same(*both("get_p", 10**rng.uniform(-10, -5), 10**rng.uniform(-10, -5)), "get_p")

	def accept_step():
		both("accept_step")

	def forget():
		both("forget", DELAY)

	def check_new_y_diff():
		same(*both("check_new_y_diff", 10**rng.uniform(-10, -5), 10**rng.uniform(-10, -5)), "check_new_y_diff")

	def past_within_step():
		same(R.past_within_step, E.past_within_step, "past_within_step")

	def orthonormalise():
		same(*both("orthonormalise", 3, rng.uniform(0.1*DELAY, DELAY)), "orthonormalise")

	def remove_projections():
		if rng.random() > 0.1:
			d = rng.uniform(0.1*DELAY, DELAY)
			vector = tuple(rng.uniform(-1, 1, (2, 2))) if rng.integers(0, 2) else tuple(rng.integers(-1, 2, (2, 2)).astype(float))
			same(*both("remove_projections", d, [vector]), "remove_projections")
		else:
			both("remove_state_component" if rng.integers(0, 2) else "remove_diff_component", int(rng.integers(0, 2)))

	def normalise_indices():
		same(*both("normalise_indices", rng.uniform(0.1*DELAY, DELAY)), "normalise_indices")

	def reduced_interval():
		full = R.get_full_state()
		return 0.9*full[0][0]+0.1*full[-1][0], 0.1*full[0][0]+0.9*full[-1][0]

	def truncate_past():
		accept_step()
		both("truncate", rng.uniform(*reduced_interval()))

	def apply_jump():
		accept_step()
		time = rng.uniform(*reduced_interval())
		A, B = both("apply_jump", rng.normal(0, 0.1, N), time, 0.1)
		same(A[0], B[0], "apply_jump minima")
		same(A[1], B[1], "apply_jump maxima")

	actions = [get_next_step, get_t, get_recent_state, get_current_state, get_full_state, get_p, accept_step, forget, check_new_y_diff,
		past_within_step, orthonormalise, remove_projections, normalise_indices, truncate_past, apply_jump]
	get_next_step()
	get_next_step()
	for _ in range(10):
		action = pyrng.sample(actions, 1)[0]
		log.append(action.__name__)
		action()
	get_full_state()       # whatever the actions were, the histories end up identical
	get_p()
	E.engine.close()


def test_random_core_actions_match_the_reference_core():
	program = _codegen.generate(system())
	rng = np.random.default_rng(seed=42)
	pyrng = random.Random(int(rng.integers(1_000_000)))
	seen = set()
	for realisation in range(150):
		log = []
		try:
			run_realisation(program, rng, pyrng, log)
		except AssertionError as error:
			raise AssertionError(f"realisation {realisation}, actions {log}: {error}") from error
		seen.update(log)
	assert len(seen) == 15      # every kind of action occurred


def test_random_core_actions_with_past_within_step():
	"""the reference's second protocol (tests/compare_C_and_Python_core_2.py): Mackey–Glass with an additional tiny delay, steps of
	up to 1 — lookups reach into the step under way (past_within_step, extrapolation beyond the tentative anchor)"""
	tau, tiny = 15, 1e-30
	program = _codegen.generate([0.25*y(0, t-tau)/(1.0+y(0, t-tau)**10) - 0.1*y(0, t-tiny)])
	rng = np.random.default_rng(seed=7)
	pyrng = random.Random(11)
	seen = set()
	for realisation in range(150):
		past = [(time, rng.random(1), rng.random(1)) for time in (0.0, 0.5, 2.0)]
		R = reference_core(program, past, name="ref_random_actions_2", n_basic=None, tangent_indices=())
		E = EngineCore(program, past, n=1, n_basic=1, max_delay=tau)
		log = []

		def both(name, *args):
			return getattr(R, name)(*args), getattr(E, name)(*args)

		def reduced_interval():
			full = R.get_full_state()
			return 0.9*full[0][0]+0.1*full[-1][0], 0.1*full[0][0]+0.9*full[-1][0]

		def truncate_past():
			both("accept_step")
			both("truncate", rng.uniform(*reduced_interval()))

		def apply_jump():
			both("accept_step")
			A, B = both("apply_jump", np.array([rng.normal(0, 0.1)]), rng.uniform(*reduced_interval()), 0.1)
			same(A[0], B[0], "apply_jump minima")
			same(A[1], B[1], "apply_jump maxima")

		def get_full_state():
			A, B = both("get_full_state")
			assert len(A) == len(B)
			for a, b in zip(A, B):
				for k in range(3):
					same(a[k], b[k], "get_full_state")

		actions = dict(
			get_next_step=lambda: both("get_next_step", rng.uniform(1e-5, 1)),
			get_t=lambda: same(*both("get_t"), "get_t"),
			get_recent_state=lambda: same(*both("get_recent_state", R.get_t()+rng.uniform(-0.1, 0.1)), "get_recent_state"),
			get_current_state=lambda: same(*both("get_current_state"), "get_current_state"),
			get_full_state=get_full_state,
			get_p=lambda: same(*both("get_p", 10**rng.uniform(-10, -5), 10**rng.uniform(-10, -5)), "get_p"),
			accept_step=lambda: both("accept_step"),
			forget=lambda: both("forget", tau),
			check_new_y_diff=lambda: same(*both("check_new_y_diff", 10**rng.uniform(-10, -5), 10**rng.uniform(-10, -5)), "check_new_y_diff"),
			past_within_step=lambda: same(R.past_within_step, E.past_within_step, "past_within_step"),
			truncate_past=truncate_past,
			apply_jump=apply_jump,
		)
		try:
			actions["get_next_step"]()
			actions["get_next_step"]()
			for _ in range(30):
				name = pyrng.choice(sorted(actions))
				log.append(name)
				actions[name]()
			get_full_state()
			actions["past_within_step"]()
		except AssertionError as error:
			raise AssertionError(f"realisation {realisation}, actions {log}: {error}") from error
		seen.update(log)
		E.engine.close()
	assert len(seen) == len(actions)

"""
oracle.py — Python access to the plain-C oracle (dde_oracle.c). TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import
this module; nothing under jitcdde_b200/ does.

`build(program)` compiles dde_oracle.c for one generated right-hand side
(jitcdde_b200._codegen.RHSProgram) with gcc into oracle/_build/ and returns a
ctypes handle; `Trajectory` wraps one trajectory with the method names of the
reference's Python API (/root/reference/jitcdde/_jitcdde.py).
"""

import ctypes
import hashlib
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_BUILD = os.path.join(_HERE, "_build")
_CSRC = os.path.join(os.path.dirname(_HERE), "jitcdde_b200", "csrc")

#: verification flags: no fast-math, no FMA contraction (hard part 1)
VERIFY_FLAGS = ["-O2", "-ffp-contract=off"]
#: timing flags for the CPU-baseline leg
FAST_FLAGS = ["-O3", "-DJDDE_LIBM"]   # no -march=native: the prebuilt .so must also run on the GPU box's host CPU

_c_double_p = ctypes.POINTER(ctypes.c_double)


class Params(ctypes.Structure):
	"""jo_params in dde_oracle.c; defaults = /root/reference/jitcdde/_jitcdde.py:621-638."""
	_fields_ = [
		("atol", ctypes.c_double), ("rtol", ctypes.c_double), ("first_step", ctypes.c_double),
		("min_step", ctypes.c_double), ("max_step", ctypes.c_double),
		("decrease_threshold", ctypes.c_double), ("increase_threshold", ctypes.c_double),
		("safety_factor", ctypes.c_double), ("max_factor", ctypes.c_double), ("min_factor", ctypes.c_double),
		("pws_factor", ctypes.c_double), ("pws_atol", ctypes.c_double), ("pws_rtol", ctypes.c_double),
		("pws_base_increase_chance", ctypes.c_double),
		("pws_max_iterations", ctypes.c_int32), ("root_mode", ctypes.c_int32),
		("fuzzy_seed", ctypes.c_int32), ("pad_", ctypes.c_int32),
	]

DEFAULTS = dict(
	atol=1e-10, rtol=1e-5, first_step=1.0, min_step=1e-10, max_step=10.0,
	decrease_threshold=1.1, increase_threshold=0.5, safety_factor=0.9, max_factor=5.0, min_factor=0.2,
	pws_factor=3, pws_atol=0.0, pws_rtol=1e-5, pws_max_iterations=10, pws_base_increase_chance=0.1,
)


def make_params(root_mode=1, fuzzy_seed=0, **kwargs):
	values = dict(DEFAULTS)
	values.update(kwargs)
	# same clipping as the reference, _jitcdde.py:688-693
	if values["first_step"] > values["max_step"]:
		values["first_step"] = values["max_step"]
	if values["min_step"] > values["first_step"]:
		values["min_step"] = values["first_step"]
	return Params(root_mode=root_mode, fuzzy_seed=int(fuzzy_seed), **values)


def _arr(x, shape=None):
	a = np.ascontiguousarray(x, dtype=np.float64)
	if shape is not None:
		a = a.reshape(shape)
	return a


def _p(a):
	return a.ctypes.data_as(_c_double_p)


def build(program, n_basic=None, flags=VERIFY_FLAGS, openmp=False, tag=""):
	"""Compile the oracle for `program`; returns the loaded ctypes library."""
	os.makedirs(_BUILD, exist_ok=True)
	n_basic = n_basic or program.n
	name = f"oracle_{program.key}_{n_basic}{tag}_{'omp' if openmp else 'st'}_{hashlib.sha256(repr(tuple(flags)).encode()).hexdigest()[:8]}"
	rhs = os.path.join(_BUILD, f"rhs_{program.key}.inc")
	so = os.path.join(_BUILD, name + ".so")
	src = os.path.join(_HERE, "dde_oracle.c")
	if not os.path.exists(rhs):
		with open(rhs, "w") as f:
			f.write(program.body)
	deps = [src, os.path.join(_CSRC, "jdde_math.h")]
	h = hashlib.sha256()
	for d in deps:
		with open(d, "rb") as f:
			h.update(f.read())
	stamp = h.hexdigest()
	try:
		with open(so + ".stamp") as f:
			current = os.path.exists(so) and f.read().strip() == stamp
	except OSError:
		current = False
	if not current:
		tmp = f"{so}.{os.getpid()}.tmp"
		cmd = [
			"gcc", "-std=c11", *flags, "-fPIC", "-shared", "-Wall", "-Wno-unused-function", "-Wno-unused-variable", "-Wno-unknown-pragmas",
			f"-DJO_N={program.n}", f"-DJO_NBASIC={n_basic}", f"-DJO_NSITES={program.n_sites}",
			f"-DJO_NPARAMS={program.n_params}", f'-DJO_RHS_FILE="{rhs}"',
			f"-I{_CSRC}", src, "-o", tmp, "-lm",
		]
		if openmp:
			cmd.insert(2, "-fopenmp")
		subprocess.run(cmd, check=True)
		os.replace(tmp, so)
		with open(so + ".stamp", "w") as f:
			f.write(stamp)
	lib = ctypes.CDLL(so)
	lib.jo_create.restype = ctypes.c_void_p
	lib.jo_create.argtypes = [ctypes.c_int, _c_double_p, _c_double_p, _c_double_p]
	lib.jo_get_t.restype = ctypes.c_double
	lib.jo_get_dt.restype = ctypes.c_double
	lib.jo_raw_error.restype = ctypes.c_double
	lib.jo_raw_pws.restype = ctypes.c_double
	lib.jo_raw_get_p.restype = ctypes.c_double
	lib.jo_n_anchors.restype = ctypes.c_long
	for fn in ("jo_destroy", "jo_set_parameters", "jo_set_integration_parameters", "jo_get_t", "jo_get_dt", "jo_status",
			"jo_n_anchors", "jo_counters", "jo_get_full_state", "jo_get_current_state", "jo_raw_next_step", "jo_raw_accept",
			"jo_raw_error", "jo_raw_pws", "jo_raw_get_p", "jo_raw_forget", "jo_raw_recent_state", "jo_raw_truncate",
			"jo_integrate", "jo_step_on_discontinuities", "jo_jump", "jo_adjust_diff", "jo_integrate_blindly"):
		getattr(lib, fn).argtypes = None
	lib._program = program
	lib._n_basic = n_basic
	return lib


class Trajectory:
	"""One trajectory in the oracle; method names follow the reference's `jitcdde` class."""

	def __init__(self, lib, past, parameters=(), max_delay=0.0, root_mode=1, **integration_parameters):
		self.lib = lib
		self.n = lib.jo_n()
		self.n_basic = lib.jo_n_basic()
		times = _arr([a[0] for a in past])
		states = _arr([a[1] for a in past], (len(past), self.n))
		diffs = _arr([a[2] for a in past], (len(past), self.n))
		self._h = ctypes.c_void_p(lib.jo_create(len(past), _p(times), _p(states), _p(diffs)))
		if not self._h:
			raise ValueError("need at least two anchors")
		if lib.jo_n_params():
			pars = _arr(parameters)
			assert pars.shape == (lib.jo_n_params(),)
			lib.jo_set_parameters(self._h, _p(pars))
		self.max_delay = float(max_delay)
		self.root_mode = root_mode
		self.set_integration_parameters(**integration_parameters)

	def __del__(self):
		if getattr(self, "_h", None):
			self.lib.jo_destroy(self._h)
			self._h = None

	def set_integration_parameters(self, **kwargs):
		self.params = make_params(self.root_mode, **kwargs)
		self.lib.jo_set_integration_parameters(self._h, ctypes.byref(self.params), ctypes.c_double(self.max_delay))

	@property
	def t(self):
		return self.lib.jo_get_t(self._h)

	@property
	def dt(self):
		return self.lib.jo_get_dt(self._h)

	@property
	def status(self):
		return self.lib.jo_status(self._h)

	@property
	def counters(self):
		out = (ctypes.c_long*3)()
		self.lib.jo_counters(self._h, out)
		return dict(accepted=out[0], rejected=out[1], rhs=out[2])

	def integrate(self, target):
		out = np.empty(self.n)
		status = self.lib.jo_integrate(self._h, ctypes.c_double(target), _p(out))
		if status:
			raise RuntimeError(f"oracle: integration failed with status {status}")
		return out

	def step_on_discontinuities(self, steps):
		steps = _arr(steps)
		out = np.empty(self.n)
		status = self.lib.jo_step_on_discontinuities(self._h, len(steps), _p(steps), _p(out))
		if status:
			raise RuntimeError(f"oracle: integration failed with status {status}")
		return out

	def adjust_diff(self, shift_ratio=1e-4):
		mi, ma = np.empty(self.n), np.empty(self.n)
		self.lib.jo_adjust_diff(self._h, ctypes.c_double(shift_ratio), _p(mi), _p(ma))
		return mi, ma

	def jump(self, amplitude, time, width=1e-5, forward=True):
		amplitude = _arr(np.broadcast_to(np.asarray(amplitude, dtype=float), (self.n,)))
		mi, ma = np.empty(self.n), np.empty(self.n)
		self.lib.jo_jump(self._h, _p(amplitude), ctypes.c_double(time), ctypes.c_double(width), int(forward), _p(mi), _p(ma))
		return mi, ma

	def integrate_blindly(self, target, step=None):
		out = np.empty(self.n)
		status = self.lib.jo_integrate_blindly(self._h, ctypes.c_double(target), ctypes.c_double(step or 0.0), _p(out))
		if status:
			raise RuntimeError(f"oracle: blind integration failed with status {status}")
		return out

	def integrate_lyap(self, target):
		n_lyap = self.n//self.n_basic-1
		state, lyaps, dt = np.empty(self.n_basic), np.empty(n_lyap), ctypes.c_double()
		status = self.lib.jo_integrate_lyap(self._h, ctypes.c_double(target), _p(state), _p(lyaps), ctypes.byref(dt))
		if status:
			raise RuntimeError(f"oracle: integration failed with status {status}")
		return state, lyaps, dt.value

	def orthonormalise(self, n_lyap, delay):
		norms = np.empty(n_lyap)
		self.lib.jo_raw_orthonormalise(self._h, n_lyap, ctypes.c_double(delay), _p(norms))
		return norms

	def get_state(self):
		k = self.lib.jo_n_anchors(self._h)
		times, states, diffs = np.empty(k), np.empty((k, self.n)), np.empty((k, self.n))
		self.lib.jo_get_full_state(self._h, _p(times), _p(states), _p(diffs))
		return [(times[i], states[i].copy(), diffs[i].copy()) for i in range(k)]

	# raw core access (names of the reference's dde_integrator methods, jitced_template.c:1183-1208)
	def get_next_step(self, h):
		self.lib.jo_raw_next_step(self._h, ctypes.c_double(h))

	def accept_step(self):
		self.lib.jo_raw_accept(self._h)

	def error(self, i=0):
		return self.lib.jo_raw_error(self._h, i)

	@property
	def past_within_step(self):
		return self.lib.jo_raw_pws(self._h)

	def get_p(self, atol, rtol):
		return self.lib.jo_raw_get_p(self._h, ctypes.c_double(atol), ctypes.c_double(rtol))

	def forget(self, delay):
		self.lib.jo_raw_forget(self._h, ctypes.c_double(delay))

	def get_recent_state(self, t):
		out = np.empty(self.n)
		self.lib.jo_raw_recent_state(self._h, ctypes.c_double(t), _p(out))
		return out

	def get_current_state(self):
		out = np.empty(self.n)
		self.lib.jo_get_current_state(self._h, _p(out))
		return out

	def truncate(self, time):
		self.lib.jo_raw_truncate(self._h, ctypes.c_double(time))


def run_ensemble(lib, past, parameters, max_delay, disc_mode, disc_steps, sample_offsets, root_mode=1, per_member_past=False, keep_output=True, lyap=False, n_members=None, absolute_samples=False, **integration_parameters):
	"""
	Run N independent members through past → discontinuity handling → sampling loop (jo_run_ensemble).
	`past` = (times[k], states[k][n], diffs[k][n]) shared by all members, or with a leading member axis if
	`per_member_past`. disc_mode: 0 none, 1 step_on_discontinuities(disc_steps[m]), 2 adjust_diff,
	3 integrate_blindly(t + disc_steps[m][0], step = disc_steps[m][1]). Samples are taken at t_after_disc + offset, or at
	the given times themselves with `absolute_samples`. With `lyap`, samples are taken with
	integrate_lyap. Returns dict(out, counters[N][3], status[N], t[N] [, lyaps, weights]).
	"""
	n, n_basic = lib.jo_n(), lib.jo_n_basic()
	times, states, diffs = (_arr(x) for x in past)
	parameters = _arr(parameters) if parameters is not None else np.zeros((0, 0))
	if n_members is not None:
		N = int(n_members)
	elif parameters.ndim == 2 and parameters.shape[0]:
		N = parameters.shape[0]
	elif per_member_past:
		N = times.shape[0]
	else:
		N = int(np.shape(disc_steps)[0])
	k = times.shape[-1]
	P = make_params(root_mode, **integration_parameters)
	disc = _arr(disc_steps if disc_steps is not None else np.zeros((N, 1)))
	if disc.ndim == 1:
		disc = np.ascontiguousarray(np.broadcast_to(disc, (N, disc.shape[0])))
	n_disc = disc.shape[1]
	offs = _arr(sample_offsets)
	n_out = n_basic if lyap else n
	n_lyap = n//n_basic-1
	out = np.empty((len(offs), N, n_out)) if keep_output else None
	lyaps = np.zeros((len(offs), N, n_lyap)) if lyap else None
	weights = np.zeros((len(offs), N)) if lyap else None
	counters = np.zeros((N, 3), dtype=np.int64)
	status = np.zeros(N, dtype=np.int32)
	final_t = np.empty(N)
	lib.jo_run_ensemble(
		ctypes.c_long(N), ctypes.c_int(k), _p(times), _p(states), _p(diffs), ctypes.c_int(int(per_member_past)),
		_p(parameters) if parameters.size else None, ctypes.byref(P), ctypes.c_double(max_delay),
		ctypes.c_int(disc_mode), ctypes.c_int(n_disc), _p(disc),
		ctypes.c_int(-len(offs) if absolute_samples else len(offs)), _p(offs),
		_p(out) if keep_output else None,
		_p(lyaps) if lyap else None, _p(weights) if lyap else None,
		counters.ctypes.data_as(ctypes.POINTER(ctypes.c_long)),
		status.ctypes.data_as(ctypes.POINTER(ctypes.c_int)),
		_p(final_t),
	)
	result = dict(out=out, counters=counters, status=status, t=final_t)
	if lyap:
		result.update(lyaps=lyaps, weights=weights)
	return result

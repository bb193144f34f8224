"""
Host-side mirror of the reference's user-facing classes `jitcdde` and `jitcdde_lyap`
(/root/reference/jitcdde/_jitcdde.py:141-1209): same method names, argument meaning
and error behaviour, so that a script written for the reference runs unchanged —
but every call is ONE whole-ensemble operation of the CUDA engine behind the C ABI
(include/jitcdde_b200.h). There is no Python-side stepping and no CPU fallback.

Ensemble extension (the reference integrates one trajectory per object): pass
`n_members=N`; `set_parameters` then accepts an (N, n_params) array, pasts may
carry a leading member axis, target times may be (N,) arrays, and results get a
leading member axis. With the default `n_members=1` all shapes are the
reference's.
"""

import os
from warnings import warn

import numpy as np
import sympy

from . import _build, _capi, _codegen
from ._past import Anchor, Past, join
from ._symbols import anchors, current_y, input_base_n, input_shift, t

_default_min_step = 1e-10


class UnsuccessfulIntegration(Exception):
	"""
	Raised when the integrator cannot meet the accuracy and step-size requirements
	(reference: _jitcdde.py:135-139). For ensembles, `.status` holds the per-member status.
	"""
	status = None


def _propagate_delays(delays, p, threshold=1e-5):
	"""_jitcdde.py:75-86"""
	result = [0]
	if p != 0:
		for delay in _propagate_delays(delays, p-1, threshold):
			for other_delay in delays:
				new_entry = delay + other_delay
				for old_entry in result:
					if abs(new_entry-old_entry) < threshold:
						break
				else:
					result.append(new_entry)
	return result


def _find_max_delay(delays):
	"""_jitcdde.py:69-73"""
	if all(sympy.sympify(delay).is_Number for delay in delays):
		return float(max(sympy.sympify(delay) for delay in delays))
	raise ValueError("Delay depends on time, dynamics, or control parameters; cannot determine max_delay automatically. You have to pass it as an argument to jitcdde.")


def _is_anchor_helper(helper):
	return getattr(helper[1], "func", None) is not None and type(helper[1]).__name__ == "anchors"


class jitcdde:
	"""
	Parameters follow the reference (_jitcdde.py:141-197): f_sym (iterable, generator
	function, or dict keyed by y(i)), helpers, n, delays, max_delay,
	automatic_anchor_helpers (accepted; lookups with equal delays are always merged),
	control_pars, verbose, rng.

	Not supported (out of the accelerated path): callback_functions (a host
	callback per right-hand-side evaluation cannot run on the device). module_location takes a
	file written by save_compiled (kernel image + dimensions).

	Extensions: n_members (ensemble size), device (CUDA ordinal), ring_capacity (anchors
	kept per member; grows automatically), verification (build kernels without FMA
	contraction: bit-identical with the reference arithmetic).
	"""

	def __init__(
			self,
			f_sym=(), *,
			helpers=None,
			n=None,
			delays=None,
			max_delay=None,
			automatic_anchor_helpers=False,
			control_pars=(),
			callback_functions=(),
			verbose=True,
			module_location=None,
			rng=None,
			n_members=1,
			device=0,
			ring_capacity=None,
			verification=False,
		):
		if callback_functions:
			raise NotImplementedError("callback_functions are not supported: a Python callback per right-hand-side evaluation cannot run on the GPU.")
		if rng is None:
			rng = np.random.default_rng(seed=42)

		self.verbose = verbose
		self._loaded = None
		if module_location is not None:
			# a module file written by save_compiled: kernel image + what the engine needs to know about the system; no
			# symbolic input needed (reference: _jitcdde.py:178-179, 585-592). `n`, if given, must match.
			meta, image = _build.load_image_file(module_location)
			if n is not None and n != meta["n"]:
				raise ValueError(f"n = {n} does not match the module file (n = {meta['n']}).")
			if len(control_pars) != meta["n_params"]:
				raise ValueError(f"The module file was compiled with {meta['n_params']} control parameters; provide them again (control_pars).")
			self._loaded = (meta, image)
			f_sym, n = (), meta["n"]
			if not hasattr(self, "n_basic") or self.n_basic is None:
				self.n_basic = meta["n_basic"]
			verification = meta["mode"] == _build.VERIFY
		self.f_sym = self._handle_input(f_sym, n)
		if not hasattr(self, "n_basic") or self.n_basic is None:
			self.n_basic = self.n
		self.helpers = [(sympy.sympify(h[0]), sympy.sympify(h[1])) for h in (helpers or [])]
		self.control_pars = list(control_pars)
		self.rng = rng
		self.n_members = int(n_members)
		self.device = device
		self.ring_capacity = ring_capacity
		self.verification = verification
		self.automatic_anchor_helpers = automatic_anchor_helpers

		self.initiate_past()
		self.integration_parameters_set = False
		self.DDE = None
		self._engine = None
		self._program = None
		self._image = None
		self._parameters = None
		self.delays = delays
		self.max_delay = max_delay
		self.initial_discontinuities_handled = False
		self.extra_compile_args = ()
		self._int_params_dirty = True
		self._pushed_max_delay = None
		if self._loaded:
			from types import SimpleNamespace
			meta, image = self._loaded
			self._program = SimpleNamespace(n=meta["n"], n_sites=meta["n_sites"], n_params=meta["n_params"], key=meta["key"])
			self._image = image
			if delays is None and max_delay is None:
				raise ValueError("With module_location you must give delays or max_delay (the module file holds no symbolic input).")

	# jitcxde._handle_input: iterable / generator function / dict keyed by y(i)
	def _handle_input(self, f_sym, n=None):
		if isinstance(f_sym, dict):
			entries = [None]*len(f_sym)
			for key, value in f_sym.items():
				key = sympy.sympify(key)
				if type(key).__name__ != "current_y":
					raise ValueError("Dictionary keys must be y(0), y(1), …")
				entries[int(key.args[0])] = value
			if any(e is None for e in entries):
				raise ValueError("Dictionary keys must be y(0), y(1), … without gaps")
		elif callable(f_sym):
			entries = list(f_sym())
		else:
			entries = list(f_sym)
		entries = [sympy.sympify(e) for e in entries]
		if n is not None and entries and len(entries) != n:
			raise ValueError("len(f_sym) and n do not match.")
		self.n = len(entries) if entries else n
		return lambda: list(entries)

	def report(self, message):
		if self.verbose:
			print(message)

	def initiate_past(self):
		self.past = Past(n=self.n, n_basic=self.n_basic, rng=self.rng)

	# ------------------------------------------------------------------ delays (_jitcdde.py:225-249)
	@property
	def delays(self):
		if self._delays is None:
			self.delays = _codegen.get_delays(self.f_sym(), self.helpers)
		return self._delays

	@delays.setter
	def delays(self, new_delays):
		self._delays = new_delays
		if (self._delays is not None) and not self._per_member_delays() and (0 not in list(self._delays)):
			self._delays = list(self._delays)+[0]
		self._max_delay = None

	def _per_member_delays(self):
		return self._delays is not None and np.ndim(self._delays) == 2

	@property
	def max_delay(self):
		if self._max_delay is None:
			if self._per_member_delays():
				self._max_delay = float(np.max(self._delays))
			else:
				self._max_delay = _find_max_delay(self.delays)
			assert self._max_delay >= 0.0, "Negative maximum delay."
		return self._max_delay

	@max_delay.setter
	def max_delay(self, new_max_delay):
		if new_max_delay is not None:
			assert new_max_delay >= 0.0, "Negative maximum delay."
		self._max_delay = new_max_delay

	# ------------------------------------------------------------------ checks (_jitcdde.py:251-280)
	def check(self, fail_fast=True):
		problems = []
		if not self.f_sym():
			problems.append("f_sym is empty.")
		valid_symbols = {t, *[h[0] for h in self.helpers], *[sympy.sympify(p) for p in self.control_pars]}
		for i, entry in enumerate(self.f_sym()):
			entry = _codegen._to_plain_functions(entry)
			for call in entry.atoms(sympy.Function):
				name = type(call).__name__
				if name == "current_y":
					index = call.args[0]
				elif name in ("past_y", "past_dy"):
					index = call.args[1]
				else:
					continue
				if index < 0:
					problems.append(f"y is called with a negative argument ({index}) in equation {i}.")
				if index >= self.n:
					problems.append(f"y is called with an argument ({index}) higher than the system’s dimension ({self.n}) in equation {i}.")
			for symbol in entry.free_symbols:
				if symbol not in valid_symbols:
					problems.append(f"Invalid symbol ({symbol.name}) in equation {i}.")
			if problems and fail_fast:
				break
		if problems:
			raise ValueError("\n".join(problems))

	# ------------------------------------------------------------------ past (_jitcdde.py:282-387)
	def add_past_point(self, time, state, derivative):
		self.reset_integrator(idc_unhandled=True)
		self.past.add((time, state, derivative))

	def add_past_points(self, anchors):
		self.reset_integrator(idc_unhandled=True)
		for anchor in anchors:
			self.past.add(anchor)

	def constant_past(self, state, time=0):
		self.reset_integrator(idc_unhandled=True)
		self.past.constant(state, time)

	def past_from_function(self, function, times_of_interest=None, max_anchors=100, tol=5):
		self.reset_integrator(idc_unhandled=True)
		if times_of_interest is None:
			times_of_interest = np.linspace(-self.max_delay, 0, 10)
		else:
			times_of_interest = sorted(times_of_interest)
		self.past.from_function(function, times_of_interest, max_anchors, tol)

	def get_state(self, member=0):
		"""All anchors currently held by the integrator (after forget(max_delay)) as a `Past`
		that `add_past_points` accepts (_jitcdde.py:357-371). Ensembles: anchors of `member`."""
		self._initiate()
		times, states, diffs = self._engine.get_state(member)
		self.past.clear()
		for k in range(len(times)):
			list.append(self.past, self.past.prepare_anchor((times[k], states[k], diffs[k])))
		return self.past

	def purge_past(self):
		self.past.clear()
		self.reset_integrator(idc_unhandled=True)

	def reset_integrator(self, idc_unhandled=False):
		if idc_unhandled:
			self.initial_discontinuities_handled = False
		self.DDE = None

	# ------------------------------------------------------------------ code generation / compilation
	def generate_lambdas(self, simplify=None):
		raise NotImplementedError("There is no Python/CPU core in this package: the integrator only runs as CUDA kernels on a B200.")

	def save_compiled(self, destination="", overwrite=False):
		"""
		Saves the compiled kernel image together with what the engine needs to know about the system (dimensions, lookup
		sites, control parameters, build mode) to one file that `module_location` accepts, so that a later session — or
		another machine with a B200 — skips code generation and nvcc (reference: jitcxde_common's save_compiled, used at
		tests/test_jitcdde.py:138-147). `destination`: a file name, or a directory (or "": the working directory) in which
		the file is named after the system. Returns the path.
		"""
		if self._program is None:
			self.compile_C()
		name = f"jitced_{self._program.key}.jdde"
		if destination == "" or os.path.isdir(destination):
			destination = os.path.join(destination or ".", name)
		meta = dict(n=self.n, n_basic=self.n_basic, n_sites=self._program.n_sites, n_params=self._program.n_params, key=self._program.key,
			mode=_build.VERIFY if self.verification else _build.FAST, format=1)
		return _build.save_image(destination, self._image, meta, overwrite=overwrite)

	def compile_C(self, simplify=None, do_cse=True, chunk_size=100, extra_compile_args=None, extra_link_args=None, verbose=False, modulename=None, omp=False):
		"""
		Translate the derivative to CUDA and compile it for sm_100a (replaces the C translation
		of _jitcdde.py:420-575). `chunk_size`, `extra_link_args`, `modulename` and `omp` have no
		meaning here and are ignored; `extra_compile_args` are passed to nvcc.
		"""
		if self._loaded:
			raise RuntimeError("This integrator was loaded from a module file (module_location); there is nothing to compile.")
		# The reference simplifies by default for n <= 10 (SymEngine's cheap simplify). SymPy's
		# simplify is slow and only reorders arithmetic, so it is opt-in here.
		simplify = bool(simplify)
		self.extra_compile_args = tuple(extra_compile_args or ())
		blocks = (self.n_basic, self.n//self.n_basic-1) if self.n_basic != self.n else None
		self._program = _codegen.generate(self.f_sym(), helpers=self.helpers, control_pars=self.control_pars, simplify=simplify, do_cse=do_cse, blocks=blocks)
		if not self._program.n_sites:
			warn("Differential equation does not include a delay term.", stacklevel=2)
		mode = _build.VERIFY if self.verification else _build.FAST
		self._image = _build.build_image(self._program, n_basic=self.n_basic, mode=mode, extra_flags=self.extra_compile_args, verbose=verbose)
		self.reset_integrator()
		self._drop_engine()

	def _drop_engine(self):
		if self._engine is not None:
			self._engine.close()
			self._engine = None

	def _default_ring_capacity(self):
		needed = max(64, 2*len(self.past)+8)
		capacity = 64
		while capacity < needed:
			capacity *= 2
		return capacity

	def _initiate(self):
		if self._program is None:
			self.compile_C()
		if self.DDE is None:
			assert len(self.past) > 1, "You need to add at least two past points first. Usually this means that you did not set an initial past at all."
			capacity = self.ring_capacity or self._default_ring_capacity()
			while capacity < len(self.past)+2:
				capacity *= 2
			if self._engine is None or self._engine.ring_capacity < capacity:
				self._drop_engine()
				self._engine = _capi.Engine(self.n, self.n_basic, self._program.n_sites, self._program.n_params, self.n_members, ring_capacity=capacity, device=self.device)
				self._engine.load_image(self._image)
				self._int_params_dirty = True
			self._engine.set_past(*self.past.as_arrays(self.n_members))
			if self._parameters is not None:
				self._engine.set_parameters(self._parameters)
			self.DDE = self._engine
		self._set_integration_parameters()
		if self._int_params_dirty:
			# Step size and past-within-step memory live per member on the device. As in the reference
			# (attributes of the Python object, _jitcdde.py:710-728) they are reset by
			# set_integration_parameters only and survive a new past.
			self._engine.set_integration_parameters(self.max_delay, **self._int_params)
			self._pushed_max_delay = self.max_delay
			self._int_params_dirty = False
		elif self._pushed_max_delay != self.max_delay:
			self._engine.set_max_delay(self.max_delay)
			self._pushed_max_delay = self.max_delay

	# ------------------------------------------------------------------ parameters (_jitcdde.py:594-743)
	def set_parameters(self, *parameters):
		"""Control parameters in the order of `control_pars`: several numbers, one sequence, or
		(ensembles) one array of shape (n_members, n_params)."""
		if len(parameters) == 1 and np.ndim(parameters[0]) >= 1:
			values = np.array(parameters[0], dtype=float)
		else:
			values = np.array(parameters, dtype=float)
		expected = [(len(self.control_pars),), (self.n_members, len(self.control_pars))]
		if values.shape not in expected:
			raise TypeError("Argument must either be a single sequence or multiple numbers.")
		self._parameters = values
		self._initiate()
		self.initial_discontinuities_handled = False
		self._engine.set_parameters(values)

	def _set_integration_parameters(self):
		if not self.integration_parameters_set:
			self.report("Using default integration parameters.")
			self.set_integration_parameters()

	def set_integration_parameters(self,
			atol=1e-10,
			rtol=1e-5,
			first_step=1.0,
			min_step=_default_min_step,
			max_step=10.0,
			decrease_threshold=1.1,
			increase_threshold=0.5,
			safety_factor=0.9,
			max_factor=5.0,
			min_factor=0.2,
			pws_factor=3,
			pws_atol=0.0,
			pws_rtol=1e-5,
			pws_max_iterations=10,
			pws_base_increase_chance=0.1,
			pws_fuzzy_increase=False,
			):
		if first_step > max_step:
			first_step = max_step
			warn("Decreasing first_step to match max_step", stacklevel=2)
		if min_step > first_step:
			min_step = first_step
			warn("Decreasing min_step to match first_step", stacklevel=2)

		assert decrease_threshold >= 1.0, "decrease_threshold smaller than 1"
		assert increase_threshold <= 1.0, "increase_threshold larger than 1"
		assert max_factor >= 1.0, "max_factor smaller than 1"
		assert min_factor <= 1.0, "min_factor larger than 1"
		assert safety_factor <= 1.0, "safety_factor larger than 1"
		assert atol >= 0.0, "negative atol"
		assert rtol >= 0.0, "negative rtol"
		if atol == 0 and rtol == 0:
			warn("atol and rtol are both 0. You probably do not want this.", stacklevel=2)
		assert pws_atol >= 0.0, "negative pws_atol"
		assert pws_rtol >= 0.0, "negative pws_rtol"
		assert 0 < pws_max_iterations, "non-positive pws_max_iterations"
		assert 2 <= pws_factor, "pws_factor smaller than 2"
		assert pws_base_increase_chance >= 0, "negative pws_base_increase_chance"

		self._int_params = dict(
			atol=atol, rtol=rtol, first_step=first_step, min_step=min_step, max_step=max_step,
			decrease_threshold=decrease_threshold, increase_threshold=increase_threshold,
			safety_factor=safety_factor, max_factor=max_factor, min_factor=min_factor,
			pws_factor=pws_factor, pws_atol=pws_atol, pws_rtol=pws_rtol,
			pws_max_iterations=pws_max_iterations, pws_base_increase_chance=pws_base_increase_chance,
			# pws_fuzzy_increase (_jitcdde.py:730-731): the draws come from a counter-based generator on the device, seeded from self.rng
			pws_fuzzy_seed=int(self.rng.integers(1, 2**31-1)) if pws_fuzzy_increase else 0,
		)
		self.atol, self.rtol, self.min_step, self.max_step = atol, rtol, min_step, max_step
		self.integration_parameters_set = True
		self._int_params_dirty = True

	@property
	def dt(self):
		"""current step size (per member for ensembles)"""
		self._initiate()
		return self._shape_scalar(self._engine.get_dt())

	# ------------------------------------------------------------------ helpers for shapes and failures
	def _shape_scalar(self, values):
		return float(values[0]) if self.n_members == 1 else values

	def _shape_state(self, states, n_out=None):
		return states[0].copy() if self.n_members == 1 else states

	def _check_status(self, status, retry=None):
		if status is None or not status.any():
			return False
		if (status == _capi.STATUS_OVERFLOW).any() and retry is not None:
			new_capacity = self._engine.ring_capacity*2
			self.report(f"Anchor ring full; growing to {new_capacity} anchors per member.")
			self._engine.grow_ring(new_capacity)
			return True
		if (status == _capi.STATUS_MIN_STEP).any():
			message = ["",
				"Could not integrate with the given tolerance parameters:\n",
				f"atol: {self.atol:e}",
				f"rtol: {self.rtol:e}",
				f"min_step: {self.min_step:e}\n",
				"The most likely reasons for this are:",
			]
			if not self.initial_discontinuities_handled:
				message.append("• You did not sufficiently address initial discontinuities. See https://jitcdde.rtfd.io/#discontinuities for details.")
			message.append("• The DDE is ill-posed or stiff.")
			if self.initial_discontinuities_handled:
				message.append("• You used `integrate_blindly` and did not adjust the maximum step or you used `adjust_diff` and did not pay attention to the extrema.")
			if self.atol == 0:
				message.append("• You did not allow for an absolute error tolerance (atol) though your DDE calls for it. Even a very small absolute tolerance (1e-16) may sometimes help.")
			if self.n_members > 1:
				message.append(f"({int((status == _capi.STATUS_MIN_STEP).sum())} of {self.n_members} members failed; see .status of this exception.)")
			error = UnsuccessfulIntegration("\n".join(message))
			error.status = status
			raise error
		if (status == _capi.STATUS_BACKWARDS).any():
			raise ValueError("Can’t integrate backwards in time.")
		if (status == _capi.STATUS_OVERFLOW).any():
			raise MemoryError("Anchor ring full; pass a larger ring_capacity to jitcdde.")
		if (status == _capi.STATUS_BUDGET).any():
			raise UnsuccessfulIntegration("A member needed more attempted steps in one call than the kernel watchdog allows (Engine.set_attempt_budget).")
		error = UnsuccessfulIntegration("The state of at least one member became non-finite or its step size collapsed to zero.")
		error.status = status
		raise error

	# ------------------------------------------------------------------ integration (_jitcdde.py:796-1044)
	@property
	def t(self):
		self._initiate()
		return self._shape_scalar(self._engine.get_t())

	def integrate(self, target_time, out=None, wait=True):
		"""
		As the reference's `integrate` (_jitcdde.py:807-836). Ensemble extension: `target_time`
		may be an array of shape (n_members,), and `out` a preallocated (n_members, n) float64
		array (e.g. from `jitcdde_b200.pinned_empty`) that receives and is returned as the result.

		Pipelined sampling: with `wait=False` (needs `out`, page-locked) the call only queues the work and
		returns `out` at once; its content is valid after `synchronize()`. The copy of one sample to the host
		then overlaps the integration up to the next one. Alternate between (at least) two `out` buffers.
		Failures of members are raised by `synchronize()`; a full ring cannot be resumed in this mode (choose
		`ring_capacity` large enough or sample synchronously).
		"""
		self._initiate()
		if self.n_members == 1 and wait and self._current_t() > target_time:
			warn("The target time is smaller than the current time. No integration step will happen. The returned state will be extrapolated from the interpolating Hermite polynomial for the last integration step. You may see this because you try to integrate backwards in time, in which case you did something wrong. You may see this just because your sampling step is small, in which case there is no need to worry (though you should think about increasing your sampling time).", stacklevel=2)
		if not self.initial_discontinuities_handled:
			warn("You did not explicitly handle initial discontinuities. Proceed only if you know what you are doing. This is only fine if you somehow chose your initial past such that the derivative of the last anchor complies with the DDE. In this case, you can set the attribute `initial_discontinuities_handled` to `True` to suppress this warning. See https://jitcdde.rtfd.io/#discontinuities for details.", stacklevel=2)
		if not wait:
			if out is None:
				raise ValueError("wait=False needs a preallocated `out` (jitcdde_b200.pinned_empty)")
			self._queued = True
			return self._engine.integrate_async(target_time, out)
		if getattr(self, "_queued", False):
			self.synchronize()
		if self.n_members == 1:
			# one trajectory (the reference's use): state, status and the time reached come back in one round trip; the time
			# serves the check above at the next call (the reference asks get_t before every call, _jitcdde.py:824)
			result, status, t_after = self._engine.integrate_t(target_time, out=out)
			self._t_known = (self._engine, float(t_after[0])) if not status.any() else None
			self._t_known_launches = self._engine.launch_count      # any later launch (jump, new past, …) voids the cached time
			status = status if status.any() else None
		else:
			# the per-member status stays on the device unless a member failed (8 bytes back instead of 4·N)
			result, status = self._engine.integrate(target_time, out=out, fetch_status=False)
			status = self._engine.get_status() if self._engine.count_failed() else None
		while self._check_status(status, retry=True):
			result, status = self._engine.resume()
			if out is not None:
				out[...] = result
				result = out
		return self._shape_state(result)

	def integrate_samples(self, times, out=None, wait=True):
		"""
		Extension: the sampling loop `[DDE.integrate(T) for T in times]` (examples/mackey_glass.py:76-78) as ONE kernel
		launch — the same steps and the same states as the separate calls, bit for bit, but a trajectory does not wait
		for the others at every sample and the host is not in the loop. `times`: ascending, shape (n_samples,) or
		(n_members, n_samples). Returns an array of shape (n_samples, n) — (n_samples, n_members, n) for ensembles —,
		`out` if given (page-locked memory from `pinned_empty` is written by the kernel itself). `wait=False` queues the
		work like `integrate(..., wait=False)`.
		"""
		self._initiate()
		times = np.asarray(times, dtype=float)
		if times.ndim not in (1, 2) or times.shape[-1] < 1 or np.any(np.diff(times, axis=-1) < 0):
			raise ValueError("times must be a non-empty ascending sequence")
		if not self.initial_discontinuities_handled:
			warn("You did not explicitly handle initial discontinuities. Proceed only if you know what you are doing. See https://jitcdde.rtfd.io/#discontinuities for details.", stacklevel=2)
		if not wait:
			if out is None:
				raise ValueError("wait=False needs a preallocated `out` (jitcdde_b200.pinned_empty)")
			self._queued = True
			return self._engine.integrate_samples_async(times, out)
		if getattr(self, "_queued", False):
			self.synchronize()
		result, status = self._engine.integrate_samples(times, out=out, fetch_status=False)
		status = self._engine.get_status() if self._engine.count_failed() else None
		while self._check_status(status, retry=True):
			result, status = self._engine.resume(samples=result)
		return result[:, 0].copy() if self.n_members == 1 and out is None else result

	def _current_t(self):
		"""time of a single trajectory: as reported by the last integrate() if nothing else has touched the engine since"""
		known = getattr(self, "_t_known", None)
		if known is not None and known[0] is self._engine and self._engine.launch_count == getattr(self, "_t_known_launches", -1):
			return known[1]
		return float(self._engine.get_t()[0])

	def wait_result(self, age=0):
		"""wait for the `out` of the `integrate(..., wait=False)` call `age` calls ago (0: the latest, 1: the one before)
		without waiting for younger work; member failures are only reported by `synchronize()`"""
		self._engine.wait_result(age)

	def synchronize(self):
		"""wait for the results of `integrate(..., wait=False)` calls; raises if a member failed meanwhile"""
		self._initiate()
		self._engine.synchronize()
		self._queued = False
		if self._engine.count_failed():
			status = self._engine.get_status()
			if np.any(status == _capi.STATUS_OVERFLOW):
				raise RuntimeError("the anchor ring of a member ran full during pipelined sampling: pass a larger ring_capacity or sample with wait=True")
			self._check_status(status)

	def adjust_diff(self, shift_ratio=1e-4):
		self._initiate()
		self.initial_discontinuities_handled = True
		self._engine.ensure_capacity(self._engine.max_live_anchors()+3)      # two new anchors; not resumable after an overflow
		minima, maxima = self._engine.adjust_diff(shift_ratio)
		self._check_status(self._engine.get_status())
		return self._shape_state(minima), self._shape_state(maxima)

	def integrate_blindly(self, target_time, step=None):
		self._initiate()
		self.initial_discontinuities_handled = True
		step = step or self.max_step
		assert step > 0, "step must be positive"
		self._room_for_blind_steps(target_time, step)
		result, status = self._engine.integrate_blindly(target_time, step)
		self._check_status(status)
		return self._shape_state(result)

	def _room_for_blind_steps(self, target_time, step):
		"""A blind call cannot be resumed after a ring overflow (the reference's anchor list is unbounded): make the ring large
		enough beforehand — what the members hold now plus the steps of the call that forget() cannot drop, i.e. those within
		max_delay of the end, plus a margin."""
		total = float(np.max(np.asarray(target_time, dtype=float)-self._engine.get_t()))
		if not total > 0:
			return self._engine.ensure_capacity(self._engine.max_live_anchors()+3)
		number = max(1, round(total/min(step, total)))
		kept = int(np.ceil(min(total, self.max_delay)/(total/number)))+2 if np.isfinite(self.max_delay) else number
		self._engine.ensure_capacity(self._engine.max_live_anchors()+min(number, kept)+4)

	def step_on_discontinuities(self, *, propagations=1, min_distance=1e-5, max_step=None):
		self._initiate()
		self.initial_discontinuities_handled = True

		assert min_distance > 0, "min_distance must be positive."
		assert isinstance(propagations, int), "Non-integer number of propagations."
		if max_step is not None:
			warn("The max_step parameter is retired and does not need to be used anymore. Instead, step_on_discontinuities now adapts the step size. This will raise an error in the future.", stacklevel=2)

		if self._per_member_delays():
			# ensemble extension: one row of numeric delays per member (e.g. a delay that is a control parameter)
			rows = []
			for row in np.asarray(self.delays, dtype=float):
				row = list(row)
				if 0 not in row:
					row.append(0.0)
				steps = sorted(_propagate_delays(row, propagations, min_distance))
				steps.remove(0)
				rows.append(steps)
			if len({len(r) for r in rows}) != 1:
				raise ValueError("Members have different numbers of discontinuity steps.")
			steps = np.array(rows, dtype=float)
			has_steps = steps.shape[1] > 0
		else:
			if not all(sympy.sympify(delay).is_number for delay in self.delays):
				raise ValueError("At least one delay depends on time, dynamics, or control parameters; cannot automatically determine steps.")
			self.delays = [float(sympy.sympify(delay)) for delay in self.delays]
			steps = _propagate_delays(self.delays, propagations, min_distance)
			steps.sort()
			steps.remove(0)
			has_steps = bool(steps)
			steps = np.array(steps, dtype=float)

		if has_steps:
			result, status = self._engine.step_on_discontinuities(steps)
			while self._check_status(status, retry=True):
				result, status = self._engine.resume()
			return self._shape_state(result)
		else:
			self.adjust_diff()
			state = self._engine.get_current_state()[:, :self.n_basic]
			return self._shape_state(state)

	def jump(self, amplitude, time, width=1e-5, forward=True):
		self._initiate()
		self.initial_discontinuities_handled = True
		assert width >= 0
		if np.ndim(amplitude) == 0:
			amplitude = np.full(self.n, amplitude, dtype=float)
		amplitude = np.array(amplitude, dtype=float)
		assert amplitude.shape in [(self.n,), (self.n_members, self.n)]
		self._engine.ensure_capacity(self._engine.max_live_anchors()+3)      # two new anchors; not resumable after an overflow
		minima, maxima = self._engine.jump(amplitude, time, width, forward)
		self._check_status(self._engine.get_status())
		return self._shape_state(minima), self._shape_state(maxima)

	# ------------------------------------------------------------------ extras
	@property
	def status(self):
		self._initiate()
		return self._engine.get_status()

	def counters(self):
		"""(accepted steps, rejected attempts, right-hand-side evaluations) per member"""
		self._initiate()
		return self._engine.get_counters()


class jitcdde_lyap(jitcdde):
	"""
	Lyapunov exponents of the dynamics (reference: _jitcdde.py:1091-1209): the system is
	extended by `n_lyap` tangent vectors (symbolic Jacobians per delay), and after every
	`integrate` the separation functions are orthonormalised over the delay window on the GPU.
	"""

	def __init__(self, f_sym=(), n_lyap=1, simplify=None, **kwargs):
		if kwargs.get("module_location") is not None:
			# a saved module holds the extended system already (save_compiled of a jitcdde_lyap)
			meta, _ = _build.load_image_file(kwargs["module_location"])
			n_basic = kwargs.pop("n", None) or meta["n_basic"]
			if n_basic != meta["n_basic"] or meta["n"] != n_basic*(n_lyap+1):
				raise ValueError(f"The module file holds a system with n_basic = {meta['n_basic']} and {meta['n']//meta['n_basic']-1} Lyapunov exponents.")
			self.n_basic = n_basic
			self._n_lyap = n_lyap
			super().__init__((), n=meta["n"], **kwargs)
			self.n_basic = n_basic
			self.initiate_past()
			assert self.max_delay > 0, "Maximum delay must be positive for calculating Lyapunov exponents."
			return
		self.n_basic = kwargs.pop("n", None)
		helpers = [(sympy.sympify(h[0]), sympy.sympify(h[1])) for h in (kwargs.pop("helpers", None) or [])]
		if any(_is_anchor_helper(helper) for helper in helpers):
			raise NotImplementedError("Cannot use anchor helpers with Lyapunov exponents.")
		kwargs.setdefault("automatic_anchor_helpers", True)

		basic = jitcdde._handle_input(self, f_sym, self.n_basic)()
		self.n_basic = len(basic)
		if simplify is None:
			simplify = self.n_basic <= 10

		if not kwargs.get("delays"):
			kwargs["delays"] = _codegen.get_delays(basic, helpers)
		assert n_lyap >= 0, "n_lyap negative"
		self._n_lyap = n_lyap

		f_lyap = _codegen.tangent_vector_f(basic, helpers, self.n_basic, n_lyap, kwargs["delays"], simplify=simplify)
		n_basic = self.n_basic
		super().__init__(f_lyap, n=n_basic*(n_lyap+1), helpers=(), **kwargs)
		self.n_basic = n_basic
		self.initiate_past()
		assert self.max_delay > 0, "Maximum delay must be positive for calculating Lyapunov exponents."

	def integrate(self, target_time):
		"""Returns (state[:n_basic], local Lyapunov exponents, integration time), _jitcdde.py:1141-1172."""
		self._initiate()
		if not self.initial_discontinuities_handled:
			warn("You did not explicitly handle initial discontinuities. See https://jitcdde.rtfd.io/#discontinuities for details.", stacklevel=2)
		state, lyaps, delta_t, status = self._engine.integrate_lyap(target_time)
		while self._check_status(status, retry=True):
			state, lyaps, delta_t, status = self._engine.resume(lyap=True)
		if self.n_members == 1:
			return state[0].copy(), lyaps[0].copy(), float(delta_t[0])
		return state, lyaps, delta_t

	def set_integration_parameters(self, **kwargs):
		if self._n_lyap/self.n_basic > 2:
			required_max_step = self.max_delay/(np.ceil(self._n_lyap/self.n_basic/2)-1)
			if "max_step" in kwargs.keys():
				if kwargs["max_step"] > required_max_step:
					kwargs["max_step"] = required_max_step
					warn(f"Decreased max_step to {required_max_step} to ensure sufficient dimensionality for Lyapunov exponents.", stacklevel=2)
			else:
				kwargs["max_step"] = required_max_step
			kwargs.setdefault("min_step", _default_min_step)
			if kwargs["min_step"] > required_max_step:
				warn("Given the number of desired Lyapunov exponents and the maximum delay in the system, the highest possible step size is lower than the default min_step or the min_step set by you. This is almost certainly a very bad thing. Nonetheless I will lower min_step accordingly.", stacklevel=2)
		super().set_integration_parameters(**kwargs)

	def integrate_blindly(self, target_time, step=None):
		"""Like jitcdde.integrate_blindly, orthonormalising after each step (_jitcdde.py:1191-1209)."""
		self._initiate()
		self.initial_discontinuities_handled = True
		step = step or self.max_step
		assert step > 0, "step must be positive"
		self._room_for_blind_steps(target_time, step)
		state, lyaps, total, status = self._engine.integrate_blindly_lyap(target_time, step)
		self._check_status(status)
		if self.n_members == 1:
			return state[0].copy(), lyaps[0].copy(), float(total[0])
		return state, lyaps, total


class jitcdde_input(jitcdde):
	"""
	Integrate the DDE (or ODE) with an external input (reference: _jitcdde.py:1519-1659). The input — a sequence of
	anchors (time, state, diff) describing a cubic Hermite spline, e.g. what `get_state()` returns — is stored in the
	past of dummy dynamical variables, which on the GPU are ordinary components of the anchors: `input(i)` in the
	equations becomes a delayed lookup of component n+i. The integration must start at t = 0.
	"""

	def __init__(self, f_sym=(), input=None, **kwargs):  # noqa: A002
		if input is None:
			raise ValueError("You must define an input; otherwise just use plain jitcdde.")
		input = [Anchor(float(a[0]), np.array(a[1], dtype=float), np.array(a[2], dtype=float)) for a in input]  # noqa: A001
		if input[0].time > 0:
			warn(f"Your input past does not begin at t=0 but at t={input[0].time}. Values before the beginning of the past will be extrapolated. You very likely do not want this.", stacklevel=2)
		n_input = len(input[0].state)

		basic = jitcdde._handle_input(self, f_sym, kwargs.pop("n", None))()
		self.input_base_n = len(basic)
		self.input_duration = input[-1].time
		substitutions = {input_base_n: self.input_base_n, input_shift: self.input_duration}

		self.regular_past = Past(n=self.input_base_n)
		self.input_past = Past(n=n_input)
		for anchor in input:
			self.input_past.add((anchor.time-self.input_duration, anchor.state, anchor.diff))

		# dummy dynamical variables with the constant derivative of the last input point (:1561-1562)
		f_full = [sympy.sympify(e).subs(substitutions) for e in basic] + [sympy.Float(d) for d in input[-1].diff]

		if kwargs.setdefault("delays", None) is not None:
			kwargs["delays"] = [*kwargs["delays"], self.input_duration]
		self._joined = None
		super().__init__(f_full, n=len(f_full), **kwargs)

	@property
	def max_delay(self):
		return jitcdde.max_delay.fget(self)

	@max_delay.setter
	def max_delay(self, new_max_delay):
		if new_max_delay is not None:
			assert new_max_delay >= 0.0, "Negative maximum delay."
			new_max_delay = max(new_max_delay, self.input_duration)
		self._max_delay = new_max_delay

	def initiate_past(self):
		pass

	@property
	def past(self):
		if self._joined is None:
			assert len(self.regular_past) > 1, "You need to add at least two past points first. Usually this means that you did not set an initial past at all."
			if self.regular_past[-1].time != 0:
				raise ValueError("For jitcdde_input, the initial past must end at t=0.")
			self._joined = join(self.regular_past, self.input_past)
		return self._joined

	@past.setter
	def past(self, value):
		raise AssertionError("For jitcdde_input, the past attribute should not be directly set.")

	def add_past_point(self, time, state, derivative):
		self._joined = None
		self.reset_integrator(idc_unhandled=True)
		self.regular_past.add((time, state, derivative))

	def add_past_points(self, anchors):
		self._joined = None
		self.reset_integrator(idc_unhandled=True)
		for anchor in anchors:
			self.regular_past.add(anchor)

	def constant_past(self, state, time=0):
		self._joined = None
		self.reset_integrator(idc_unhandled=True)
		self.regular_past.constant(state, time)

	def past_from_function(self, function, times_of_interest=None, max_anchors=100, tol=5):
		self._joined = None
		self.reset_integrator(idc_unhandled=True)
		if times_of_interest is None:
			times_of_interest = np.linspace(-self.max_delay, 0, 10)
		else:
			times_of_interest = sorted(times_of_interest)
		self.regular_past.from_function(function, times_of_interest, max_anchors, tol)

	def get_state(self, member=0):
		self._initiate()
		times, states, diffs = self._engine.get_state(member)
		self._joined = None
		self.regular_past.clear()
		for k in range(len(times)):
			list.append(self.regular_past, Anchor(times[k], states[k][:self.input_base_n].copy(), diffs[k][:self.input_base_n].copy()))
		return self.regular_past

	def purge_past(self):
		self._joined = None
		self.regular_past.clear()
		self.reset_integrator(idc_unhandled=True)

	def integrate(self, target_time, out=None):
		"""`out`, if given, receives the state of the system proper (without the dummy variables of the input) and is returned"""
		if np.any(np.asarray(target_time) > self.input_duration):
			warn("Integrating beyond duration of input. From now on, the input will be extrapolated. You very likely do not want this. Instead define a longer input or stop integrating.", stacklevel=2)
		result = super().integrate(target_time)[..., :self.input_base_n]
		if out is not None:
			if np.shape(out) != np.shape(result):
				raise ValueError(f"out must have shape {np.shape(result)}")
			out[...] = result
			return out
		return result

	def integrate_blindly(self, *args, **kwargs):
		return super().integrate_blindly(*args, **kwargs)[..., :self.input_base_n]

	def step_on_discontinuities(self, *args, **kwargs):
		raise NotImplementedError("Stepping on discontinuities is not implemented for input yet. Use integrate_blindly or adjust_diff instead.")

	def adjust_diff(self, shift_ratio=1e-4):
		minima, maxima = super().adjust_diff(shift_ratio)
		return minima[..., :self.input_base_n], maxima[..., :self.input_base_n]

	def jump(self, amplitude, *args, **kwargs):
		if np.ndim(amplitude) == 0:
			amplitude = np.full(self.input_base_n, amplitude, dtype=float)
		amplitude = np.asarray(amplitude, dtype=float)
		assert amplitude.shape == (self.input_base_n,)
		extended_amplitude = np.hstack((amplitude, np.zeros(self.input_past.n)))
		minima, maxima = super().jump(extended_amplitude, *args, **kwargs)
		return minima[..., :self.input_base_n], maxima[..., :self.input_base_n]


def test(omp=True, sympy=True):
	"""
	Quick sanity check of the installation (reference: _jitcdde.py:1661-1677): code generation, nvcc, the runtime
	library and a CUDA device. Finishes without any message if everything works. `omp` has no meaning here.
	"""
	if sympy:
		import sympy  # noqa: F401
	from ._symbols import t, y
	DDE = jitcdde([y(1, t-1), -y(0, t-2)], verbose=False)
	DDE.compile_C(chunk_size=1, omp=omp)
	DDE.constant_past([1, 2])
	DDE.step_on_discontinuities()
	DDE.integrate(DDE.t+1)


test.__test__ = False      # not a pytest test

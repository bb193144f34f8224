"""
ctypes binding of the C ABI (include/jitcdde_b200.h) — the thin layer between the
Python drop-in classes (`_jitcdde.py`) and libjitcdde_b200.so.

In the reference this place is taken by the generated CPython extension type
`dde_integrator` (/root/reference/jitcdde/jitced_template.c:1183-1232).
There is no fallback: if the library is missing or no CUDA device is present,
an exception is raised.
"""

import ctypes
import os

import numpy as np

from . import _build

c_double_p = ctypes.POINTER(ctypes.c_double)
c_int32_p = ctypes.POINTER(ctypes.c_int32)
c_int64_p = ctypes.POINTER(ctypes.c_int64)

PROJECT_RESTRICTED, PROJECT_TRANSVERSAL = 1, 2
STATUS_OK, STATUS_MIN_STEP, STATUS_OVERFLOW, STATUS_NONFINITE, STATUS_BACKWARDS, STATUS_BUDGET = range(6)


class Desc(ctypes.Structure):
	_fields_ = [
		("n", ctypes.c_int32), ("n_basic", ctypes.c_int32), ("n_sites", ctypes.c_int32), ("n_params", ctypes.c_int32),
		("ring_capacity", ctypes.c_int32), ("device", ctypes.c_int32), ("n_members", ctypes.c_int64),
	]


class IntParams(ctypes.Structure):
	_fields_ = [
		("atol", ctypes.c_double), ("rtol", ctypes.c_double), ("first_step", ctypes.c_double),
		("min_step", ctypes.c_double), ("max_step", ctypes.c_double),
		("decrease_threshold", ctypes.c_double), ("increase_threshold", ctypes.c_double),
		("safety_factor", ctypes.c_double), ("max_factor", ctypes.c_double), ("min_factor", ctypes.c_double),
		("pws_factor", ctypes.c_double), ("pws_atol", ctypes.c_double), ("pws_rtol", ctypes.c_double),
		("pws_base_increase_chance", ctypes.c_double),
		("pws_max_iterations", ctypes.c_int32), ("pws_fuzzy_seed", ctypes.c_int32),
	]


#: every symbol include/jitcdde_b200.h declares, with (restype, argtypes)
_H = ctypes.c_void_p
SIGNATURES = {
	"jdde_create": (ctypes.c_int, [ctypes.POINTER(Desc), ctypes.POINTER(_H)]),
	"jdde_destroy": (ctypes.c_int, [_H]),
	"jdde_last_error": (ctypes.c_char_p, [_H]),
	"jdde_load_rhs": (ctypes.c_int, [_H, ctypes.c_void_p, ctypes.c_size_t]),
	"jdde_set_stream": (ctypes.c_int, [_H, ctypes.c_void_p]),
	"jdde_synchronize": (ctypes.c_int, [_H]),
	"jdde_set_past": (ctypes.c_int, [_H, ctypes.c_int32, c_double_p, c_double_p, c_double_p, ctypes.c_int32]),
	"jdde_set_parameters": (ctypes.c_int, [_H, c_double_p, ctypes.c_int32]),
	"jdde_set_integration_parameters": (ctypes.c_int, [_H, ctypes.POINTER(IntParams), ctypes.c_double]),
	"jdde_integrate": (ctypes.c_int, [_H, c_double_p, ctypes.c_int32, c_double_p, c_int32_p]),
	"jdde_integrate_t": (ctypes.c_int, [_H, c_double_p, ctypes.c_int32, c_double_p, c_int32_p, c_double_p]),
	"jdde_integrate_async": (ctypes.c_int, [_H, c_double_p, ctypes.c_int32, c_double_p]),
	"jdde_wait_result": (ctypes.c_int, [_H, ctypes.c_int32]),
	"jdde_integrate_samples": (ctypes.c_int, [_H, c_double_p, ctypes.c_int32, ctypes.c_int32, c_double_p, c_int32_p]),
	"jdde_integrate_samples_async": (ctypes.c_int, [_H, c_double_p, ctypes.c_int32, ctypes.c_int32, c_double_p]),
	"jdde_step_on_discontinuities": (ctypes.c_int, [_H, ctypes.c_int32, c_double_p, ctypes.c_int32, c_double_p, c_int32_p]),
	"jdde_resume": (ctypes.c_int, [_H, c_double_p, c_double_p, c_double_p, c_int32_p]),
	"jdde_integrate_blindly": (ctypes.c_int, [_H, c_double_p, ctypes.c_int32, ctypes.c_double, c_double_p, c_int32_p]),
	"jdde_integrate_blindly_lyap": (ctypes.c_int, [_H, c_double_p, ctypes.c_int32, ctypes.c_double, c_double_p, c_double_p, c_double_p, c_int32_p]),
	"jdde_set_projector": (ctypes.c_int, [_H, ctypes.c_int32, c_double_p, ctypes.c_int32, c_int32_p, ctypes.c_int32, c_int32_p, ctypes.c_int32, c_int32_p, ctypes.c_int32]),
	"jdde_project": (ctypes.c_int, [_H, ctypes.c_double, c_double_p]),
	"jdde_integrate_blindly_projected": (ctypes.c_int, [_H, c_double_p, ctypes.c_int32, ctypes.c_double, c_double_p, c_double_p, c_double_p, c_int32_p]),
	"jdde_set_max_delay": (ctypes.c_int, [_H, ctypes.c_double]),
	"jdde_jump": (ctypes.c_int, [_H, c_double_p, ctypes.c_int32, c_double_p, ctypes.c_int32, ctypes.c_double, ctypes.c_int32, c_double_p, c_double_p]),
	"jdde_adjust_diff": (ctypes.c_int, [_H, ctypes.c_double, c_double_p, c_double_p]),
	"jdde_orthonormalise": (ctypes.c_int, [_H, ctypes.c_int32, ctypes.c_double, c_double_p]),
	"jdde_integrate_lyap": (ctypes.c_int, [_H, c_double_p, ctypes.c_int32, c_double_p, c_double_p, c_double_p, c_int32_p]),
	"jdde_get_t": (ctypes.c_int, [_H, c_double_p]),
	"jdde_get_dt": (ctypes.c_int, [_H, c_double_p]),
	"jdde_get_current_state": (ctypes.c_int, [_H, c_double_p]),
	"jdde_get_status": (ctypes.c_int, [_H, c_int32_p]),
	"jdde_count_failed": (ctypes.c_int, [_H, c_int64_p]),
	"jdde_get_counters": (ctypes.c_int, [_H, c_int64_p, c_int64_p, c_int64_p]),
	"jdde_sum_counters": (ctypes.c_int, [_H, c_int64_p, c_int64_p, c_int64_p]),
	"jdde_get_state": (ctypes.c_int, [_H, ctypes.c_int64, ctypes.c_int32, c_int32_p, c_double_p, c_double_p, c_double_p]),
	"jdde_max_live_anchors": (ctypes.c_int, [_H, c_int32_p]),
	"jdde_grow_ring": (ctypes.c_int, [_H, ctypes.c_int32]),
	"jdde_clear_status": (ctypes.c_int, [_H, ctypes.c_int32]),
	"jdde_set_attempt_budget": (ctypes.c_int, [_H, ctypes.c_int64]),
	"jdde_host_alloc": (ctypes.c_int, [ctypes.c_size_t, ctypes.POINTER(ctypes.c_void_p)]),
	"jdde_host_free": (ctypes.c_int, [ctypes.c_void_p]),
	"jdde_measure_fp64_peak": (ctypes.c_int, [ctypes.c_int32, ctypes.c_void_p, ctypes.c_size_t, c_double_p]),
	"jdde_launch_count": (ctypes.c_int64, [_H]),
	"jdde_device_result": (ctypes.c_uint64, [_H]),
}

class NetDesc(ctypes.Structure):
	_fields_ = [
		("n", ctypes.c_int64), ("lo", ctypes.c_int64), ("hi", ctypes.c_int64), ("n_edges", ctypes.c_int64),
		("n_node_params", ctypes.c_int32), ("ring_capacity", ctypes.c_int32), ("device", ctypes.c_int32), ("reserved", ctypes.c_int32),
	]


c_uint64_p = ctypes.POINTER(ctypes.c_uint64)
SIGNATURES.update({
	"jdde_net_create": (ctypes.c_int, [ctypes.POINTER(NetDesc), ctypes.POINTER(_H)]),
	"jdde_net_destroy": (ctypes.c_int, [_H]),
	"jdde_net_last_error": (ctypes.c_char_p, [_H]),
	"jdde_net_load_image": (ctypes.c_int, [_H, ctypes.c_void_p, ctypes.c_size_t]),
	"jdde_net_set_stream": (ctypes.c_int, [_H, ctypes.c_void_p]),
	"jdde_net_set_graph": (ctypes.c_int, [_H, c_int64_p, c_int32_p, c_double_p, c_double_p]),
	"jdde_net_set_node_parameters": (ctypes.c_int, [_H, c_double_p]),
	"jdde_net_set_past": (ctypes.c_int, [_H, ctypes.c_int32, c_double_p, c_double_p, c_double_p]),
	"jdde_net_set_integration_parameters": (ctypes.c_int, [_H, ctypes.POINTER(IntParams), ctypes.c_double]),
	"jdde_net_begin": (ctypes.c_int, [_H, ctypes.c_int32, ctypes.c_double, ctypes.c_double, ctypes.c_int64]),
	"jdde_net_attempt": (ctypes.c_int, [_H]),
	"jdde_net_control": (ctypes.c_int, [_H]),
	"jdde_net_poll": (ctypes.c_int, [_H, c_int32_p, c_double_p, c_int32_p]),
	"jdde_net_exchange_pointers": (ctypes.c_int, [_H, c_uint64_p, c_uint64_p]),
	"jdde_net_exchange_handle": (ctypes.c_int, [_H, ctypes.c_void_p, ctypes.c_size_t]),
	"jdde_net_connect_peers": (ctypes.c_int, [_H, ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p, ctypes.c_size_t]),
	"jdde_net_integrate": (ctypes.c_int, [_H, ctypes.c_double, c_double_p, c_int32_p]),
	"jdde_net_integrate_blindly": (ctypes.c_int, [_H, ctypes.c_double, ctypes.c_double, c_double_p, c_int32_p]),
	"jdde_net_recent_state": (ctypes.c_int, [_H, ctypes.c_double, ctypes.c_int32, c_double_p]),
	"jdde_net_grow_ring": (ctypes.c_int, [_H, ctypes.c_int32]),
	"jdde_net_resume": (ctypes.c_int, [_H, c_double_p, c_int32_p]),
	"jdde_net_step_to": (ctypes.c_int, [_H, ctypes.c_double, c_double_p, c_int32_p]),
	"jdde_net_get_history": (ctypes.c_int, [_H, c_int32_p, c_double_p, c_double_p, c_double_p, ctypes.c_int32]),
	"jdde_net_adjust_diff": (ctypes.c_int, [_H, ctypes.c_double, c_double_p, c_double_p]),
	"jdde_net_adjust_diff_begin": (ctypes.c_int, [_H, ctypes.c_double]),
	"jdde_net_jump": (ctypes.c_int, [_H, c_double_p, ctypes.c_double, ctypes.c_double, ctypes.c_int32, c_double_p, c_double_p]),
	"jdde_net_jump_begin": (ctypes.c_int, [_H, c_double_p, ctypes.c_double, ctypes.c_double, ctypes.c_int32]),
	"jdde_net_adjust_diff_end": (ctypes.c_int, [_H, c_double_p, c_double_p]),
	"jdde_net_get_counters": (ctypes.c_int, [_H, c_int64_p, c_int64_p, c_int64_p]),
	"jdde_net_get_dt": (ctypes.c_int, [_H, c_double_p]),
	"jdde_net_launch_count": (ctypes.c_int64, [_H]),
})

_lib = None


def load_library(path=None):
	"""Load libjitcdde_b200.so (building it with nvcc if it is not there yet)."""
	global _lib
	if _lib is not None and path is None:
		return _lib
	if path is None:
		path = _build.build_runtime()      # no-op unless the sources are newer than the library
	lib = ctypes.CDLL(path)
	for name, (restype, argtypes) in SIGNATURES.items():
		fn = getattr(lib, name)      # AttributeError if the library lacks a declared symbol
		fn.restype = restype
		fn.argtypes = argtypes
	_lib = lib
	return lib


class EngineError(RuntimeError):
	pass


def measure_fp64_peak(device=0):
	"""FP64 FMA throughput of the device in FLOP/s (microbenchmark csrc/jdde_peak.cu)"""
	lib = load_library()
	with open(_build.build_peak_image(), "rb") as f:
		image = f.read()
	flops = ctypes.c_double()
	rc = lib.jdde_measure_fp64_peak(device, image, len(image), ctypes.byref(flops))
	if rc:
		raise EngineError(f"jdde_measure_fp64_peak failed ({rc}): {lib.jdde_last_error(None).decode()}")
	return flops.value


class _PinnedBuffer:
	def __init__(self, lib, nbytes):
		self.lib = lib
		self.ptr = ctypes.c_void_p()
		if lib.jdde_host_alloc(nbytes, ctypes.byref(self.ptr)):
			raise EngineError("jdde_host_alloc failed: " + lib.jdde_last_error(None).decode())
		self.nbytes = nbytes

	def __del__(self):
		if self.ptr:
			self.lib.jdde_host_free(self.ptr)
			self.ptr = None


def pinned_empty(shape, dtype=np.float64):
	"""NumPy array in page-locked host memory (jdde_host_alloc); freed with the array."""
	dtype = np.dtype(dtype)
	count = int(np.prod(shape))
	buffer = _PinnedBuffer(load_library(), max(1, count*dtype.itemsize))
	raw = (ctypes.c_char*buffer.nbytes).from_address(buffer.ptr.value)
	array = np.frombuffer(raw, dtype=dtype, count=count).reshape(shape)
	_keep_alive[id(raw)] = buffer
	import weakref
	weakref.finalize(raw, _keep_alive.pop, id(raw), None)
	return array


_keep_alive = {}


def _f64(a):
	return np.ascontiguousarray(a, dtype=np.float64)


def _dp(a):
	return a.ctypes.data_as(c_double_p) if a is not None else None


class Engine:
	"""One ensemble on one GPU: owns a jdde_handle."""

	def __init__(self, n, n_basic, n_sites, n_params, n_members, ring_capacity=64, device=0):
		self.lib = load_library()
		self.n, self.n_basic, self.n_sites, self.n_params = n, n_basic, n_sites, n_params
		self.N = int(n_members)
		self.ring_capacity = ring_capacity
		self.device = device
		desc = Desc(n=n, n_basic=n_basic, n_sites=n_sites, n_params=n_params,
			ring_capacity=ring_capacity, device=device, n_members=self.N)
		self._h = _H()
		rc = self.lib.jdde_create(ctypes.byref(desc), ctypes.byref(self._h))
		if rc:
			message = self.lib.jdde_last_error(None).decode()
			self._h = None
			raise EngineError(f"jdde_create failed ({rc}): {message}")

	def close(self):
		if getattr(self, "_h", None):
			self.lib.jdde_destroy(self._h)
			self._h = None

	__del__ = close

	def _check(self, rc):
		if rc:
			raise EngineError(f"error {rc}: {self.lib.jdde_last_error(self._h).decode()}")

	def load_image(self, path):
		with open(path, "rb") as f:
			image = f.read()
		self._image = image   # keep alive
		self._check(self.lib.jdde_load_rhs(self._h, image, len(image)))

	def set_stream(self, stream_ptr):
		self._check(self.lib.jdde_set_stream(self._h, ctypes.c_void_p(stream_ptr)))

	def synchronize(self):
		self._check(self.lib.jdde_synchronize(self._h))

	def set_past(self, times, states, diffs):
		times, states, diffs = _f64(times), _f64(states), _f64(diffs)
		per_member = times.ndim == 2
		k = times.shape[-1]
		expected = (self.N, k, self.n) if per_member else (k, self.n)
		if states.shape != expected or diffs.shape != expected or (per_member and times.shape[0] != self.N):
			raise ValueError(f"past has wrong shape: times {times.shape}, states {states.shape}, expected states {expected}")
		self._check(self.lib.jdde_set_past(self._h, k, _dp(times), _dp(states), _dp(diffs), int(per_member)))

	def set_parameters(self, params):
		params = _f64(params)
		per_member = params.ndim == 2
		if params.shape != ((self.N, self.n_params) if per_member else (self.n_params,)):
			raise ValueError(f"parameters have wrong shape {params.shape}")
		self._check(self.lib.jdde_set_parameters(self._h, _dp(params), int(per_member)))

	def set_integration_parameters(self, max_delay, **kwargs):
		p = IntParams(**kwargs)
		self._check(self.lib.jdde_set_integration_parameters(self._h, ctypes.byref(p), float(max_delay)))

	def _targets(self, target):
		if np.ndim(target) == 0:
			return np.array([target], dtype=np.float64), 0
		target = _f64(target)
		if target.shape != (self.N,):
			raise ValueError(f"target times must be a scalar or have shape ({self.N},)")
		return target, 1

	def integrate(self, target, out=None, status=None, fetch=True, fetch_status=True):
		"""fetch=False: leave everything on the device, do not synchronise. fetch_status=False: copy the
		states back but not the per-member status (use count_failed / get_status afterwards)."""
		target, pm = self._targets(target)
		if fetch:
			out = np.empty((self.N, self.n)) if out is None else self._checked_out(out, (self.N, self.n))
		if fetch and fetch_status and status is None:
			status = self._pinned_status()
		if not (fetch and fetch_status):
			status = None
		self._check(self.lib.jdde_integrate(self._h, _dp(target), pm, _dp(out) if fetch else None,
			status.ctypes.data_as(c_int32_p) if status is not None else None))
		return out, status

	def integrate_t(self, target, out=None):
		"""integrate() that returns (states, status, times of the members afterwards) after ONE synchronisation"""
		target, pm = self._targets(target)
		out = np.empty((self.N, self.n)) if out is None else self._checked_out(out, (self.N, self.n))
		status = np.empty(self.N, dtype=np.int32)
		t_after = np.empty(self.N)
		self._check(self.lib.jdde_integrate_t(self._h, _dp(target), pm, _dp(out), status.ctypes.data_as(c_int32_p), _dp(t_after)))
		return out, status, t_after

	def integrate_async(self, target, out):
		"""queue an integrate() whose result lands in `out` (page-locked, (N, n) float64) by the next synchronize()"""
		target, pm = self._targets(target)
		self._checked_out(out, (self.N, self.n))
		self._check(self.lib.jdde_integrate_async(self._h, _dp(target), pm, _dp(out)))
		return out

	def wait_result(self, age=0):
		self._check(self.lib.jdde_wait_result(self._h, int(age)))

	def _sample_targets(self, targets):
		targets = _f64(targets)
		if targets.ndim == 1:
			return targets, 0
		if targets.ndim != 2 or targets.shape[0] != self.N:
			raise ValueError(f"sample times must have shape (n_samples,) or ({self.N}, n_samples)")
		return targets, 1

	def integrate_samples(self, targets, out=None, status=None, fetch=True, fetch_status=True):
		"""a series of integrate() calls in one launch; out: (n_samples, N, n). fetch / fetch_status as for integrate()."""
		targets, pm = self._sample_targets(targets)
		K = targets.shape[-1]
		if fetch:
			out = np.empty((K, self.N, self.n)) if out is None else self._checked_out(out, (K, self.N, self.n))
		if fetch and fetch_status and status is None:
			status = self._pinned_status()
		if not (fetch and fetch_status):
			status = None
		self._check(self.lib.jdde_integrate_samples(self._h, _dp(targets), K, pm, _dp(out) if fetch else None,
			status.ctypes.data_as(c_int32_p) if status is not None else None))
		return out, status

	def integrate_samples_async(self, targets, out):
		targets, pm = self._sample_targets(targets)
		K = targets.shape[-1]
		self._checked_out(out, (K, self.N, self.n))
		self._check(self.lib.jdde_integrate_samples_async(self._h, _dp(targets), K, pm, _dp(out)))
		return out

	@staticmethod
	def _checked_out(out, shape):
		"""the runtime writes prod(shape)·8 bytes through this pointer (the kernel itself, if the memory is page-locked)"""
		if not isinstance(out, np.ndarray) or out.shape != tuple(shape) or out.dtype != np.float64 or not out.flags.c_contiguous or not out.flags.writeable:
			raise ValueError(f"out must be a writeable C-contiguous float64 array of shape {tuple(shape)}")
		return out

	def _pinned_status(self):
		"""per-member status lands in page-locked memory (full-bandwidth D2H); callers get a fresh copy only if they keep it"""
		if getattr(self, "_status_buffer", None) is None:
			self._status_buffer = pinned_empty((self.N,), np.int32)
		return self._status_buffer

	def step_on_discontinuities(self, steps, out=None, status=None):
		steps = _f64(steps)
		pm = int(steps.ndim == 2)
		if pm and steps.shape[0] != self.N:
			raise ValueError("steps have wrong shape")
		out = np.empty((self.N, self.n)) if out is None else self._checked_out(out, (self.N, self.n))
		status = np.empty(self.N, dtype=np.int32) if status is None else status
		self._check(self.lib.jdde_step_on_discontinuities(self._h, steps.shape[-1], _dp(steps), pm, _dp(out), status.ctypes.data_as(c_int32_p)))
		return out, status

	def resume(self, lyap=False, samples=None):
		"""repeat the last integrate / step_on_discontinuities / integrate_lyap / integrate_samples call for members that
		stopped (samples: the (n_samples, N, n) result array of the integrate_samples call that is resumed)"""
		status = np.empty(self.N, dtype=np.int32)
		if lyap:
			n_lyap = self.n//self.n_basic-1
			state, lyaps, delta_t = np.empty((self.N, self.n_basic)), np.empty((self.N, n_lyap)), np.empty(self.N)
			self._check(self.lib.jdde_resume(self._h, _dp(state), _dp(lyaps), _dp(delta_t), status.ctypes.data_as(c_int32_p)))
			return state, lyaps, delta_t, status
		out = np.empty((self.N, self.n)) if samples is None else samples
		self._check(self.lib.jdde_resume(self._h, _dp(out), None, None, status.ctypes.data_as(c_int32_p)))
		return out, status

	def integrate_blindly(self, target, step):
		target, pm = self._targets(target)
		out = np.empty((self.N, self.n))
		status = np.empty(self.N, dtype=np.int32)
		self._check(self.lib.jdde_integrate_blindly(self._h, _dp(target), pm, float(step or 0.0), _dp(out), status.ctypes.data_as(c_int32_p)))
		return out, status

	def integrate_blindly_lyap(self, target, step):
		target, pm = self._targets(target)
		n_lyap = self.n//self.n_basic-1
		state = np.empty((self.N, self.n_basic))
		lyaps = np.empty((self.N, n_lyap))
		total = np.empty(self.N)
		status = np.empty(self.N, dtype=np.int32)
		self._check(self.lib.jdde_integrate_blindly_lyap(self._h, _dp(target), pm, float(step or 0.0), _dp(state), _dp(lyaps), _dp(total), status.ctypes.data_as(c_int32_p)))
		return state, lyaps, total, status

	# restricted / transversal Lyapunov exponents
	def set_projector(self, mode, vectors=(), state_components=(), diff_components=(), indices=()):
		"""mode: PROJECT_RESTRICTED (vectors: [(state part, derivative part), …] of length n_basic each, plus components
		handled by zeroing) or PROJECT_TRANSVERSAL (indices = tangent indices)"""
		vectors = _f64(np.array([[v[0], v[1]] for v in vectors], dtype=np.float64).reshape(len(vectors), 2, self.n_basic)) if len(vectors) else np.zeros((0, 2, self.n_basic))
		ints = [np.ascontiguousarray(np.array(x, dtype=np.int32).reshape(-1)) for x in (state_components, diff_components, indices)]
		ip = lambda a: a.ctypes.data_as(c_int32_p)
		self._check(self.lib.jdde_set_projector(self._h, int(mode), _dp(vectors), len(vectors), ip(ints[0]), len(ints[0]), ip(ints[1]), len(ints[1]), ip(ints[2]), len(ints[2])))

	def project(self, delay):
		norms = np.empty(self.N)
		self._check(self.lib.jdde_project(self._h, float(delay), _dp(norms)))
		return norms

	def integrate_blindly_projected(self, target, step):
		target, pm = self._targets(target)
		state, lyap, total = np.empty((self.N, self.n)), np.empty(self.N), np.empty(self.N)
		status = np.empty(self.N, dtype=np.int32)
		self._check(self.lib.jdde_integrate_blindly_projected(self._h, _dp(target), pm, float(step or 0.0), _dp(state), _dp(lyap), _dp(total), status.ctypes.data_as(c_int32_p)))
		return state, lyap, total, status

	def set_max_delay(self, max_delay):
		self._check(self.lib.jdde_set_max_delay(self._h, float(max_delay)))

	def jump(self, amplitude, time, width, forward):
		amplitude = _f64(amplitude)
		amp_pm = int(amplitude.ndim == 2)
		time = _f64(time)
		time_pm = int(time.ndim == 1)
		if not time_pm:
			time = time.reshape(1)
		mi, ma = np.empty((self.N, self.n)), np.empty((self.N, self.n))
		self._check(self.lib.jdde_jump(self._h, _dp(amplitude), amp_pm, _dp(time), time_pm, float(width), int(forward), _dp(mi), _dp(ma)))
		return mi, ma

	def adjust_diff(self, shift_ratio):
		mi, ma = np.empty((self.N, self.n)), np.empty((self.N, self.n))
		self._check(self.lib.jdde_adjust_diff(self._h, float(shift_ratio), _dp(mi), _dp(ma)))
		return mi, ma

	def orthonormalise(self, n_lyap, delay):
		norms = np.empty((self.N, n_lyap))
		self._check(self.lib.jdde_orthonormalise(self._h, n_lyap, float(delay), _dp(norms)))
		return norms

	def integrate_lyap(self, target):
		target, pm = self._targets(target)
		n_lyap = self.n//self.n_basic-1
		state = np.empty((self.N, self.n_basic))
		lyaps = np.empty((self.N, n_lyap))
		delta_t = np.empty(self.N)
		status = np.empty(self.N, dtype=np.int32)
		self._check(self.lib.jdde_integrate_lyap(self._h, _dp(target), pm, _dp(state), _dp(lyaps), _dp(delta_t), status.ctypes.data_as(c_int32_p)))
		return state, lyaps, delta_t, status

	def get_t(self):
		t = np.empty(self.N)
		self._check(self.lib.jdde_get_t(self._h, _dp(t)))
		return t

	def get_dt(self):
		dt = np.empty(self.N)
		self._check(self.lib.jdde_get_dt(self._h, _dp(dt)))
		return dt

	def get_current_state(self):
		out = np.empty((self.N, self.n))
		self._check(self.lib.jdde_get_current_state(self._h, _dp(out)))
		return out

	def get_status(self):
		status = np.empty(self.N, dtype=np.int32)
		self._check(self.lib.jdde_get_status(self._h, status.ctypes.data_as(c_int32_p)))
		return status

	def count_failed(self):
		n = ctypes.c_int64()
		self._check(self.lib.jdde_count_failed(self._h, ctypes.byref(n)))
		return n.value

	def get_counters(self):
		a, r, f = (np.empty(self.N, dtype=np.int64) for _ in range(3))
		self._check(self.lib.jdde_get_counters(self._h, a.ctypes.data_as(c_int64_p), r.ctypes.data_as(c_int64_p), f.ctypes.data_as(c_int64_p)))
		return a, r, f

	def sum_counters(self):
		a, r, f = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64()
		self._check(self.lib.jdde_sum_counters(self._h, ctypes.byref(a), ctypes.byref(r), ctypes.byref(f)))
		return a.value, r.value, f.value

	def get_state(self, member=0):
		count = ctypes.c_int32()
		self._check(self.lib.jdde_get_state(self._h, member, 0, ctypes.byref(count), None, None, None))
		k = count.value
		times, states, diffs = np.empty(k), np.empty((k, self.n)), np.empty((k, self.n))
		self._check(self.lib.jdde_get_state(self._h, member, k, ctypes.byref(count), _dp(times), _dp(states), _dp(diffs)))
		return times, states, diffs

	def grow_ring(self, new_capacity):
		self._check(self.lib.jdde_grow_ring(self._h, int(new_capacity)))
		self.ring_capacity = int(new_capacity)

	def max_live_anchors(self):
		count = ctypes.c_int32()
		self._check(self.lib.jdde_max_live_anchors(self._h, ctypes.byref(count)))
		return count.value

	def ensure_capacity(self, anchors):
		"""grow the ring so that every member can hold `anchors` anchors (for calls that cannot be resumed after an overflow)"""
		capacity = self.ring_capacity
		while capacity < anchors+1:
			capacity *= 2
		if capacity > self.ring_capacity:
			self.grow_ring(capacity)

	def clear_status(self, code):
		self._check(self.lib.jdde_clear_status(self._h, int(code)))

	def set_attempt_budget(self, attempts):
		self._check(self.lib.jdde_set_attempt_budget(self._h, int(attempts)))

	@property
	def launch_count(self):
		return self.lib.jdde_launch_count(self._h)

	@property
	def device_result(self):
		return self.lib.jdde_device_result(self._h)


class NetEngine:
	"""One node-partitioned network (or this rank's part of it) on one GPU: owns a jdde_net handle."""

	def __init__(self, n, lo, hi, n_edges, n_node_params, ring_capacity=128, device=0):
		self.lib = load_library()
		self.n, self.lo, self.hi = int(n), int(lo), int(hi)
		self.ring_capacity = int(ring_capacity)
		desc = NetDesc(n=n, lo=lo, hi=hi, n_edges=n_edges, n_node_params=n_node_params, ring_capacity=ring_capacity, device=device)
		self._h = _H()
		rc = self.lib.jdde_net_create(ctypes.byref(desc), ctypes.byref(self._h))
		if rc:
			message = self.lib.jdde_net_last_error(None).decode()
			self._h = None
			raise EngineError(f"jdde_net_create failed ({rc}): {message}")

	def close(self):
		if getattr(self, "_h", None):
			self.lib.jdde_net_destroy(self._h)
			self._h = None

	__del__ = close

	def _check(self, rc):
		if rc:
			raise EngineError(f"error {rc}: {self.lib.jdde_net_last_error(self._h).decode()}")

	def load_image(self, path):
		with open(path, "rb") as f:
			image = f.read()
		self._image = image
		self._check(self.lib.jdde_net_load_image(self._h, image, len(image)))

	def set_stream(self, stream_ptr):
		self._check(self.lib.jdde_net_set_stream(self._h, ctypes.c_void_p(stream_ptr)))

	def set_graph(self, row_ptr, src, tau, weight):
		row_ptr = np.ascontiguousarray(row_ptr, dtype=np.int64)
		src = np.ascontiguousarray(src, dtype=np.int32)
		tau, weight = _f64(tau), _f64(weight)
		self._check(self.lib.jdde_net_set_graph(self._h, row_ptr.ctypes.data_as(c_int64_p), src.ctypes.data_as(c_int32_p), _dp(tau), _dp(weight)))

	def set_node_parameters(self, par):
		par = _f64(par)
		self._check(self.lib.jdde_net_set_node_parameters(self._h, _dp(par)))

	def set_past(self, times, states, diffs):
		times, states, diffs = _f64(times), _f64(states), _f64(diffs)
		if states.shape != (len(times), self.n) or diffs.shape != states.shape:
			raise ValueError("past has wrong shape")
		self._check(self.lib.jdde_net_set_past(self._h, len(times), _dp(times), _dp(states), _dp(diffs)))

	def set_integration_parameters(self, max_delay, **kwargs):
		p = IntParams(**kwargs)
		self._check(self.lib.jdde_net_set_integration_parameters(self._h, ctypes.byref(p), float(max_delay)))

	def begin(self, mode, target, step=0.0, steps=0):
		self._check(self.lib.jdde_net_begin(self._h, int(mode), float(target), float(step), int(steps)))

	def attempt(self):
		self._check(self.lib.jdde_net_attempt(self._h))

	def control(self):
		self._check(self.lib.jdde_net_control(self._h))

	def poll(self):
		done, status, t = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_double()
		self._check(self.lib.jdde_net_poll(self._h, ctypes.byref(done), ctypes.byref(t), ctypes.byref(status)))
		return bool(done.value), t.value, status.value

	def exchange_pointers(self):
		a, b = ctypes.c_uint64(), ctypes.c_uint64()
		self._check(self.lib.jdde_net_exchange_pointers(self._h, ctypes.byref(a), ctypes.byref(b)))
		return a.value, b.value

	def integrate(self, target):
		out = np.empty(self.n)
		status = ctypes.c_int32()
		self._check(self.lib.jdde_net_integrate(self._h, float(target), _dp(out), ctypes.byref(status)))
		return out, status.value

	def integrate_blindly(self, target, step):
		out = np.empty(self.n)
		status = ctypes.c_int32()
		self._check(self.lib.jdde_net_integrate_blindly(self._h, float(target), float(step or 0.0), _dp(out), ctypes.byref(status)))
		return out, status.value

	def get_history(self):
		"""the live anchors in order: (times [k], states [k, n], diffs [k, n])"""
		count = ctypes.c_int32()
		self._check(self.lib.jdde_net_get_history(self._h, ctypes.byref(count), None, None, None, 0))
		k = count.value
		times, states, diffs = np.empty(k), np.empty((k, self.n)), np.empty((k, self.n))
		self._check(self.lib.jdde_net_get_history(self._h, ctypes.byref(count), _dp(times), _dp(states), _dp(diffs), k))
		return times, states, diffs

	def step_to(self, target):
		"""adaptive steps capped at the target (one entry of the step list of step_on_discontinuities); returns (state, status)"""
		out = np.empty(self.n)
		status = ctypes.c_int32()
		self._check(self.lib.jdde_net_step_to(self._h, float(target), _dp(out), ctypes.byref(status)))
		return out, status.value

	def grow_ring(self, new_capacity):
		self._check(self.lib.jdde_net_grow_ring(self._h, int(new_capacity)))
		self.ring_capacity = int(new_capacity)

	def resume(self):
		"""continue the call that stopped with STATUS_OVERFLOW (after grow_ring); returns (state, status) like that call"""
		out = np.empty(self.n)
		status = ctypes.c_int32()
		self._check(self.lib.jdde_net_resume(self._h, _dp(out), ctypes.byref(status)))
		return out, status.value

	def adjust_diff(self, shift_ratio=1e-4, barrier=None):
		"""adjust_diff of the reference; returns (minima, maxima). Several ranks: `barrier` is called between the two halves."""
		minima, maxima = np.empty(self.n), np.empty(self.n)
		if barrier is None:
			self._check(self.lib.jdde_net_adjust_diff(self._h, float(shift_ratio), _dp(minima), _dp(maxima)))
		else:
			self._check(self.lib.jdde_net_adjust_diff_begin(self._h, float(shift_ratio)))
			barrier()
			self._check(self.lib.jdde_net_adjust_diff_end(self._h, _dp(minima), _dp(maxima)))
		return minima, maxima

	def jump(self, amplitude, time, width=1e-5, forward=True, barrier=None):
		"""jump of the reference; amplitude: [n]; returns (minima, maxima). Several ranks: `barrier` is called between the two halves."""
		amplitude = np.ascontiguousarray(np.broadcast_to(np.asarray(amplitude, dtype=np.float64), (self.n,)))
		minima, maxima = np.empty(self.n), np.empty(self.n)
		if barrier is None:
			self._check(self.lib.jdde_net_jump(self._h, _dp(amplitude), float(time), float(width), int(bool(forward)), _dp(minima), _dp(maxima)))
		else:
			self._check(self.lib.jdde_net_jump_begin(self._h, _dp(amplitude), float(time), float(width), int(bool(forward))))
			barrier()
			self._check(self.lib.jdde_net_adjust_diff_end(self._h, _dp(minima), _dp(maxima)))
		return minima, maxima

	def exchange_handle(self):
		"""CUDA IPC handle of this rank's exchange buffer (bytes), to be passed to the peers"""
		buffer = ctypes.create_string_buffer(64)
		self._check(self.lib.jdde_net_exchange_handle(self._h, buffer, 64))
		return buffer.raw

	def connect_peers(self, rank, world, handles):
		"""handles: list of the `world` handles in rank order; afterwards attempt() exchanges through peer memory"""
		blob = b"".join(handles)
		assert len(blob) == 64*world
		self._check(self.lib.jdde_net_connect_peers(self._h, int(rank), int(world), blob, 64))

	def recent_state(self, t, current_only=False):
		out = np.empty(self.n)
		self._check(self.lib.jdde_net_recent_state(self._h, float(t), int(current_only), _dp(out)))
		return out

	def counters(self):
		a, r, k = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64()
		self._check(self.lib.jdde_net_get_counters(self._h, ctypes.byref(a), ctypes.byref(r), ctypes.byref(k)))
		return a.value, r.value, k.value

	def get_dt(self):
		dt = ctypes.c_double()
		self._check(self.lib.jdde_net_get_dt(self._h, ctypes.byref(dt)))
		return dt.value

	@property
	def launch_count(self):
		return self.lib.jdde_net_launch_count(self._h)

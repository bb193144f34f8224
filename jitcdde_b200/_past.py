"""
Initial-past container: a time-ordered list of anchors (time, state, diff).

The reference uses `Past(chspy.CubicHermiteSpline)`
(/root/reference/jitcdde/past.py:9-50); chspy is not installed here, so this is a
small stand-alone container with the methods the hot path's callers use:
`add`, `constant`, `from_function`, `prepare_anchor` (random tangent vectors for
Lyapunov exponents, past.py:21-50) and interpolation helpers.

Ensemble extension: an anchor's state/diff may carry a leading member axis
(shape (N, n)) and its time may have shape (N,); `as_arrays` then yields
per-member arrays for jdde_set_past.
"""

from collections import namedtuple

import numpy as np

Anchor = namedtuple("Anchor", ["time", "state", "diff"])


def random_direction(d, rng):
	"""uniformly random unit vector (jitcxde_common.numerical.random_direction)"""
	vector = rng.normal(0, 1, d)
	return vector/np.linalg.norm(vector)


def _hermite(t, a, b):
	"""value and derivative of the cubic Hermite interpolant between anchors a and b
	(formulas of /root/reference/jitcdde/jitced_template.c:221-251)"""
	q = b.time-a.time
	x = (t-a.time)/q
	A, B, C, D = a.state, a.diff*q, b.state, b.diff*q
	value = (1-x)*((1-x)*(B*x+(A-C)*(2*x+1))-D*x*x)+C
	diff = ((1-x)*(B-x*3*(2*(A-C)+B+D))+D*x)/q
	return value, diff


class Past(list):
	def __init__(self, n, anchors=(), n_basic=None, rng=None, tangent_indices=()):
		super().__init__()
		self.n = n
		self.n_basic = n_basic or n
		self.tangent_indices = list(tangent_indices)
		self.main_indices = [i for i in range(n or 0) if i not in self.tangent_indices]
		self.rng = rng if rng is not None else np.random.default_rng(42)
		for anchor in anchors:
			self.add(anchor)

	# past.py:21-50
	def prepare_anchor(self, x):
		time, state, diff = x
		state = np.array(state, dtype=float)
		diff = np.array(diff, dtype=float)
		time = np.array(time, dtype=float) if np.ndim(time) else float(time)
		if state.shape != diff.shape:
			raise ValueError("State and derivative have different shapes.")
		if self.tangent_indices and state.shape[-1] < self.n:
			# transversal Lyapunov exponents: one value per group given, the separation function starts in a random direction
			assert state.shape[-1] == len(self.main_indices)
			lead = state.shape[:-1]
			new_state, new_diff = np.empty(lead+(self.n,)), np.empty(lead+(self.n,))
			new_state[..., self.main_indices] = state
			new_state[..., self.tangent_indices] = random_direction(len(self.tangent_indices), self.rng)
			new_diff[..., self.main_indices] = diff
			new_diff[..., self.tangent_indices] = random_direction(len(self.tangent_indices), self.rng)
			state, diff = new_state, new_diff
		if state.shape[-1:] not in [(self.n,), (self.n_basic,)] or state.ndim > 2:
			raise ValueError("State has wrong shape.")
		if self.n_basic < self.n and state.shape[-1] == self.n_basic:
			assert self.n % self.n_basic == 0
			lead = state.shape[:-1]
			new_state = np.empty(lead+(self.n,))
			new_diff = np.empty(lead+(self.n,))
			new_state[..., :self.n_basic] = state
			new_diff[..., :self.n_basic] = diff
			for i in range(1, self.n//self.n_basic):
				block = slice(i*self.n_basic, (i+1)*self.n_basic)
				# one random direction per tangent vector, shared by all members of an ensemble
				new_state[..., block] = random_direction(self.n_basic, self.rng)
				new_diff[..., block] = random_direction(self.n_basic, self.rng)
			state, diff = new_state, new_diff
		return Anchor(time, state, diff)

	def add(self, anchor):
		anchor = self.prepare_anchor(anchor)
		key = np.max(anchor.time)
		index = len(self)
		while index > 0 and np.max(self[index-1].time) > key:
			index -= 1
		if index > 0 and np.any(np.asarray(self[index-1].time) == np.asarray(anchor.time)):
			raise ValueError("Anchor with this time already exists.")
		self.insert(index, anchor)

	def constant(self, state, time=0):
		"""chspy semantics: anchors (time−1, state, 0) and (time, state, 0)"""
		if self:
			raise AssertionError("Past is not empty.")
		state = np.array(state, dtype=float)
		self.add((np.asarray(time)-1. if np.ndim(time) else time-1., state, np.zeros_like(state)))
		self.add((time, state, np.zeros_like(state)))

	def from_function(self, function, times_of_interest, max_anchors=100, tol=5):
		"""
		Anchors from a callable t → state or from symbolic expressions of t.
		Heuristic (the reference delegates to chspy; its anchor placement is not pinned by any
		golden vector): start from times_of_interest, then repeatedly split the
		segment whose midpoint the interpolant misses most until all are within 10**-tol
		(relative) or max_anchors is reached.
		"""
		if self:
			raise AssertionError("Past is not empty.")
		if callable(function):
			def value(time):
				return np.asarray(function(time), dtype=float)
			def derivative(time):
				h = 1e-6*max(1.0, abs(time))
				return (value(time+h)-value(time-h))/(2*h)
		else:
			import sympy
			from ._symbols import t as t_symbol
			expressions = [sympy.sympify(e) for e in function]
			variable = t_symbol
			for expression in expressions:
				for symbol in expression.free_symbols:
					if symbol.name == "t":
						variable = symbol
			f_value = sympy.lambdify(variable, expressions, "math")
			f_diff = sympy.lambdify(variable, [e.diff(variable) for e in expressions], "math")
			value = lambda time: np.asarray(f_value(time), dtype=float)
			derivative = lambda time: np.asarray(f_diff(time), dtype=float)

		make = lambda time: Anchor(float(time), value(time), derivative(time))
		anchors = [make(time) for time in sorted(times_of_interest)]
		threshold = 10.0**(-tol)
		while len(anchors) < max_anchors:
			worst, where = 0.0, None
			for i in range(len(anchors)-1):
				mid = (anchors[i].time+anchors[i+1].time)/2
				interpolated, _ = _hermite(mid, anchors[i], anchors[i+1])
				exact = value(mid)
				deviation = np.max(np.abs(interpolated-exact)/(np.abs(exact)+threshold))
				if deviation > worst:
					worst, where = deviation, i
			if where is None or worst < threshold:
				break
			anchors.insert(where+1, make((anchors[where].time+anchors[where+1].time)/2))
		for anchor in anchors:
			self.add(anchor)

	@property
	def t(self):
		return self[-1].time

	@property
	def times(self):
		return [anchor.time for anchor in self]

	@property
	def per_member(self):
		return any(np.ndim(a.time) > 0 or a.state.ndim > 1 for a in self)

	def get_state(self, time):
		"""interpolated (or extrapolated) state at `time` (single-trajectory pasts)"""
		times = np.array(self.times)
		i = int(np.clip(np.searchsorted(times, time, side="right")-1, 0, len(self)-2))
		return _hermite(time, self[i], self[i+1])[0]

	def as_arrays(self, n_members):
		"""(times, states, diffs) for jdde_set_past: shared ([k], [k][n]) or per member ([N][k], [N][k][n])"""
		k = len(self)
		if not self.per_member:
			return (
				np.array([a.time for a in self], dtype=float),
				np.array([a.state for a in self], dtype=float).reshape(k, self.n),
				np.array([a.diff for a in self], dtype=float).reshape(k, self.n),
			)
		times = np.empty((n_members, k))
		states = np.empty((n_members, k, self.n))
		diffs = np.empty((n_members, k, self.n))
		for j, a in enumerate(self):
			times[:, j] = a.time
			states[:, j, :] = a.state
			diffs[:, j, :] = a.diff
		return times, states, diffs


def _evaluate(spline, time):
	"""(state, diff) of a single-trajectory spline at `time`; outside its range the first/last segment is extrapolated"""
	times = np.array([a.time for a in spline], dtype=float)
	hit = np.nonzero(times == time)[0]
	if len(hit):
		return spline[hit[0]].state, spline[hit[0]].diff
	i = int(np.clip(np.searchsorted(times, time, side="right")-1, 0, len(spline)-2))
	return _hermite(time, spline[i], spline[i+1])


def join(*splines):
	"""
	Glue splines along the component axis (chspy.join, used by the reference at _jitcdde.py:1598): the result has
	anchors at the union of all anchor times; a spline without an anchor at one of these times contributes its
	interpolated (outside its range: extrapolated) value and derivative.
	"""
	times = sorted({float(anchor.time) for spline in splines for anchor in spline})
	joined = Past(n=sum(spline.n for spline in splines))
	for time in times:
		parts = [_evaluate(spline, time) for spline in splines]
		list.append(joined, Anchor(time, np.hstack([p[0] for p in parts]), np.hstack([p[1] for p in parts])))
	return joined

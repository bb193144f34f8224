"""
Largest Lyapunov exponent transversal to a plane or to a synchronisation manifold:
`jitcdde_restricted_lyap` and `jitcdde_transversal_lyap`
(/root/reference/jitcdde/_jitcdde.py:1211-1345 and :1348-1507).

Both integrate a system extended by ONE separation function on the GPU engine like any other
DDE and, after every call of `integrate` (or every step of `integrate_blindly`), apply a
projector to the separation function over the whole delay window, on the device
(csrc/jdde_engine.cuh: `remove_projections`, `normalise_indices`; C ABI: `jdde_set_projector`,
`jdde_project`, `jdde_integrate_blindly_projected`):

  restricted   removes the projections onto the given vectors (unit vectors by zeroing the
               component, others by the reference's dummy-function Gram–Schmidt,
               /root/reference/jitcdde/jitced_template.c:885-1003) and normalises;
  transversal  the system is first transformed to differences between the synchronised
               components, so that only the norm over the tangent indices has to be taken
               (jitced_template.c:1007-1083).

`GroupHandler` restates what the reference imports from jitcxde_common.transversal (not
vendored, not installed here; the transformation is fixed by how _jitcdde.py:1366-1431 uses
it): for a group of synchronised components i0 < i1 < … < ik the transformed system keeps the
synchronised state at position i0 and the differences z(i_j) = δy(i_{j−1}) − δy(i_j) of the
tangent vector at the positions i_j; the tangent vector itself follows from the differences
and from its vanishing projection onto the manifold (Σ_j δy(i_j) = 0).
"""

from warnings import warn

import numpy as np
import sympy

from . import _capi, _codegen
from ._jitcdde import jitcdde, _is_anchor_helper
from ._past import Past
from ._symbols import anchors, current_y, past_y, t

_NO_INTEGRATION = "No actual integration happened in this call of integrate. This happens because the sampling step became smaller than the actual integration step. While this is not a problem per se, I cannot return a meaningful local Lyapunov exponent; therefore I return 0 instead."


class GroupHandler:
	def __init__(self, groups):
		self.groups = [sorted(int(i) for i in group) for group in groups]
		members = [i for group in self.groups for i in group]
		if sorted(members) != list(range(len(members))):
			raise ValueError("Groups must partition the dynamical variables 0 … n−1.")
		self.n = len(members)
		self.main_indices = [group[0] for group in self.groups]
		self.tangent_indices = sorted(i for group in self.groups for i in group[1:])
		self._group_of = {i: group for group in self.groups for i in group}

	def map_to_main(self, index):
		return self._group_of[int(index)][0]

	def extract_main(self, f):
		"""(f, f): the Jacobians need all of f; the main dynamics are its entries at the main indices"""
		f = list(f)
		return f, f

	def iterate(self, iterable):
		"""For position 0 … n−1 of the transformed system: the number of the group whose synchronised state lives
		there, or the pair (entry of the previous member of the group, entry of this member)."""
		entries = list(iterable)
		for position in range(self.n):
			group = self._group_of[position]
			rank = group.index(position)
			if rank == 0:
				yield self.groups.index(group)
			else:
				yield entries[group[rank-1]], entries[position]

	def back_transform(self, z_vector):
		"""tangent vector δy in terms of the differences z (entries of z_vector at the tangent indices)"""
		out = [None]*self.n
		for group in self.groups:
			k = len(group)-1
			first = sum((k+1-j)*z_vector[group[j]] for j in range(1, k+1))/sympy.Integer(k+1) if k else sympy.Integer(0)
			out[group[0]] = first
			for j in range(1, k+1):
				out[group[j]] = out[group[j-1]] - z_vector[group[j]]
		return out


def _shape(values, n_members):
	return values[0] if n_members == 1 else values


class _ProjectedLyap(jitcdde):
	"""what the two classes share: the projector on the device and the result conventions"""

	def compile_C(self, *args, **kwargs):
		kwargs["extra_compile_args"] = [*(kwargs.get("extra_compile_args") or ()), "-DJD_PROJECT=1"]
		super().compile_C(*args, **kwargs)

	def _initiate(self):
		fresh = self._engine is None or self.DDE is None
		super()._initiate()
		if fresh or getattr(self, "_projector_engine", None) is not self._engine:
			self._engine.set_projector(**self._projector())
			self._projector_engine = self._engine

	def _select(self, states):
		raise NotImplementedError

	def integrate(self, target_time):
		"""(state, local largest transversal Lyapunov exponent, integration time), _jitcdde.py:1312-1321, 1471-1480"""
		self._initiate()
		old_t = self._engine.get_t()
		result = jitcdde.integrate(self, target_time)
		result = np.atleast_2d(result)
		delta_t = self._engine.get_t()-old_t
		lyap = np.zeros(self.n_members)
		if np.any(delta_t == 0):
			warn(_NO_INTEGRATION, stacklevel=2)
		if np.any(delta_t != 0):
			# (ensembles: the projector runs for all members; those that did not move report 0 as above)
			norm = self._engine.project(self.max_delay)
			moved = delta_t != 0
			lyap[moved] = np.log(norm[moved])/delta_t[moved]
		return _shape(self._select(result), self.n_members), _shape(lyap, self.n_members), _shape(delta_t, self.n_members)

	def integrate_blindly(self, target_time, step=None):
		"""_jitcdde.py:1323-1345, 1484-1507: the projector after every step, exponent averaged over the steps"""
		self._initiate()
		self.initial_discontinuities_handled = True
		step = step or self.max_step
		assert step > 0, "step must be positive"
		self._room_for_blind_steps(target_time, step)
		state, lyap, total, status = self._engine.integrate_blindly_projected(target_time, step)
		self._check_status(status)
		return _shape(self._select(state), self.n_members), _shape(lyap, self.n_members), _shape(total, self.n_members)


class jitcdde_restricted_lyap(_ProjectedLyap):
	"""
	Largest Lyapunov exponent orthogonal to a plane (reference: _jitcdde.py:1211-1345).
	vectors: pairs (state part, derivative part) spanning the plane whose projection is removed.
	"""

	def __init__(self, f_sym=(), vectors=(), **kwargs):
		self.n_basic = kwargs.pop("n", None)
		helpers = [(sympy.sympify(h[0]), sympy.sympify(h[1])) for h in (kwargs.pop("helpers", None) or [])]
		if any(_is_anchor_helper(helper) for helper in helpers):
			raise NotImplementedError("Cannot use anchor helpers with Lyapunov exponents.")
		kwargs.setdefault("automatic_anchor_helpers", True)

		basic = jitcdde._handle_input(self, f_sym, self.n_basic)()
		self.n_basic = len(basic)
		simplify = kwargs.pop("simplify", None)
		if simplify is None:
			simplify = self.n_basic <= 10
		if not kwargs.get("delays"):
			kwargs["delays"] = _codegen.get_delays(basic, helpers)

		self.vectors, self.state_components, self.diff_components = [], [], []
		for vector in vectors:
			state, diff = np.array(vector[0], dtype=float), np.array(vector[1], dtype=float)
			assert len(state) == self.n_basic
			assert len(diff) == self.n_basic
			if np.count_nonzero(state)+np.count_nonzero(diff) > 1:
				self.vectors.append((state, diff))
			elif np.count_nonzero(state) == 1 and np.count_nonzero(diff) == 0:
				self.state_components.append(int(state.nonzero()[0][0]))
			elif np.count_nonzero(diff) == 1 and np.count_nonzero(state) == 0:
				self.diff_components.append(int(diff.nonzero()[0][0]))
			else:
				raise ValueError("One vector contains only zeros.")

		f_lyap = _codegen.tangent_vector_f(basic, helpers, self.n_basic, 1, kwargs["delays"], simplify=simplify)
		f_lyap += [sympy.Integer(0)]*(2*self.n_basic*len(self.vectors))
		n_basic = self.n_basic
		super().__init__(f_lyap, n=n_basic*(2+2*len(self.vectors)), helpers=(), **kwargs)
		self.n_basic = n_basic
		self.initiate_past()
		assert self.max_delay > 0, "Maximum delay must be positive for calculating Lyapunov exponents."

	def _projector(self):
		return dict(mode=_capi.PROJECT_RESTRICTED, vectors=self.vectors, state_components=self.state_components, diff_components=self.diff_components)

	def _select(self, states):
		return states[:, :self.n_basic]

	def remove_projections(self):
		"""_jitcdde.py:1278-1284; returns the norm of the separation function before its normalisation"""
		self._initiate()
		return _shape(self._engine.project(self.max_delay), self.n_members)

	def set_integration_parameters(self, *args, **kwargs):
		super().set_integration_parameters(*args, **kwargs)
		if (self.state_components or self.diff_components) and not self.atol:
			warn("At least one of your vectors has only one component while your absolute error (atol) is 0. This may cause problems due to spuriously high relative errors. Consider setting atol to some small, non-zero value (e.g., 1e-10) to avoid this.", stacklevel=2)


class jitcdde_transversal_lyap(_ProjectedLyap, GroupHandler):
	"""
	Largest Lyapunov exponent transversal to a synchronisation manifold (reference: _jitcdde.py:1348-1507).
	groups: iterable of iterables of indices of dynamical variables that are synchronised on the manifold. The past
	is given with one value per group (in the order of `groups`), and so is the returned state.
	"""

	def __init__(self, f_sym=(), groups=(), simplify=None, **kwargs):
		GroupHandler.__init__(self, groups)
		n = kwargs.pop("n", None)
		helpers = [(sympy.sympify(h[0]), sympy.sympify(h[1])) for h in (kwargs.pop("helpers", None) or [])]
		if any(_is_anchor_helper(helper) for helper in helpers):
			raise NotImplementedError("Cannot use anchor helpers with Lyapunov exponents.")
		kwargs.setdefault("automatic_anchor_helpers", True)

		f_basic = jitcdde._handle_input(self, f_sym, n)()
		if len(f_basic) != self.n:
			raise ValueError("Groups and f_sym do not have the same number of dynamical variables.")
		f_basic = [_codegen._to_plain_functions(e) for e in _codegen._substitute_helpers(f_basic, [(s, _codegen._to_plain_functions(e)) for s, e in helpers])]
		f_basic, extracted = self.extract_main(f_basic)
		if simplify is None:
			simplify = self.n <= 10
		delays = kwargs.pop("delays", ()) or _codegen.get_delays(f_basic, ())

		past_z, current_z = sympy.Function("past_z"), sympy.Function("current_z")

		def z(index, time=t):
			return current_z(index) if time == t else past_z(time, index, anchors(time))

		tangent_vectors = {delay: self.back_transform([z(i, t-delay) for i in range(self.n)]) for delay in delays}
		jacs = _codegen.jacobians(f_basic, self.n, delays, simplify=simplify)

		def tangent_vector_f():
			for c in range(self.n):
				expression = sympy.Integer(0)
				for delay, jac in zip(delays, jacs, strict=True):
					for k, entry in enumerate(jac[c]):
						expression += entry*tangent_vectors[delay][k]
				yield expression

		def replace(expression, function, new):
			return expression.replace(lambda e: isinstance(e, sympy.Function) and type(e).__name__ == function.__name__, new)

		def finalise(entry):
			entry = replace(entry, current_y, lambda e: current_z(self.map_to_main(e.args[0])))
			entry = replace(entry, past_y, lambda e: past_z(e.args[0], self.map_to_main(e.args[1]), e.args[2]))
			if simplify:
				entry = sympy.simplify(entry, ratio=1)
			entry = replace(entry, current_z, lambda e: current_y(*e.args))
			entry = replace(entry, past_z, lambda e: past_y(*e.args))
			return entry

		f_lyap = []
		for entry in self.iterate(tangent_vector_f()):
			if isinstance(entry, int):
				f_lyap.append(finalise(extracted[self.main_indices[entry]]))
			else:
				f_lyap.append(finalise(entry[0]-entry[1]))

		jitcdde.__init__(self, f_lyap, n=self.n, delays=delays, helpers=(), **kwargs)

	def initiate_past(self):
		self.past = Past(n=self.n, n_basic=self.n_basic, rng=self.rng, tangent_indices=self.tangent_indices)

	def _projector(self):
		return dict(mode=_capi.PROJECT_TRANSVERSAL, indices=self.tangent_indices)

	def _select(self, states):
		return states[:, self.main_indices]

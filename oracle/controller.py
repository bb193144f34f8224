"""
controller.py — the reference's Python-side control loop, restated. TEST INFRASTRUCTURE ONLY.

The reference splits the integrator in two: the C core does one attempted step per
call, and step-size control, accept/reject, the past-within-step iteration,
discontinuity stepping and sampling run in Python
(/root/reference/jitcdde/_jitcdde.py:621-1000). That module cannot be imported here
(symengine, chspy and jitcxde_common are not installed), so its control flow is
restated below and drives the reference's *unmodified* C core as built by
oracle/build_ref.py (`core` = `<module>.dde_integrator(past)`).

Besides the reference's behaviour this class only adds counters (accepted,
rejected) and `root_mode`: 0 = `p**(-1/q)` through libm as in the reference,
1 = the libm-free roots of csrc/jdde_math.h that the GPU can reproduce bit for bit
(`roots` argument = the compiled twin from dde_oracle's shared object, so both
sides call the very same C routine).
"""

import math
from warnings import warn

import numpy as np


class UnsuccessfulIntegration(Exception):
	pass


def propagate_delays(delays, p, threshold=1e-5):
	"""_jitcdde.py:75-86"""
	result = [0]
	if p != 0:
		for delay in propagate_delays(delays, p-1, threshold):
			for other in delays:
				candidate = delay+other
				if all(abs(candidate-old) >= threshold for old in result):
					result.append(candidate)
	return result


def discontinuity_steps(delays, propagations=1, min_distance=1e-5):
	"""Step list of step_on_discontinuities, _jitcdde.py:976-983 (delays include 0)."""
	delays = [float(d) for d in delays]
	if 0 not in delays:
		delays.append(0.0)
	steps = propagate_delays(delays, propagations, min_distance)
	steps.sort()
	steps.remove(0)
	return steps


def _sigmoid(x, tanh=np.tanh):
	return (tanh(x)+1)/2


class Controller:
	def __init__(self, core, max_delay, root_mode=0, roots=None, tanh=None, **kwargs):
		self.DDE = core
		self.tanh = tanh or getattr(roots, "tanh", None) or np.tanh      # jdde_tanh of the compiled oracle for runs the GPU is to reproduce bit for bit
		self.max_delay = max_delay
		self.root_mode = root_mode
		self.roots = roots
		self.accepted = 0
		self.rejected = 0
		self.set_integration_parameters(**kwargs)

	# _jitcdde.py:621-743
	def set_integration_parameters(self,
			atol=1e-10, rtol=1e-5, first_step=1.0, min_step=1e-10, max_step=10.0,
			decrease_threshold=1.1, increase_threshold=0.5, safety_factor=0.9,
			max_factor=5.0, min_factor=0.2, pws_factor=3, pws_atol=0.0, pws_rtol=1e-5,
			pws_max_iterations=10, pws_base_increase_chance=0.1):
		if first_step > max_step:
			first_step = max_step
		if min_step > first_step:
			min_step = first_step
		self.atol, self.rtol = atol, rtol
		self.dt = first_step
		self.min_step, self.max_step = min_step, max_step
		self.decrease_threshold, self.increase_threshold = decrease_threshold, increase_threshold
		self.safety_factor, self.max_factor, self.min_factor = safety_factor, max_factor, min_factor
		self.pws_atol, self.pws_rtol = pws_atol, pws_rtol
		self.pws_max_iterations, self.pws_factor = pws_max_iterations, pws_factor
		self.pws_base_increase_chance = pws_base_increase_chance
		self.q = 3.
		self.last_pws = False
		self.count = 0
		self.increase_credit = 0.0

	# _jitcdde.py:733-741
	def _do_increase(self, p):
		self.increase_credit += p
		if self.increase_credit >= 0.9:
			self.increase_credit = 0.0
			return True
		return False

	# _jitcdde.py:745-765
	def _control_for_min_step(self):
		if self.dt < self.min_step:
			raise UnsuccessfulIntegration(f"dt={self.dt} < min_step={self.min_step}")

	# _jitcdde.py:767-773
	def _increase_chance(self, new_dt):
		q = new_dt/self.last_pws
		is_explicit = _sigmoid(1-q, self.tanh)
		far_from_explicit = _sigmoid(q-self.pws_factor, self.tanh)
		few_iterations = _sigmoid(1-self.count/self.pws_factor, self.tanh)
		profile = is_explicit+far_from_explicit*few_iterations
		return profile + self.pws_base_increase_chance*(1-profile)

	def _root(self, p, order):
		if self.root_mode:
			return self.roots(p, order)
		return p**(-1/order)

	# _jitcdde.py:775-794
	def _adjust_step_size(self):
		p = self.DDE.get_p(self.atol, self.rtol)
		if p > self.decrease_threshold:
			self.dt *= max(self.safety_factor*self._root(p, self.q), self.min_factor)
			self._control_for_min_step()
			return False
		if p <= self.increase_threshold:
			factor = self.safety_factor*self._root(p, self.q+1) if p else self.max_factor
			new_dt = min(self.dt*min(factor, self.max_factor), self.max_step)
			if (not self.last_pws) or self._do_increase(self._increase_chance(new_dt)):
				self.dt = new_dt
				self.count = 0
				self.last_pws = False
		return True

	# _jitcdde.py:838-861
	def try_single_step(self, length):
		self.DDE.get_next_step(length)
		if self.DDE.past_within_step:
			self.last_pws = self.DDE.past_within_step
			if self.dt > self.pws_factor*self.DDE.past_within_step:
				self.dt /= self.pws_factor
				self._control_for_min_step()
				return False
			for count in range(1, self.pws_max_iterations+1):
				self.count = count
				self.DDE.get_next_step(self.dt)
				if self.DDE.check_new_y_diff(self.pws_atol, self.pws_rtol):
					break
			else:
				self.dt /= self.pws_factor
				self._control_for_min_step()
				return False
		return self._adjust_step_size()

	def _step(self, length):
		if self.try_single_step(length):
			self.DDE.accept_step()
			self.accepted += 1
		else:
			self.rejected += 1

	@property
	def t(self):
		return self.DDE.get_t()

	# _jitcdde.py:807-836
	def integrate(self, target_time):
		while self.DDE.get_t() < target_time:
			self._step(self.dt)
		result = self.DDE.get_recent_state(target_time)
		self.DDE.forget(self.max_delay)
		return result

	# _jitcdde.py:985-997
	def step_on_discontinuities(self, steps):
		target_time = self.t
		for step in steps:
			target_time = self.t + step
			while True:
				current_time = self.DDE.get_t()
				if current_time >= target_time:
					break
				self._step(min(self.dt, target_time-current_time))
		result = self.DDE.get_recent_state(target_time)
		self.DDE.forget(self.max_delay)
		return result

	# _jitcdde.py:1002-1044
	def jump(self, amplitude, time, width=1e-5, forward=True):
		n = len(self.DDE.get_current_state())
		amplitude = np.atleast_1d(np.array(np.broadcast_to(np.asarray(amplitude, dtype=float), (n,)), dtype=float))
		if forward:
			return self.DDE.apply_jump(amplitude, time, width)
		return self.DDE.apply_jump(amplitude, time-width, width)

	# _jitcdde.py:863-878 with get_state, :357-371
	def adjust_diff(self, shift_ratio=1e-4):
		self.DDE.forget(self.max_delay)
		past = self.DDE.get_full_state()
		width = shift_ratio*(past[-1][0]-past[-2][0])
		time = past[-1][0]
		return self.jump(0, time, width, forward=False)

	# _jitcdde.py:880-933
	def integrate_blindly(self, target_time, step=None):
		step = step or self.max_step
		total = target_time-self.DDE.get_t()
		if total > 0:
			if total < step:
				step = total
			number = round(total/step)
			dt = total/number
		elif total == 0:
			number, dt = 0, 0
		else:
			raise ValueError("Can’t integrate backwards in time.")
		if number == 0:
			self.adjust_diff()
		else:
			for _ in range(number):
				self.DDE.get_next_step(dt)
				self.DDE.accept_step()
				self.accepted += 1
				self.DDE.forget(self.max_delay)
		return self.DDE.get_current_state().copy()

	# jitcdde_lyap.integrate, _jitcdde.py:1141-1172
	def integrate_lyap(self, target_time, n_basic, n_lyap):
		old_t = self.DDE.get_t()
		result = self.integrate(target_time)[:n_basic]
		delta_t = self.DDE.get_t()-old_t
		if delta_t != 0:
			norms = self.DDE.orthonormalise(n_lyap, self.max_delay)
			lyaps = np.log(norms) / delta_t
		else:
			lyaps = np.zeros(n_lyap)
		return result, lyaps, delta_t

	# jitcdde_restricted_lyap.remove_projections (_jitcdde.py:1278-1284) / DDE.normalise_indices (:1478)
	def project(self, projector):
		if projector["mode"] == "restricted":
			for component in projector["state_components"]:
				self.DDE.remove_state_component(component)
			for component in projector["diff_components"]:
				self.DDE.remove_diff_component(component)
			return self.DDE.remove_projections(self.max_delay, [(np.array(v[0], dtype=float), np.array(v[1], dtype=float)) for v in projector["vectors"]])
		return self.DDE.normalise_indices(self.max_delay)

	# integrate of the restricted / transversal classes, _jitcdde.py:1312-1321, 1471-1480 (the caller selects the components)
	def integrate_projected(self, target_time, projector):
		old_t = self.DDE.get_t()
		result = self.integrate(target_time)
		delta_t = self.DDE.get_t()-old_t
		lyap = np.log(self.project(projector))/delta_t if delta_t != 0 else 0.0
		return result, lyap, delta_t

	# integrate_blindly of the two classes, _jitcdde.py:1323-1345, 1484-1507
	def integrate_blindly_projected(self, target_time, projector, step=None):
		step = step or self.max_step
		total = target_time-self.DDE.get_t()
		if total < step:
			step = total
		number = round(total/step)
		dt = total/number
		instantaneous = []
		for _ in range(number):
			self.DDE.get_next_step(dt)
			self.DDE.accept_step()
			self.accepted += 1
			self.DDE.forget(self.max_delay)
			instantaneous.append(np.log(self.project(projector))/dt)
		return self.DDE.get_current_state().copy(), np.average(instantaneous), total

	def get_state(self):
		self.DDE.forget(self.max_delay)
		return [(a[0], np.array(a[1]), np.array(a[2])) for a in self.DDE.get_full_state()]

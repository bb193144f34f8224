/*
 * net_twin.cpp — TEST SCAFFOLDING: the network engine header (jitcdde_b200/csrc/jdn_engine.cuh) compiled with g++
 * behind the jdde_net_* C ABI on host memory, nodes processed in a loop. Selected only by tests (tests/twin.py).
 */
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <algorithm>

#include "jdn_engine.cuh"

struct jdde_net
{
	jdde_net_desc desc{};
	jdn_problem P{};
	jdn_control ctrl{};
	std::vector<double> times, hist, stage, tent[2], par, tau, weight, scratch;
	std::vector<long long> row_ptr;
	std::vector<int> src, cursor;
	bool have_graph = false, have_past = false, have_int = false;
	long long launches = 0;
	std::string err;
};

static std::string g_err;
static int fail(jdde_net * h, int code, const char * msg) { (h ? h->err : g_err) = msg; return code; }

extern "C" {

int jdde_net_create(const jdde_net_desc * d, jdde_net ** out)
{
	if (d->n_node_params != JDN_NPARAMS) return fail(nullptr, JDDE_E_IMAGE, "twin compiled for another system");
	if (d->ring_capacity < 4 || (d->ring_capacity & (d->ring_capacity-1))) return fail(nullptr, JDDE_E_INVALID, "ring_capacity must be a power of two >= 4");
	jdde_net * h = new jdde_net;
	h->desc = *d;
	size_t const n = (size_t)d->n, cap = (size_t)d->ring_capacity, E = (size_t)d->n_edges;
	h->times.assign(cap, 0.0); h->hist.assign(2*cap*n, 0.0);
	h->stage.assign(2*n, 0.0);
	h->tent[0].assign(2*n, 0.0); h->tent[1].assign(2*n, 0.0);
	h->par.assign(n*(JDN_NPARAMS ? JDN_NPARAMS : 1), 0.0);
	h->row_ptr.assign((size_t)(d->hi-d->lo)+1, 0); h->src.assign(E, 0); h->tau.assign(E, 0.0); h->weight.assign(E, 0.0);
	h->cursor.assign(E, 0); h->scratch.assign(3*E, 0.0);
	jdn_problem & P = h->P;
	P.ctrl = &h->ctrl; P.times = h->times.data(); P.hist = h->hist.data();
	P.stage = h->stage.data(); P.node_par = h->par.data();
	P.tentative[0] = h->tent[0].data(); P.tentative[1] = h->tent[1].data();
	P.world = 1;
	P.iterate = d->lo == 0 && d->hi == d->n;      /* (a rank of a partitioned network, exchange driven from Python: no iteration) */
	P.row_ptr = h->row_ptr.data(); P.src = h->src.data(); P.tau = h->tau.data(); P.weight = h->weight.data(); P.cursor = h->cursor.data();
	P.n = d->n; P.lo = d->lo; P.hi = d->hi; P.cap = d->ring_capacity; P.n_node_params = d->n_node_params;
	*out = h;
	return JDDE_OK;
}
int jdde_net_destroy(jdde_net * h) { delete h; return JDDE_OK; }
const char * jdde_net_last_error(const jdde_net * h) { return h ? h->err.c_str() : g_err.c_str(); }
int jdde_net_load_image(jdde_net *, const void *, size_t) { return JDDE_OK; }
int jdde_net_set_stream(jdde_net *, void *) { return JDDE_OK; }

int jdde_net_set_graph(jdde_net * h, const int64_t * row_ptr, const int32_t * src, const double * tau, const double * weight)
{
	size_t const owned = (size_t)(h->desc.hi-h->desc.lo), E = (size_t)h->desc.n_edges;
	if (row_ptr[0] != 0 || (size_t)row_ptr[owned] != E) return fail(h, JDDE_E_INVALID, "row_ptr does not match n_edges");
	for (size_t k = 0; k <= owned; k++) h->row_ptr[k] = row_ptr[k];
	for (size_t e = 0; e < E; e++)
	{
		if (src[e] < 0 || src[e] >= h->desc.n || !(tau[e] > 0.0)) return fail(h, JDDE_E_INVALID, "edge with source out of range or non-positive delay");
		h->src[e] = src[e]; h->tau[e] = tau[e]; h->weight[e] = weight[e];
	}
	h->have_graph = true;
	return JDDE_OK;
}
int jdde_net_set_node_parameters(jdde_net * h, const double * par) { if (JDN_NPARAMS) memcpy(h->par.data(), par, h->par.size()*sizeof(double)); return JDDE_OK; }

int jdde_net_set_past(jdde_net * h, int32_t k, const double * times, const double * states, const double * diffs)
{
	if (k < 2) return fail(h, JDDE_E_INVALID, "need at least two anchors");
	if (k+1 > h->P.cap) return fail(h, JDDE_E_INVALID, "initial past does not fit into the ring; increase ring_capacity");
	size_t const n = (size_t)h->desc.n;
	memcpy(h->times.data(), times, k*sizeof(double));
	for (int a = 0; a < k; a++)
		for (size_t i = 0; i < n; i++)
		{
			h->hist[2*jdn_pair_index(a, (long long)i, (long long)n)] = states[(size_t)a*n+i];
			h->hist[2*jdn_pair_index(a, (long long)i, (long long)n)+1] = diffs[(size_t)a*n+i];
		}
	std::fill(h->cursor.begin(), h->cursor.end(), k-2);
	jdn_control & C = h->ctrl;
	C.first = 0; C.current = k-1; C.t = times[k-1]; C.last_start = times[0]; C.status = JDDE_STATUS_OK; C.done = 1; C.commit_slot = -1;
	C.accepted = C.rejected = C.attempts = 0;
	C.tentative = 0; C.pws_pass = 0; C.pws_bits = C.diff_bits = 0;
	h->have_past = true;
	return JDDE_OK;
}

int jdde_net_set_integration_parameters(jdde_net * h, const jdde_int_params * p, double max_delay)
{
	jdde_int_params q = *p;
	if (q.first_step > q.max_step) q.first_step = q.max_step;
	if (q.min_step > q.first_step) q.min_step = q.first_step;
	h->ctrl.P = q; h->ctrl.dt = q.first_step; h->ctrl.max_delay = max_delay; h->ctrl.cap = h->P.cap;
	h->ctrl.last_pws = 0.0; h->ctrl.count = 0; h->ctrl.credit = 0.0;
	h->have_int = true;
	return JDDE_OK;
}

static int ready(jdde_net * h)
{
	if (!h->have_graph) return fail(h, JDDE_E_INVALID, "no graph set (jdde_net_set_graph)");
	if (!h->have_past) return fail(h, JDDE_E_INVALID, "no past set (jdde_net_set_past)");
	if (!h->have_int) return fail(h, JDDE_E_INVALID, "integration parameters not set (jdde_net_set_integration_parameters)");
	return 0;
}

int jdde_net_begin(jdde_net * h, int32_t mode, double target, double step, int64_t steps)
{
	if (int rc = ready(h)) return rc;
	jdn_control & C = h->ctrl;
	if (mode == JDN_MODE_CONTINUE) { jdn::continue_call(C); h->launches++; return JDDE_OK; }
	if (C.status != JDDE_STATUS_OK) { C.done = 1; return JDDE_OK; }
	C.mode = mode; C.target = target; C.commit_slot = -1; C.pws = 0; C.p_bits = 0; C.pws_bits = C.diff_bits = 0; C.pws_pass = 0;
	if (mode == JDN_MODE_BLIND) { C.h = step; C.remaining = steps; C.done = steps <= 0; }
	else { C.h = jdn::next_length(C); C.done = !(C.t < target); if (C.done) jdn::forget(h->P, C); }
	C.last_start = C.t;
	h->launches++;
	return JDDE_OK;
}

int jdde_net_attempt(jdde_net * h)
{
	if (int rc = ready(h)) return rc;
	jdn_control & C = h->ctrl;
	h->launches++;
	if (C.done) return JDDE_OK;
	jdn::View V = jdn::view_of(h->P);
	bool const iterate = h->P.iterate;
	jdn::pair * const fresh = jdn::tentative_buffer(h->P, (int)(C.epoch & 1));
	double overlap = 0.0;
	for (long long e = 0; e < h->desc.n_edges; e++)
	{
		if (iterate)
			jdn::lookup_edge_values(h->P, V, C.h, e, h->scratch[3*e], h->scratch[3*e+1], h->scratch[3*e+2], overlap);
		else
		{
			int pws = 0;
			jdn::lookup_edge(h->P, V, C.h, e, h->scratch.data(), pws);
			if (pws) C.pws = 1;
		}
	}
	if (overlap > 0.0)
	{
		union { unsigned long long u; double d; } bits; bits.d = overlap;
		C.pws_bits = bits.u;
	}
	double m = -1.0;
	for (long long i = h->P.lo; i < h->P.hi; i++)
	{
		jdn::pair staged;
		double c = jdn::attempt_node(h->P, V, C.h, i, h->scratch.data(), staged);
		reinterpret_cast<jdn::pair *>(h->P.stage)[i] = staged;
		if (iterate)
		{
			if (C.pws_pass > 0)      /* check_new_y_diff against the result of the pass before */
			{
				double const difference = fabs(staged.state - jdn::last_tentative(h->P)[i].state);
				double const tolerance = C.P.pws_atol + fabs(C.P.pws_rtol*staged.state);
				if (!(tolerance >= difference))
					C.diff_bits = 1;
			}
			fresh[i] = staged;
		}
		if (c >= 0.0 && c > m) m = c;
	}
	if (m > 0.0)
	{
		union { unsigned long long u; double d; } bits; bits.d = m;
		if (bits.u > C.p_bits) C.p_bits = bits.u;
	}
	return JDDE_OK;
}

int jdde_net_control(jdde_net * h)
{
	if (int rc = ready(h)) return rc;
	h->ctrl.epoch++;      /* (the parity of the attempts selects the tentative buffer) */
	jdn::control(h->P);
	int const slot = h->ctrl.commit_slot;
	if (slot >= 0)
	{
		size_t const n = (size_t)h->desc.n;
		for (size_t i = 0; i < n; i++)
		{
			h->hist[2*jdn_pair_index(slot, (long long)i, (long long)n)] = h->stage[2*i];
			h->hist[2*jdn_pair_index(slot, (long long)i, (long long)n)+1] = h->stage[2*i+1];
		}
	}
	h->launches += 2;
	return JDDE_OK;
}

int jdde_net_poll(jdde_net * h, int32_t * done, double * t, int32_t * status)
{
	if (done) *done = h->ctrl.done;
	if (t) *t = h->ctrl.t;
	if (status) *status = h->ctrl.status;
	return JDDE_OK;
}

int jdde_net_exchange_pointers(jdde_net * h, uint64_t * a, uint64_t * c)
{
	if (a) *a = (uint64_t)(uintptr_t)h->stage.data();
	if (c) *c = (uint64_t)(uintptr_t)&h->ctrl.p_bits;
	return JDDE_OK;
}

int jdde_net_recent_state(jdde_net * h, double t, int32_t current_only, double * out)
{
	size_t const n = (size_t)h->desc.n;
	for (size_t i = 0; i < n; i++)
		out[i] = current_only ? h->hist[2*jdn_pair_index(h->ctrl.current & (h->P.cap-1), (long long)i, (long long)n)] : jdn::recent_state(h->P, t, (long long)i);
	return JDDE_OK;
}

static int run(jdde_net * h)
{
	while (!h->ctrl.done)
	{
		if (int rc = jdde_net_attempt(h)) return rc;
		if (int rc = jdde_net_control(h)) return rc;
	}
	return JDDE_OK;
}

int jdde_net_integrate(jdde_net * h, double target, double * out, int32_t * status)
{
	if (int rc = jdde_net_begin(h, JDN_MODE_ADAPTIVE, target, 0.0, 0)) return rc;
	if (int rc = run(h)) return rc;
	if (status) *status = h->ctrl.status;
	if (h->ctrl.status == JDDE_STATUS_OK) return jdde_net_recent_state(h, target, 0, out);
	return JDDE_OK;
}

int jdde_net_integrate_blindly(jdde_net * h, double target, double step, double * out, int32_t * status)
{
	if (int rc = ready(h)) return rc;
	if (!(step > 0)) step = h->ctrl.P.max_step;
	double const total = target-h->ctrl.t;
	if (total < 0) return fail(h, JDDE_E_INVALID, "Can't integrate backwards in time.");
	long long number = 0; double dt = 0;
	if (total > 0) { if (total < step) step = total; number = (long long) rint(total/step); dt = total/number; }
	if (int rc = jdde_net_begin(h, JDN_MODE_BLIND, target, dt, number)) return rc;
	if (int rc = run(h)) return rc;
	if (status) *status = h->ctrl.status;
	return jdde_net_recent_state(h, target, 1, out);
}

int jdde_net_grow_ring(jdde_net * h, int32_t new_cap)
{
	if (new_cap <= h->P.cap || (new_cap & (new_cap-1))) return fail(h, JDDE_E_INVALID, "new capacity must be a larger power of two");
	size_t const n = (size_t)h->desc.n;
	std::vector<double> new_times((size_t)new_cap, 0.0), new_hist(2*(size_t)new_cap*n, 0.0);
	for (size_t i = 0; i < n; i++)
		jdn::regrow_node(h->P, new_hist.data(), new_cap, (long long)i);
	for (int k = h->ctrl.first; k <= h->ctrl.current; k++)
		new_times[(size_t)(k & (new_cap-1))] = h->times[(size_t)(k & (h->P.cap-1))];
	h->times.swap(new_times); h->hist.swap(new_hist);
	h->P.times = h->times.data(); h->P.hist = h->hist.data();
	h->P.cap = new_cap; h->ctrl.cap = new_cap; h->desc.ring_capacity = new_cap;
	return JDDE_OK;
}

static int twin_jump(jdde_net * h, const double * amplitude, double time, double width, double * minima, double * maxima)
{
	jdn_control & C = h->ctrl;
	int const mask = h->P.cap-1;
	if (C.current+2-C.first > h->P.cap) return fail(h, JDDE_E_NOMEM, "anchor ring full; increase ring_capacity");
	if (!(time > h->times[(C.current-1) & mask]) || !(time <= C.t) || !(width > 0)) return fail(h, JDDE_E_INVALID, "jump: the time must lie on the last segment of the history");
	double const new_t = time+width;
	jdn::View const V = jdn::view_of(h->P);
	size_t const n = (size_t)h->desc.n;
	std::vector<jdn::pair> truncated(n), fresh(n);
	for (long long i = h->P.lo; i < h->P.hi; i++)
		jdn::adjust_diff_node(h->P, V, i, time, new_t, amplitude ? amplitude[i] : 0.0, truncated[i], fresh[i]);
	for (long long i = h->P.lo; i < h->P.hi; i++)
	{
		double lo, hi;
		jdn::adjust_diff_commit_node(h->P, i, new_t-time, truncated[i], fresh[i], lo, hi);
		if (minima) minima[i] = lo;
		if (maxima) maxima[i] = hi;
	}
	jdn::adjust_diff_control(h->P, time, new_t);
	h->launches += 3;
	return JDDE_OK;
}
int jdde_net_adjust_diff(jdde_net * h, double shift_ratio, double * minima, double * maxima)
{
	if (int rc = ready(h)) return rc;
	double const width = shift_ratio*(h->ctrl.t-h->times[(h->ctrl.current-1) & (h->P.cap-1)]);
	return twin_jump(h, nullptr, h->ctrl.t-width, width, minima, maxima);
}
int jdde_net_jump(jdde_net * h, const double * amplitude, double time, double width, int32_t forward, double * minima, double * maxima)
{
	if (int rc = ready(h)) return rc;
	return twin_jump(h, amplitude, forward ? time : time-width, width, minima, maxima);
}
int jdde_net_jump_begin(jdde_net * h, const double *, double, double, int32_t) { return fail(h, JDDE_E_INVALID, "host twin: one rank only"); }
int jdde_net_adjust_diff_begin(jdde_net * h, double) { return fail(h, JDDE_E_INVALID, "host twin: one rank only"); }
int jdde_net_adjust_diff_end(jdde_net * h, double *, double *) { return fail(h, JDDE_E_INVALID, "host twin: one rank only"); }

int jdde_net_get_history(jdde_net * h, int32_t * count, double * times, double * states, double * diffs, int32_t capacity)
{
	jdn_control const & C = h->ctrl;
	int const live = C.current-C.first+1;
	if (count) *count = live;
	if (!times) return JDDE_OK;
	if (capacity < live) return fail(h, JDDE_E_INVALID, "room for fewer anchors than there are");
	size_t const n = (size_t)h->desc.n;
	for (int k = 0; k < live; k++)
	{
		int const slot = (C.first+k) & (h->P.cap-1);
		times[k] = h->times[(size_t)slot];
		for (size_t i = 0; i < n; i++)
		{
			states[(size_t)k*n+i] = h->hist[2*jdn_pair_index(slot, (long long)i, (long long)n)];
			diffs[(size_t)k*n+i] = h->hist[2*jdn_pair_index(slot, (long long)i, (long long)n)+1];
		}
	}
	return JDDE_OK;
}

int jdde_net_step_to(jdde_net * h, double target, double * out, int32_t * status)
{
	if (int rc = jdde_net_begin(h, JDN_MODE_CAPPED, target, 0.0, 0)) return rc;
	if (int rc = run(h)) return rc;
	if (status) *status = h->ctrl.status;
	if (h->ctrl.status == JDDE_STATUS_OK) return jdde_net_recent_state(h, target, 0, out);
	return JDDE_OK;
}

int jdde_net_resume(jdde_net * h, double * out, int32_t * status)
{
	if (int rc = jdde_net_begin(h, JDN_MODE_CONTINUE, 0.0, 0.0, 0)) return rc;
	if (int rc = run(h)) return rc;
	if (status) *status = h->ctrl.status;
	if (h->ctrl.mode == JDN_MODE_BLIND) return jdde_net_recent_state(h, h->ctrl.target, 1, out);
	if (h->ctrl.status == JDDE_STATUS_OK) return jdde_net_recent_state(h, h->ctrl.target, 0, out);
	return JDDE_OK;
}

int jdde_net_get_counters(jdde_net * h, int64_t * a, int64_t * r, int64_t * k)
{
	if (a) *a = h->ctrl.accepted;
	if (r) *r = h->ctrl.rejected;
	if (k) *k = h->ctrl.attempts;
	return JDDE_OK;
}
int jdde_net_get_dt(jdde_net * h, double * dt) { *dt = h->ctrl.dt; return JDDE_OK; }
int64_t jdde_net_launch_count(const jdde_net * h) { return h->launches; }

} // extern "C"

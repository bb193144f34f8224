/*
 * jdn_engine.cuh — network engine: one large delay-coupled system, node-parallel.
 *
 * Same numerical method as jdde_engine.cuh (Bogacki–Shampine stages with FSAL, cubic-Hermite history, the
 * reference's step-size controller; /root/reference/jitcdde/jitced_template.c:193-251, 421-498 and
 * /root/reference/jitcdde/_jitcdde.py:775-836, 880-933), organised for ONE trajectory with 10^4–10^6 components:
 *
 *   - every delay is at least as long as a step (enforced: max_step <= smallest delay; a lookup into the step
 *     under way sets a flag and stops the integration), so the three stages of a node depend on the other nodes
 *     only through the HISTORY. An attempted step is therefore two kernels without any inter-node synchronisation
 *     inside: an EDGE-parallel lookup kernel (one thread per in-edge interpolates the delayed source state at the
 *     three stage times in one pass — the three lookups share their two history rows — into a per-edge scratch),
 *     then a NODE-parallel step kernel (stages, new anchor, error contribution from the scratch values);
 *   - the error norm is a max-reduction over nodes (block reduction + one atomic per block, then — with several
 *     GPUs — one all-reduce); accept/reject, the step-size controller, forget and the termination test run in a
 *     one-thread kernel (net_control) on the device, so that the host only polls a flag every few attempts;
 *   - each edge keeps its own search cursor, exactly like a call site of `anchors()` in the reference.
 *
 * Functions are __host__ __device__ (tests compile them with g++ as the host twin).
 */
#ifndef JDN_ENGINE_CUH
#define JDN_ENGINE_CUH

#include <math.h>
#include "jdde_math.h"
#include "jdn_abi.h"

#ifndef JDN_RHS_FILE
#error "JDN_RHS_FILE (generated JDN_LOCAL / JDN_COUPLING macros) must be defined"
#endif
#include JDN_RHS_FILE

#if defined(__CUDACC__)
#  define JDN_FN __host__ __device__ __forceinline__
#else
#  define JDN_FN inline
#endif

#define JDDE_IPOW(x, e) jdde_ipow(x, e)

namespace jdn {

struct alignas(16) pair { double state, diff; };

struct View
{
	double const * times;
	int mask;
	int first, current;
	double t_cur;
};

#if defined(__GNUC__) || defined(__CUDACC__)
#  define JDN_UNLIKELY(x) __builtin_expect(!!(x), 0)
#else
#  define JDN_UNLIKELY(x) (x)
#endif

JDN_FN double anchor_time(View const & V, int const index) { return V.times[index & V.mask]; }

/* tentative buffer b of this rank: the staging buffer of the exchange with several GPUs, else the rank's own pair of buffers */
JDN_FN pair * tentative_buffer(jdn_problem const & P, int const b)
{
	if (P.world > 1)
		return reinterpret_cast<pair *>(jdn_exchange_stage(P.peers[P.rank], b, P.n));
	return reinterpret_cast<pair *>(P.tentative[b]);
}

/* the (state, diff) pairs of the anchor after `current`, if there is one: the result of the attempt before, in the buffer
 * that attempt wrote */
JDN_FN pair const * last_tentative(jdn_problem const & P)
{
	return tentative_buffer(P, (int)((P.ctrl->epoch & 1) ^ 1));
}

/* get_past_anchors (jitced_template.c:193-217) for one edge: back while the anchor is later than t, forward while the
 * next one is earlier and not the last of the list (`last`: the current anchor, or the tentative one after it) */
JDN_FN int walk(View const & V, int ca, double const t, int const last)
{
	while (anchor_time(V, ca) > t && ca > V.first)
		--ca;
	while (anchor_time(V, ca+1) < t && ca+2 <= last)
		++ca;
	return ca;
}

JDN_FN int walk(View const & V, int const ca, double const t) { return walk(V, ca, t, V.current); }

/* get_past_value (jitced_template.c:221-235) of node j on the segment from anchor ca (pair v) to the next one (pair w) */
JDN_FN double hermite(View const & V, int const ca, double const t, pair const v, pair const w)
{
	double const tv = anchor_time(V, ca);
	double const q = anchor_time(V, ca+1)-tv;
	double const x = (t - tv)/q;
	double const a = v.state;
	double const b = v.diff * q;
	double const c = w.state;
	double const d = w.diff * q;
	return (1-x) * ( (1-x) * (b*x + (a-c)*(2*x+1)) - d*x*x) + c;
}

JDN_FN double past_value(jdn_problem const & P, View const & V, int const ca, long long const j, double const t)
{
	pair const * const H = reinterpret_cast<pair const *>(P.hist);
	return hermite(V, ca, t, H[jdn_pair_index(ca & V.mask, j, P.n)], H[jdn_pair_index((ca+1) & V.mask, j, P.n)]);
}

/*
 * Past within step: the lookups of an edge whose delay is shorter than the step (jitced_template.c:193-235 with
 * t > current->time). The last anchor of the list is then either `current` — the lookup extrapolates the last segment —
 * or the result of an attempt that was not accepted (jdn_control::tentative), towards which it interpolates. All three
 * stages see the anchors as they were before the attempt: get_next_step replaces the last anchor only after its last
 * evaluation of f (:465-468), so the lookups of an attempt still do not depend on its stages. Rare, kept out of line.
 */
JDN_FN void lookup_edge_beyond(jdn_problem const & P, View const & V, double const d1, double const d2, double const d3, long long const e, double & v1, double & v2, double & v3)
{
	long long const j = P.src[e];
	bool const tentative = P.ctrl->tentative && P.iterate;
	int const last = tentative ? V.current+1 : V.current;
	pair const * const H = reinterpret_cast<pair const *>(P.hist);
	double const d[3] = { d1, d2, d3 };
	double v[3];
	int ca = P.cursor[e];
	for (int k = 0; k < 3; k++)
	{
		ca = walk(V, ca, d[k], last);
		pair const w = ca+1 > V.current ? last_tentative(P)[j] : H[jdn_pair_index((ca+1) & V.mask, j, P.n)];
		v[k] = hermite(V, ca, d[k], H[jdn_pair_index(ca & V.mask, j, P.n)], w);
	}
	P.cursor[e] = ca;
	v1 = v[0];
	v2 = v[1];
	v3 = v[2];
}

/*
 * Lookup phase of one attempted step, ONE EDGE: the delayed state of the source node at the three stage times
 * t + h/2, t + 3h/4, t + h (get_past_anchors + get_past_value, jitced_template.c:193-235, three times for this
 * edge's call site, in the order the reference evaluates them). The three lookups fall on the same or neighbouring
 * history rows, so they share their gathers. The edge's cursor is updated.
 * Edge-parallel: 10^7 independent threads for a 10^6-node network of in-degree 10.
 *
 * `overlap` grows to the largest t − t_current of a lookup beyond the current anchor (past_within_step,
 * jitced_template.c:211-213); such an edge takes the path above. (A lookup at or before the current anchor never walks
 * onto the segment after it, so the anchors up to `current` are all the common path has to know.)
 */
JDN_FN void lookup_edge_values(jdn_problem const & P, View const & V, double const h, long long const e, double & v1, double & v2, double & v3, double & overlap)
{
	double const tau = P.tau[e];
	double const d1 = (V.t_cur+0.5*h)-tau, d2 = (V.t_cur+0.75*h)-tau, d3 = (V.t_cur+h)-tau;
	if (JDN_UNLIKELY(d3 > V.t_cur))      /* (d1 < d2 < d3) */
	{
		overlap = fmax(overlap, d3-V.t_cur);
		lookup_edge_beyond(P, V, d1, d2, d3, e, v1, v2, v3);
		return;
	}
	long long const j = P.src[e];
	int ca = walk(V, P.cursor[e], d1);
	v1 = past_value(P, V, ca, j, d1);
	ca = walk(V, ca, d2);
	v2 = past_value(P, V, ca, j, d2);
	ca = walk(V, ca, d3);
	v3 = past_value(P, V, ca, j, d3);
	P.cursor[e] = ca;
}

/* the same where nothing iterates: the flag `pws` stops the integration. Tested on h and tau themselves there: with h == tau,
 * (t+h)−tau lands one ulp above t in a few percent of the attempts, and flagging that would stop an integration whose step
 * size has saturated at max_step = smallest delay. */
JDN_FN void lookup_edge_values(jdn_problem const & P, View const & V, double const h, long long const e, double & v1, double & v2, double & v3, int & pws)
{
	double overlap = 0.0;
	if (h > P.tau[e])
		pws = 1;
	lookup_edge_values(P, V, h, e, v1, v2, v3, overlap);
}

JDN_FN void lookup_edge(jdn_problem const & P, View const & V, double const h, long long const e, double * const scratch, int & pws)
{
	lookup_edge_values(P, V, h, e, scratch[3*e], scratch[3*e+1], scratch[3*e+2], pws);
}

/*
 * Step phase, ONE NODE (get_next_step, jitced_template.c:421-472): the three stages with the delayed inputs from
 * `scratch`, the node's part of the tentative anchor, and its contribution |err|/(atol+rtol|y_new|) to the error
 * norm (get_p, :474-498; negative if it is to be skipped). Couplings are summed in the order of the edge list.
 */
JDN_FN double attempt_node(jdn_problem const & P, View const & V, double const h, long long const i, double const * const scratch, pair & staged)
{
	long long const row = i-P.lo;
	long long const e0 = P.row_ptr[row], e1 = P.row_ptr[row+1];
	pair const cur = reinterpret_cast<pair const *>(P.hist)[jdn_pair_index(V.current & V.mask, i, P.n)];
	double const y = cur.state;
	double const k_1 = cur.diff;
	double par[JDN_NPARAMS > 0 ? JDN_NPARAMS : 1];
	for (int k = 0; k < JDN_NPARAMS; k++)
		par[k] = P.node_par[(size_t)k*P.n+i];

	double const t1 = V.t_cur+0.5*h, t2 = V.t_cur+0.75*h, t3 = V.t_cur+h;

	double const y1 = y + 0.5*h*k_1;
	double acc = JDN_LOCAL(y1, t1, par);
	for (long long e = e0; e < e1; e++)
		acc += P.weight[e]*JDN_COUPLING(scratch[3*e], y1);
	double const k_2 = acc;

	double const y2 = y + 0.75*h*k_2;
	acc = JDN_LOCAL(y2, t2, par);
	for (long long e = e0; e < e1; e++)
		acc += P.weight[e]*JDN_COUPLING(scratch[3*e+1], y2);
	double const k_3 = acc;

	double const y_new = y + (h/9.) * (2*k_1+3*k_2+4*k_3);
	acc = JDN_LOCAL(y_new, t3, par);
	for (long long e = e0; e < e1; e++)
		acc += P.weight[e]*JDN_COUPLING(scratch[3*e+2], y_new);
	double const k_4 = acc;

	double const error = fabs((5*k_1-6*k_2-8*k_3+9*k_4) * (h/72.));
	staged.state = y_new;
	staged.diff = k_4;      /* the caller stores it: into the staging buffer, or straight into every peer's (jdn_kernels.cu) */

	double const tolerance = P.ctrl->P.atol + P.ctrl->P.rtol*fabs(y_new);
	if (error != 0.0 || tolerance != 0.0)
		return error/tolerance;
	return -1.0;
}

JDN_FN View view_of(jdn_problem const & P)
{
	jdn_control const & C = *P.ctrl;
	View V;
	V.times = P.times;
	V.mask = P.cap-1;
	V.first = C.first;
	V.current = C.current;
	V.t_cur = C.t;
	return V;
}

/* forget, jitced_template.c:534-561 */
JDN_FN void forget(jdn_problem const & P, jdn_control & C)
{
	int const mask = P.cap-1;
	double const threshold = fmin(C.t - C.max_delay, C.last_start);
	while (C.first+1 < C.current && P.times[(C.first+1) & mask] < threshold)
		C.first++;
}

/* _jitcdde.py:21 */
JDN_FN double sigmoid(double const x) { return (JDDE_TANH(x)+1)/2; }

/* _increase_chance and _do_increase (the credit rule), _jitcdde.py:733-773: after a past-within-step event the step size
 * only grows when enough credit has gathered */
JDN_FN bool pws_allows_increase(jdn_control & C, double const new_dt)
{
	jdde_int_params const & I = C.P;
	double const q = new_dt/C.last_pws;
	double const is_explicit = sigmoid(1-q);
	double const far_from_explicit = sigmoid(q-I.pws_factor);
	double const few_iterations = sigmoid(1-C.count/I.pws_factor);
	double const profile = is_explicit+far_from_explicit*few_iterations;
	double const chance = profile + I.pws_base_increase_chance*(1-profile);
	C.credit += chance;
	if (C.credit >= 0.9)
	{
		C.credit = 0.0;
		return true;
	}
	return false;
}

/* length of the next first pass of try_single_step: dt (integrate, _jitcdde.py:831), min(dt, target − t) (step_on_discontinuities,
 * :993), the fixed step (integrate_blindly) */
JDN_FN double next_length(jdn_control const & C)
{
	if (C.mode == JDN_MODE_BLIND)
		return C.h;
	if (C.mode == JDN_MODE_CAPPED)
		return fmin(C.dt, C.target-C.t);
	return C.dt;
}

/* the time of the anchor after `current` becomes that of the attempt just made (it stays the last anchor) */
JDN_FN void keep_tentative(jdn_problem const & P, jdn_control & C)
{
	C.tentative = 1;
	P.times[(C.current+1) & (P.cap-1)] = C.t + C.h;
}

/* what follows an attempt on the Python side of the reference: try_single_step (_jitcdde.py:838-861) — written as a state
 * machine, one call per get_next_step: C.pws_pass says which pass of it the attempt was —, _adjust_step_size (:775-794),
 * accept_step, the loop test of integrate (:830) or the step count of integrate_blindly (:928-931), forget, and the
 * preparation of the next attempt. Run by ONE thread. */
JDN_FN void control(jdn_problem const & P)
{
	jdn_control & C = *P.ctrl;
	C.commit_slot = -1;
	if (C.done)
		return;
	jdde_int_params const & I = C.P;
	int const mask = P.cap-1;
	union { unsigned long long u; double d; } bits;
	bits.u = C.p_bits;
	double const p = bits.d;
	bits.u = C.pws_bits;
	double const overlap = bits.d;
	bool const differs = C.diff_bits != 0;
	C.attempts++;
	C.p_bits = 0;
	C.pws_bits = 0;
	C.diff_bits = 0;

	if (C.pws)
	{
		/* a delay shorter than the step on a path that does not iterate */
		C.status = JDDE_STATUS_NONFINITE;
		C.done = 1;
		return;
	}

	bool accept = true;
	if (C.mode != JDN_MODE_BLIND)
	{
		bool shrink = false;      /* dt /= pws_factor and a rejected attempt */
		if (C.pws_pass == 0)
		{
			if (overlap != 0.0)
			{
				C.last_pws = overlap;
				if (C.dt > I.pws_factor*overlap || I.pws_max_iterations < 1)
					shrink = true;      /* if possible, shrink the step to make the integration explicit */
				else
				{
					/* repeat the step (with self.dt, :853) until two successive results agree */
					C.pws_pass = C.count = 1;
					C.last_start = C.t;
					keep_tentative(P, C);
					C.h = C.dt;
					return;
				}
			}
		}
		else if (!differs)
			C.pws_pass = 0;      /* converged: on to the error control */
		else if (C.pws_pass >= I.pws_max_iterations)
			shrink = true;
		else
		{
			C.pws_pass = C.count = C.pws_pass+1;
			C.last_start = C.t;
			keep_tentative(P, C);
			C.h = C.dt;
			return;
		}

		if (shrink)
		{
			C.pws_pass = 0;
			C.dt /= I.pws_factor;
			accept = false;
		}
		else if (p > I.decrease_threshold)
		{
			C.dt *= fmax(I.safety_factor*jdde_inv_root3(p), I.min_factor);
			accept = false;
		}
		else if (p <= I.increase_threshold)
		{
			double const factor = (p != 0.0) ? I.safety_factor*jdde_inv_root4(p) : I.max_factor;
			double const new_dt = fmin(C.dt*fmin(factor, I.max_factor), I.max_step);
			if (C.last_pws == 0.0 || pws_allows_increase(C, new_dt))
			{
				C.dt = new_dt;
				C.count = 0;
				C.last_pws = 0.0;
			}
		}
		if (!accept)
		{
			C.rejected++;
			if (C.dt < I.min_step)
			{
				C.status = JDDE_STATUS_MIN_STEP;
				C.done = 1;
				return;
			}
		}
	}

	if (accept)
	{
		/* the staged anchor becomes anchor current+1, the end of the step (t, t+h) */
		C.last_start = C.t;
		C.t = C.t + C.h;
		C.current++;
		P.times[C.current & mask] = C.t;
		C.commit_slot = C.current & mask;
		C.tentative = 0;
		C.accepted++;
		if (C.mode == JDN_MODE_BLIND)
		{
			forget(P, C);
			if (--C.remaining <= 0)
				C.done = 1;
		}
		else if (!(C.t < C.target))
		{
			forget(P, C);
			C.done = 1;
		}
	}
	else
	{
		C.last_start = C.t;
		keep_tentative(P, C);
	}

	if (C.done)
		return;
	if (!(C.dt > 0.0))
	{
		C.status = JDDE_STATUS_NONFINITE;
		C.done = 1;
		return;
	}
	if (C.current+2-C.first > P.cap)
	{
		forget(P, C);
		if (C.current+2-C.first > P.cap)
		{
			C.status = JDDE_STATUS_OVERFLOW;
			C.done = 1;
			return;
		}
	}
	C.h = next_length(C);
}

/* The reference's anchor list is unbounded; the ring is not. A call that stopped with JDDE_STATUS_OVERFLOW goes on after the
 * ring has grown (regrow_node for every node, then this): the controller continues exactly where control() left off. */
JDN_FN void continue_call(jdn_control & C)
{
	if (C.status == JDDE_STATUS_OVERFLOW)
		C.status = JDDE_STATUS_OK;
	C.commit_slot = -1;
	if (C.status != JDDE_STATUS_OK)
	{
		C.done = 1;
		return;
	}
	C.pws = 0;
	C.p_bits = 0;
	C.pws_bits = 0;
	C.diff_bits = 0;
	C.done = 0;
	C.h = next_length(C);
}

/* copy the live anchors of node i into a ring of another capacity */
JDN_FN void regrow_node(jdn_problem const & P, double * const new_hist, int const new_cap, long long const i)
{
	jdn_control const & C = *P.ctrl;
	pair const * const from = reinterpret_cast<pair const *>(P.hist);
	pair * const to = reinterpret_cast<pair *>(new_hist);
	for (int k = C.first; k <= C.current; k++)
		to[jdn_pair_index(k & (new_cap-1), i, P.n)] = from[jdn_pair_index(k & (P.cap-1), i, P.n)];
}

/*
 * adjust_diff (_jitcdde.py:863-878): a jump of amplitude zero at the last anchor, i.e. apply_jump(0, time, width) with
 * time = t_last − width, width = shift_ratio·(t_last − t_before) (jitced_template.c:1122-1174). It replaces the last
 * anchor (t_last) by two: one at `time` with the interpolated state and derivative of the last segment (truncate_past,
 * :1089-1108) and one at new_t = time + width with the interpolated state and the derivative that f gives there — so
 * that the derivative at the start of the integration is the one of the differential equation instead of that of the
 * initial past. `jump` (_jitcdde.py:935-974) is the same with an amplitude added to the state of the second anchor, at
 * time (forward) or time − width, here for a time on the last segment of the history. Phase 1, ONE NODE: both anchors from the history as it is (f's lookups use the untouched list, its
 * call sites keep their cursors); phase 2 writes them (adjust_diff_commit_node), so that no lookup of phase 1 sees them.
 */
JDN_FN void adjust_diff_node(jdn_problem const & P, View const & V, long long const i, double const time, double const new_t, double const change, pair & truncated, pair & fresh)
{
	pair const * const H = reinterpret_cast<pair const *>(P.hist);
	int const left = V.current-1;      /* (the last anchor before new_t: time < new_t <= t_last + an ulp) */
	pair const v = H[jdn_pair_index(left & V.mask, i, P.n)];
	pair const w = H[jdn_pair_index(V.current & V.mask, i, P.n)];
	fresh.state = hermite(V, left, new_t, v, w) + change;      /* (jump: the amplitude of this node; adjust_diff: zero) */
	double par[JDN_NPARAMS > 0 ? JDN_NPARAMS : 1];
	for (int k = 0; k < JDN_NPARAMS; k++)
		par[k] = P.node_par[(size_t)k*P.n+i];
	double acc = JDN_LOCAL(fresh.state, new_t, par);
	long long const row = i-P.lo;
	for (long long e = P.row_ptr[row]; e < P.row_ptr[row+1]; e++)
	{
		double const d = new_t-P.tau[e];
		int const ca = walk(V, P.cursor[e], d);
		P.cursor[e] = ca;
		acc += P.weight[e]*JDN_COUPLING(past_value(P, V, ca, P.src[e], d), fresh.state);
	}
	fresh.diff = acc;
	/* truncate_past: get_past_value and get_past_diff (jitced_template.c:237-251) at `time` */
	truncated.state = hermite(V, left, time, v, w);
	double const tv = anchor_time(V, left);
	double const q = anchor_time(V, left+1)-tv;
	double const x = (time - tv)/q;
	double const a = v.state;
	double const b = v.diff * q;
	double const c = w.state;
	double const d = w.diff * q;
	truncated.diff = ( (1-x)*(b-x*3*(2*(a-c)+b+d)) + d*x ) / q;
}

/* phase 2, one node of the WHOLE network (every rank keeps the whole history): the two anchors go into ring slots current and
 * current+1; minimum and maximum of the node on the segment between them (extrema, jitced_template.c:253-285) */
JDN_FN void adjust_diff_commit_node(jdn_problem const & P, long long const i, double const q, pair const truncated, pair const fresh, double & minimum, double & maximum)
{
	jdn_control const & C = *P.ctrl;
	pair * const H = reinterpret_cast<pair *>(P.hist);
	H[jdn_pair_index(C.current & (P.cap-1), i, P.n)] = truncated;
	H[jdn_pair_index((C.current+1) & (P.cap-1), i, P.n)] = fresh;
	double const a = truncated.state;
	double const b = truncated.diff * q;
	double const c = fresh.state;
	double const d = fresh.diff * q;
	minimum = fmin(a, c);
	maximum = fmax(a, c);
	double const radicant = b*b + b*d + d*d + 3*(a-c)*(3*(a-c) + 2*(b+d));
	if (radicant >= 0)
	{
		double const A = 1/(2*a + b - 2*c + d);
		double const B = a + 2*b/3 - c + d/3;
		for (int sign = -1; sign <= 1; sign += 2)
		{
			double const x = (B+sign*sqrt(radicant)/3) * A;
			if (0 < x && x < 1)
			{
				double const value = (1-x) * ( (1-x) * (b*x + (a-c)*(2*x+1)) - d*x*x) + c;
				minimum = fmin(minimum, value);
				maximum = fmax(maximum, value);
			}
		}
	}
}

/* after phase 2, one thread: the list now ends with the two new anchors, the second one is current (accept_step) */
JDN_FN void adjust_diff_control(jdn_problem const & P, double const time, double const new_t)
{
	jdn_control & C = *P.ctrl;
	int const mask = P.cap-1;
	P.times[C.current & mask] = time;
	C.current++;
	P.times[C.current & mask] = new_t;
	C.t = new_t;
	C.tentative = 0;
	C.commit_slot = -1;
}

/* get_recent_state, jitced_template.c:287-320: Hermite interpolation on the last segment */
JDN_FN double recent_state(jdn_problem const & P, double const t, long long const i)
{
	jdn_control const & C = *P.ctrl;
	int const mask = P.cap-1;
	pair const * const H = reinterpret_cast<pair const *>(P.hist);
	pair const v = H[jdn_pair_index((C.current-1) & mask, i, P.n)];
	pair const w = H[jdn_pair_index(C.current & mask, i, P.n)];
	double const tv = P.times[(C.current-1) & mask];
	double const q = P.times[C.current & mask]-tv;
	double const x = (t - tv)/q;
	double const a = v.state;
	double const b = v.diff * q;
	double const c = w.state;
	double const d = w.diff * q;
	return (1-x) * ( (1-x) * (b*x + (a-c)*(2*x+1)) - d*x*x) + c;
}

} // namespace jdn

#endif

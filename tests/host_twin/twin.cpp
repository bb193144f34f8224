/*
 * twin.cpp — TEST SCAFFOLDING, never part of the product.
 *
 * Compiles the engine header (jitcdde_b200/csrc/jdde_engine.cuh, whose functions are
 * __host__ __device__) with g++ and implements the C ABI of include/jitcdde_b200.h on host
 * memory by looping over the members. This lets the CPU test-suite (`-m "not gpu"`) drive the
 * very same step/controller/lookup code through the very same Python classes and compare it
 * bit for bit with the oracle — without a GPU. The product library (jdde_runtime.cu) has no
 * such path: it fails without a CUDA device. Tests select this twin explicitly by
 * monkeypatching jitcdde_b200._capi.load_library (tests/conftest.py).
 */
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#define JD_PROJECT 1      /* restricted / transversal projections are always part of the twin */
#include "jdde_engine.cuh"

struct jdde_handle
{
	jdde_desc desc{};
	jdde_problem P{};
	std::vector<double> ring, par, dt, last_pws, credit, last_start, out, out2, aux, aux2, samples;
	std::vector<int> first, last, current, count, status, cursor, resume_index;
	std::vector<double> resume_target, last_targets;
	int last_kind = 0, last_pm = 0, last_n_steps = 0;
	std::vector<long long> accepted, rejected, attempts, extra_rhs;
	bool have_past = false, have_params = false, have_int = false;
	long long budget = 1ll << 24;
	long long launches = 0;
	std::string err;
	std::vector<double> projector_vectors;
	std::vector<int> projector_ints;
	jdde_projector projector{};
};

static std::string g_err;
static int fail(jdde_handle * h, int code, const char * msg) { (h ? h->err : g_err) = msg; return code; }

static void bind(jdde_handle * h)
{
	jdde_problem & P = h->P;
	P.ring = h->ring.data(); P.par = h->par.data(); P.dt = h->dt.data(); P.last_pws = h->last_pws.data();
	P.credit = h->credit.data(); P.last_start = h->last_start.data(); P.first = h->first.data(); P.last = h->last.data();
	P.current = h->current.data(); P.count = h->count.data(); P.status = h->status.data(); P.cursor = h->cursor.data();
	P.resume_index = h->resume_index.data(); P.resume_target = h->resume_target.data();
	P.accepted = h->accepted.data(); P.rejected = h->rejected.data(); P.attempts = h->attempts.data(); P.extra_rhs = h->extra_rhs.data();
}

extern "C" {

int jdde_create(const jdde_desc * d, jdde_handle ** out)
{
	if (!d || !out) return fail(nullptr, JDDE_E_INVALID, "null argument");
	if (d->n != JD_N || d->n_basic != JD_NBASIC || d->n_sites != JD_NSITES || d->n_params != JD_NPARAMS)
		return fail(nullptr, JDDE_E_IMAGE, "twin compiled for another system");
	if (d->ring_capacity < 4 || (d->ring_capacity & (d->ring_capacity-1)))
		return fail(nullptr, JDDE_E_INVALID, "ring_capacity must be a power of two >= 4");
	jdde_handle * h = new jdde_handle;
	h->desc = *d;
	size_t const N = (size_t)d->n_members;
	h->ring.assign(N*(size_t)jdde_ring_stride(d->ring_capacity, JD_A), 0.0);
	h->par.assign(N*(JD_NPARAMS ? JD_NPARAMS : 1), 0.0);
	for (auto v : { &h->dt, &h->last_pws, &h->credit, &h->last_start, &h->aux, &h->aux2 }) v->assign(N, 0.0);
	h->out.assign(N*JD_N, 0.0); h->out2.assign(N*JD_N, 0.0);
	for (auto v : { &h->first, &h->last, &h->current, &h->count, &h->status }) v->assign(N, 0);
	h->cursor.assign(N*(JD_NSITES ? JD_NSITES : 1), 0);
	h->resume_index.assign(N, 0); h->resume_target.assign(N, 0.0);
	for (auto v : { &h->accepted, &h->rejected, &h->attempts, &h->extra_rhs }) v->assign(N, 0);
	h->P.n_members = d->n_members; h->P.cap = d->ring_capacity; h->P.anchor_stride = JD_A; h->P.ring_stride = jdde_ring_stride(d->ring_capacity, JD_A);
	bind(h);
	*out = h;
	return JDDE_OK;
}

int jdde_destroy(jdde_handle * h) { delete h; return JDDE_OK; }
const char * jdde_last_error(const jdde_handle * h) { return h ? h->err.c_str() : g_err.c_str(); }
int jdde_load_rhs(jdde_handle *, const void *, size_t) { return JDDE_OK; }
int jdde_set_stream(jdde_handle *, void *) { return JDDE_OK; }
int jdde_synchronize(jdde_handle *) { return JDDE_OK; }

int jdde_set_past(jdde_handle * h, int32_t k, const double * times, const double * states, const double * diffs, int32_t pm)
{
	if (k < 2) return fail(h, JDDE_E_INVALID, "need at least two anchors");
	if (k+1 > h->P.cap) return fail(h, JDDE_E_INVALID, "initial past does not fit into the ring; increase ring_capacity");
	for (long long m = 0; m < h->P.n_members; m++)
		jdde::member_set_past(h->P, m, k, times, states, diffs, pm);
	h->have_past = true; h->launches++;
	return JDDE_OK;
}

int jdde_set_parameters(jdde_handle * h, const double * params, int32_t pm)
{
	size_t const N = (size_t)h->P.n_members;
	for (size_t m = 0; m < N; m++)
		for (int k = 0; k < JD_NPARAMS; k++)
			h->par[(size_t)k*N+m] = pm ? params[m*JD_NPARAMS+k] : params[k];
	h->have_params = true;
	return JDDE_OK;
}

int jdde_set_integration_parameters(jdde_handle * h, const jdde_int_params * p, double max_delay)
{
	if (!(p->decrease_threshold >= 1.0) || !(p->increase_threshold <= 1.0) || !(p->max_factor >= 1.0) || !(p->min_factor <= 1.0)
		|| !(p->safety_factor <= 1.0) || !(p->atol >= 0.0) || !(p->rtol >= 0.0) || !(p->pws_atol >= 0.0) || !(p->pws_rtol >= 0.0)
		|| !(p->pws_max_iterations > 0) || !(p->pws_factor >= 2) || !(p->pws_base_increase_chance >= 0) || !(max_delay >= 0.0))
		return fail(h, JDDE_E_INVALID, "integration parameter out of range (see _jitcdde.py:695-708)");
	h->P.P = *p;
	if (h->P.P.first_step > h->P.P.max_step) h->P.P.first_step = h->P.P.max_step;
	if (h->P.P.min_step > h->P.P.first_step) h->P.P.min_step = h->P.P.first_step;
	h->P.max_delay = max_delay;
	for (long long m = 0; m < h->P.n_members; m++)
		jdde::member_reset_controller(h->P, m);
	h->have_int = true; h->launches++;
	return JDDE_OK;
}

static int ready(jdde_handle * h)
{
	if (!h->have_past) return fail(h, JDDE_E_INVALID, "no past set (jdde_set_past)");
	if (!h->have_int) return fail(h, JDDE_E_INVALID, "integration parameters not set (jdde_set_integration_parameters)");
	if (JD_NPARAMS && !h->have_params) return fail(h, JDDE_E_INVALID, "control parameters not set (jdde_set_parameters)");
	return 0;
}

static void give(jdde_handle * h, double * out, const std::vector<double> & src, size_t count, int32_t * status)
{
	if (out) memcpy(out, src.data(), count*sizeof(double));
	if (status) memcpy(status, h->status.data(), (size_t)h->P.n_members*sizeof(int32_t));
}

static void remember(jdde_handle * h, int kind, const double * targets, size_t count, int pm, int n_steps)
{
	h->last_kind = kind; h->last_pm = pm; h->last_n_steps = n_steps;
	h->last_targets.assign(targets, targets+count);
}

extern "C" int jdde_get_t(jdde_handle * h, double * t);

static void run_integrate(jdde_handle * h, const double * targets, int pm, int n_steps, int resume, int n_out)
{
	for (long long m = 0; m < h->P.n_members; m++)
		jdde::member_integrate(h->P, m, targets, pm, n_steps, resume, h->out.data(), n_out, h->budget);
	h->launches++;
}

int jdde_integrate(jdde_handle * h, const double * target, int32_t pm, double * out, int32_t * status)
{
	if (int rc = ready(h)) return rc;
	remember(h, 1, target, pm ? (size_t)h->P.n_members : 1, pm, 0);
	run_integrate(h, target, pm, 0, 0, JD_N);
	give(h, out, h->out, (size_t)h->P.n_members*JD_N, status);
	return JDDE_OK;
}

int jdde_integrate_t(jdde_handle * h, const double * target, int32_t pm, double * out, int32_t * status, double * t_after)
{
	if (int rc = jdde_integrate(h, target, pm, out, status)) return rc;
	return jdde_get_t(h, t_after);
}

int jdde_integrate_async(jdde_handle * h, const double * target, int32_t pm, double * out)
{
	if (!out) return fail(h, JDDE_E_INVALID, "null argument");
	int const rc = jdde_integrate(h, target, pm, out, nullptr);
	h->last_kind = 0;
	return rc;
}

int jdde_integrate_samples(jdde_handle * h, const double * targets, int32_t n_samples, int32_t pm, double * out, int32_t * status)
{
	if (int rc = ready(h)) return rc;
	if (n_samples < 1) return fail(h, JDDE_E_INVALID, "n_samples must be positive");
	remember(h, 4, targets, (pm ? (size_t)h->P.n_members : 1)*(size_t)n_samples, pm, -n_samples);
	if (h->samples.size() < (size_t)n_samples*h->P.n_members*JD_N)
		h->samples.resize((size_t)n_samples*h->P.n_members*JD_N);
	for (long long m = 0; m < h->P.n_members; m++)
		jdde::member_integrate(h->P, m, targets, pm, -n_samples, 0, h->samples.data(), JD_N, h->budget);
	h->launches++;
	give(h, out, h->samples, (size_t)n_samples*h->P.n_members*JD_N, status);
	return JDDE_OK;
}

int jdde_integrate_samples_async(jdde_handle * h, const double * targets, int32_t n_samples, int32_t pm, double * out)
{
	if (!out) return fail(h, JDDE_E_INVALID, "null argument");
	int const rc = jdde_integrate_samples(h, targets, n_samples, pm, out, nullptr);
	h->last_kind = 0;
	return rc;
}

int jdde_wait_result(jdde_handle * h, int32_t age)
{
	if (!h) return JDDE_E_INVALID;
	if (age < 0 || age > 1) return fail(h, JDDE_E_INVALID, "age must be 0 or 1");
	return JDDE_OK;
}

int jdde_step_on_discontinuities(jdde_handle * h, int32_t n_steps, const double * steps, int32_t pm, double * out, int32_t * status)
{
	if (int rc = ready(h)) return rc;
	if (n_steps < 1) return fail(h, JDDE_E_INVALID, "n_steps must be positive");
	remember(h, 2, steps, (pm ? (size_t)h->P.n_members : 1)*(size_t)n_steps, pm, n_steps);
	run_integrate(h, steps, pm, n_steps, 0, JD_N);
	give(h, out, h->out, (size_t)h->P.n_members*JD_N, status);
	return JDDE_OK;
}

int jdde_integrate_blindly(jdde_handle * h, const double * target, int32_t pm, double step, double * out, int32_t * status)
{
	if (int rc = ready(h)) return rc;
	for (long long m = 0; m < h->P.n_members; m++)
		jdde::member_integrate_blindly(h->P, m, target, pm, step, h->out.data(), JD_N, nullptr, nullptr);
	h->launches++;
	give(h, out, h->out, (size_t)h->P.n_members*JD_N, status);
	return JDDE_OK;
}

int jdde_integrate_blindly_lyap(jdde_handle * h, const double * target, int32_t pm, double step, double * out, double * lyaps, double * total_time, int32_t * status)
{
	if (int rc = ready(h)) return rc;
	if (JD_N == JD_NBASIC) return fail(h, JDDE_E_INVALID, "integrate_blindly_lyap needs n_basic < n");
	for (long long m = 0; m < h->P.n_members; m++)
		jdde::member_integrate_blindly(h->P, m, target, pm, step, h->out.data(), JD_NBASIC, h->out2.data(), h->aux2.data());
	h->launches++;
	give(h, lyaps, h->out2, (size_t)h->P.n_members*(JD_N/JD_NBASIC-1), nullptr);
	give(h, total_time, h->aux2, (size_t)h->P.n_members, nullptr);
	give(h, out, h->out, (size_t)h->P.n_members*JD_NBASIC, status);
	return JDDE_OK;
}

int jdde_set_projector(jdde_handle * h, int32_t mode, const double * vectors, int32_t n_vectors,
	const int32_t * state_components, int32_t n_state, const int32_t * diff_components, int32_t n_diff, const int32_t * indices, int32_t n_indices)
{
	if (mode == JDDE_PROJECT_RESTRICTED)
	{
		if (JD_N != JD_NBASIC*(2+2*n_vectors)) return fail(h, JDDE_E_INVALID, "restricted projector needs n = n_basic·(2 + 2·n_vectors) (_jitcdde.py:1270-1274)");
		n_indices = 0;
	}
	else if (mode == JDDE_PROJECT_TRANSVERSAL)
	{
		if (n_indices < 1 || !indices) return fail(h, JDDE_E_INVALID, "transversal projector needs tangent indices");
		n_vectors = n_state = n_diff = 0;
	}
	else
		return fail(h, JDDE_E_INVALID, "unknown projector mode");
	h->projector_vectors.assign(vectors, vectors+(size_t)n_vectors*2*JD_NBASIC);
	h->projector_ints.clear();
	h->projector_ints.insert(h->projector_ints.end(), state_components, state_components+n_state);
	h->projector_ints.insert(h->projector_ints.end(), diff_components, diff_components+n_diff);
	h->projector_ints.insert(h->projector_ints.end(), indices, indices+n_indices);
	jdde_projector & R = h->projector;
	R.mode = mode;
	R.n_vectors = n_vectors; R.vectors = h->projector_vectors.data();
	R.n_state_components = n_state; R.state_components = h->projector_ints.data();
	R.n_diff_components = n_diff; R.diff_components = h->projector_ints.data()+n_state;
	R.n_indices = n_indices; R.indices = h->projector_ints.data()+n_state+n_diff;
	return JDDE_OK;
}

int jdde_project(jdde_handle * h, double delay, double * norms)
{
	if (int rc = ready(h)) return rc;
	if (!h->projector.mode) return fail(h, JDDE_E_INVALID, "no projector set (jdde_set_projector)");
	for (long long m = 0; m < h->P.n_members; m++)
		jdde::member_project(h->P, m, h->projector, delay, h->aux2.data());
	h->launches++;
	give(h, norms, h->aux2, (size_t)h->P.n_members, nullptr);
	return JDDE_OK;
}

int jdde_integrate_blindly_projected(jdde_handle * h, const double * target, int32_t pm, double step, double * out, double * lyap, double * total_time, int32_t * status)
{
	if (int rc = ready(h)) return rc;
	if (!h->projector.mode) return fail(h, JDDE_E_INVALID, "no projector set (jdde_set_projector)");
	for (long long m = 0; m < h->P.n_members; m++)
		jdde::member_integrate_blindly(h->P, m, target, pm, step, h->out.data(), JD_N, h->aux.data(), h->aux2.data(), &h->projector);
	h->launches++;
	give(h, lyap, h->aux, (size_t)h->P.n_members, nullptr);
	give(h, total_time, h->aux2, (size_t)h->P.n_members, nullptr);
	give(h, out, h->out, (size_t)h->P.n_members*JD_N, status);
	return JDDE_OK;
}

int jdde_set_max_delay(jdde_handle * h, double max_delay) { if (!(max_delay >= 0.0)) return fail(h, JDDE_E_INVALID, "max_delay must be non-negative"); h->P.max_delay = max_delay; return JDDE_OK; }

int jdde_jump(jdde_handle * h, const double * amplitude, int32_t apm, const double * time, int32_t tpm, double width, int32_t forward, double * minima, double * maxima)
{
	if (int rc = ready(h)) return rc;
	if (!(width >= 0)) return fail(h, JDDE_E_INVALID, "width must be non-negative (_jitcdde.py:1035)");
	for (long long m = 0; m < h->P.n_members; m++)
		jdde::member_jump(h->P, m, 0, amplitude, apm, time, tpm, width, forward, h->out.data(), h->out2.data());
	h->launches++;
	give(h, minima, h->out, (size_t)h->P.n_members*JD_N, nullptr);
	give(h, maxima, h->out2, (size_t)h->P.n_members*JD_N, nullptr);
	return JDDE_OK;
}

int jdde_adjust_diff(jdde_handle * h, double shift_ratio, double * minima, double * maxima)
{
	if (int rc = ready(h)) return rc;
	for (long long m = 0; m < h->P.n_members; m++)
		jdde::member_jump(h->P, m, 1, nullptr, 0, nullptr, 0, shift_ratio, 0, h->out.data(), h->out2.data());
	h->launches++;
	give(h, minima, h->out, (size_t)h->P.n_members*JD_N, nullptr);
	give(h, maxima, h->out2, (size_t)h->P.n_members*JD_N, nullptr);
	return JDDE_OK;
}

#if JD_NBASIC != JD_N
/* JITCDDE_TWIN_TEAM = staging capacity in anchors (>= 4; shorter than a history: streamed in tiles) runs the team code of csrc/jdde_team.cuh with a
 * one-thread team instead of member_orthonormalise */
static void orthonormalise_all(jdde_handle * h, int n_lyap, double delay, double * norms, double * old_t, double * lyaps, double * delta_t)
{
	const char * const team = getenv("JITCDDE_TWIN_TEAM");
	if (!team)
	{
		for (long long m = 0; m < h->P.n_members; m++)
			jdde::member_orthonormalise(h->P, m, n_lyap, delay, norms, old_t, lyaps, delta_t);
		return;
	}
	int const capacity = atoi(team);
	std::vector<double> shared(1 + (size_t)(capacity | 1)*(1 + 2*JD_NBASIC*n_lyap));
	jdde::Team const T = { 0, 1, shared.data() };
	for (long long m = 0; m < h->P.n_members; m++)
		jdde::team_member_orthonormalise(T, h->P, m, n_lyap, delay, norms, old_t, lyaps, delta_t, shared.data()+1, capacity);
}
#endif

int jdde_orthonormalise(jdde_handle * h, int32_t n_lyap, double delay, double * norms)
{
#if JD_NBASIC != JD_N
	if (int rc = ready(h)) return rc;
	orthonormalise_all(h, n_lyap, delay, h->out2.data(), nullptr, nullptr, nullptr);
	h->launches++;
	give(h, norms, h->out2, (size_t)h->P.n_members*n_lyap, nullptr);
	return JDDE_OK;
#else
	(void)n_lyap; (void)delay; (void)norms;
	return fail(h, JDDE_E_INVALID, "orthonormalise needs n_basic < n and (n_lyap+1)·n_basic <= n");
#endif
}

static int run_lyap(jdde_handle * h, const double * target, int pm, int resume, double * out_state, double * lyaps, double * delta_t, int32_t * status)
{
#if JD_NBASIC != JD_N
	int const n_lyap = JD_N/JD_NBASIC-1;
	if (!resume)
		for (long long m = 0; m < h->P.n_members; m++)
			h->aux[m] = h->ring[(size_t)m*h->P.ring_stride + (size_t)(h->current[m] & (h->P.cap-1))*JD_A];
	run_integrate(h, target, pm, 0, resume, JD_NBASIC);
	orthonormalise_all(h, n_lyap, h->P.max_delay, h->out2.data(), h->aux.data(), h->out2.data(), h->aux2.data());
	h->launches += 2;
	give(h, lyaps, h->out2, (size_t)h->P.n_members*n_lyap, nullptr);
	give(h, delta_t, h->aux2, (size_t)h->P.n_members, nullptr);
	give(h, out_state, h->out, (size_t)h->P.n_members*JD_NBASIC, status);
	return JDDE_OK;
#else
	(void)target; (void)pm; (void)resume; (void)out_state; (void)lyaps; (void)delta_t; (void)status;
	return fail(h, JDDE_E_INVALID, "integrate_lyap needs n_basic < n");
#endif
}

int jdde_integrate_lyap(jdde_handle * h, const double * target, int32_t pm, double * out_state, double * lyaps, double * delta_t, int32_t * status)
{
	if (int rc = ready(h)) return rc;
	remember(h, 3, target, pm ? (size_t)h->P.n_members : 1, pm, 0);
	return run_lyap(h, target, pm, 0, out_state, lyaps, delta_t, status);
}

int jdde_resume(jdde_handle * h, double * out_state, double * lyaps, double * delta_t, int32_t * status)
{
	if (int rc = ready(h)) return rc;
	std::vector<double> const targets = h->last_targets;
	if (h->last_kind == 1 || h->last_kind == 2)
	{
		run_integrate(h, targets.data(), h->last_pm, h->last_n_steps, 1, JD_N);
		give(h, out_state, h->out, (size_t)h->P.n_members*JD_N, status);
		return JDDE_OK;
	}
	if (h->last_kind == 3)
		return run_lyap(h, targets.data(), h->last_pm, 1, out_state, lyaps, delta_t, status);
	if (h->last_kind == 4)
	{
		int const n_samples = -h->last_n_steps;
		for (long long m = 0; m < h->P.n_members; m++)
			jdde::member_integrate(h->P, m, targets.data(), h->last_pm, h->last_n_steps, 1, h->samples.data(), JD_N, h->budget);
		h->launches++;
		give(h, out_state, h->samples, (size_t)n_samples*h->P.n_members*JD_N, status);
		return JDDE_OK;
	}
	return fail(h, JDDE_E_INVALID, "nothing to resume");
}

int jdde_get_t(jdde_handle * h, double * t)
{
	for (long long m = 0; m < h->P.n_members; m++)
		t[m] = h->ring[(size_t)m*h->P.ring_stride + (size_t)(h->current[m] & (h->P.cap-1))*JD_A];
	return JDDE_OK;
}
int jdde_get_dt(jdde_handle * h, double * dt) { memcpy(dt, h->dt.data(), h->dt.size()*sizeof(double)); return JDDE_OK; }
int jdde_get_current_state(jdde_handle * h, double * s)
{
	for (long long m = 0; m < h->P.n_members; m++)
		memcpy(s+(size_t)m*JD_N, &h->ring[(size_t)m*h->P.ring_stride + (size_t)(h->last[m] & (h->P.cap-1))*JD_A+1], JD_N*sizeof(double));
	return JDDE_OK;
}
int jdde_get_status(jdde_handle * h, int32_t * status) { give(h, nullptr, h->out, 0, status); return JDDE_OK; }

int jdde_count_failed(jdde_handle * h, int64_t * count)
{
	int64_t n = 0;
	for (auto s : h->status) n += s != JDDE_STATUS_OK;
	*count = n;
	return JDDE_OK;
}

int jdde_get_counters(jdde_handle * h, int64_t * a, int64_t * r, int64_t * f)
{
	for (long long m = 0; m < h->P.n_members; m++)
	{
		if (a) a[m] = h->accepted[m];
		if (r) r[m] = h->rejected[m];
		if (f) f[m] = 3*h->attempts[m]+h->extra_rhs[m];
	}
	return JDDE_OK;
}

int jdde_sum_counters(jdde_handle * h, int64_t * a, int64_t * r, int64_t * f)
{
	int64_t sa = 0, sr = 0, sf = 0;
	for (long long m = 0; m < h->P.n_members; m++)
	{
		sa += h->accepted[m]; sr += h->rejected[m]; sf += 3*h->attempts[m]+h->extra_rhs[m];
	}
	if (a) *a = sa;
	if (r) *r = sr;
	if (f) *f = sf;
	return JDDE_OK;
}

int jdde_get_state(jdde_handle * h, int64_t member, int32_t max_anchors, int32_t * n_anchors, double * times, double * states, double * diffs)
{
	if (!h->have_past) return fail(h, JDDE_E_INVALID, "no past set (jdde_set_past)");
	if (member < 0 || member >= h->P.n_members || !n_anchors) return fail(h, JDDE_E_INVALID, "member out of range or null argument");
	if (h->have_int)
		for (long long m = 0; m < h->P.n_members; m++)
			jdde::member_forget(h->P, m);
	int const first = h->first[member], count = h->last[member]-first+1;
	*n_anchors = count;
	if (count > max_anchors || !times) return JDDE_OK;
	for (int k = 0; k < count; k++)
	{
		const double * a = &h->ring[(size_t)member*h->P.ring_stride + (size_t)((first+k) & (h->P.cap-1))*JD_A];
		times[k] = a[0];
		memcpy(states+(size_t)k*JD_N, a+1, JD_N*sizeof(double));
		memcpy(diffs+(size_t)k*JD_N, a+1+JD_N, JD_N*sizeof(double));
	}
	return JDDE_OK;
}

/*
 * Per-step access to the engine's functions for member 0 — the counterpart of the methods of the reference's dde_integrator
 * (jitced_template.c:1183-1208), which the product does not export (the boundary is one level higher). Used by
 * tests/test_random_actions.py, the restatement of the reference's tests/compare_C_and_Python_core.py: random sequences
 * of core calls on the reference's C core and on this engine must give identical results.
 */
static jdde::Member * step_member(jdde_handle * h)
{
	static thread_local jdde::Member M;
	return &M;
}

static void step_resync(jdde::Member & M)
{
	/* anchors were modified in place: refresh the copies of the current and of the last anchor */
	M.invalidate_windows();
	double const * const c = M.slot(M.current);
	double const * const l = M.slot(M.last);
	M.cur_t = c[0];
	M.new_t = l[0];
	for (int i = 0; i < JD_N; i++)
	{
		M.cur_y[i] = c[1+i]; M.cur_d[i] = c[1+JD_N+i];
		M.new_y[i] = l[1+i]; M.new_d[i] = l[1+JD_N+i];
	}
}

int twin_step_begin(jdde_handle * h)
{
	if (int rc = ready(h)) return rc;
	jdde::Member & M = *step_member(h);
	jdde::load_member(h->P, 0, M);
	step_resync(M);
	h->aux[0] = 0.0;                                   /* have_old */
	return JDDE_OK;
}

void twin_get_next_step(jdde_handle * h, double dt)
{
	jdde::Member & M = *step_member(h);
	if (M.last != M.current)
	{
		/* the anchor replaced now is the reference's old_last (jitced_template.c:106-114) */
		for (int i = 0; i < JD_N; i++) h->out2[i] = M.new_y[i];
		h->aux[0] = 1.0;
	}
	jdde::ensure_room(M, h->P.max_delay);
	jdde::get_next_step(M, dt);
}

double twin_get_t(jdde_handle * h) { return step_member(h)->cur_t; }
double twin_past_within_step(jdde_handle * h) { return step_member(h)->pws; }
double twin_get_p(jdde_handle * h, double atol, double rtol) { return jdde::get_p(*step_member(h), atol, rtol); }

int twin_check_new_y_diff(jdde_handle * h, double atol, double rtol)
{
	if (h->aux[0] == 0.0) return 0;                    /* old_last == NULL, jitced_template.c:513 */
	return jdde::check_new_y_diff(*step_member(h), h->out2.data(), atol, rtol) ? 1 : 0;
}

void twin_accept_step(jdde_handle * h)
{
	jdde::accept_step(*step_member(h));
	h->aux[0] = 0.0;
}

void twin_forget(jdde_handle * h, double delay) { jdde::forget(*step_member(h), delay); }

void twin_get_recent_state(jdde_handle * h, double t, double * out) { jdde::get_recent_state(*step_member(h), t, out, JD_N); }

void twin_get_current_state(jdde_handle * h, double * out)
{
	jdde::Member & M = *step_member(h);
	for (int i = 0; i < JD_N; i++) out[i] = M.slot(M.last)[1+i];
}

int twin_get_full_state(jdde_handle * h, int max_anchors, double * times, double * states, double * diffs)
{
	jdde::Member & M = *step_member(h);
	int const count = M.last-M.first+1;
	if (count > max_anchors) return count;
	for (int k = 0; k < count; k++)
	{
		double const * const a = M.slot(M.first+k);
		times[k] = a[0];
		memcpy(states+(size_t)k*JD_N, a+1, JD_N*sizeof(double));
		memcpy(diffs+(size_t)k*JD_N, a+1+JD_N, JD_N*sizeof(double));
	}
	return count;
}

void twin_truncate(jdde_handle * h, double time) { jdde::truncate_past(*step_member(h), time); }

void twin_apply_jump(jdde_handle * h, const double * change, double time, double width, double * minima, double * maxima)
{
	jdde::apply_jump(*step_member(h), h->P, change, time, width, minima, maxima);
}

#if JD_NBASIC != JD_N
void twin_orthonormalise(jdde_handle * h, int n_lyap, double delay, double * norms)
{
	jdde::Member & M = *step_member(h);
	jdde::orthonormalise(M, n_lyap, delay, norms);
	step_resync(M);
}

double twin_remove_projections(jdde_handle * h, double delay, const double * vector)
{
	jdde::Member & M = *step_member(h);
	double const norm = jdde::remove_projections(M, delay, 1, vector);
	step_resync(M);
	return norm;
}

void twin_remove_component(jdde_handle * h, int diff, int index)
{
	jdde::Member & M = *step_member(h);
	for (int ca = M.first; ca <= M.last; ca++)
		M.slot(ca)[1+(diff ? JD_N : 0)+JD_NBASIC+index] = 0.0;
	step_resync(M);
}
#endif

double twin_normalise_indices(jdde_handle * h, double delay, int n_indices, const int * indices)
{
	jdde::Member & M = *step_member(h);
	double const norm = jdde::normalise_indices(M, delay, n_indices, indices);
	step_resync(M);
	return norm;
}

int jdde_grow_ring(jdde_handle * h, int32_t new_cap)
{
	if (new_cap <= h->P.cap || (new_cap & (new_cap-1))) return fail(h, JDDE_E_INVALID, "new capacity must be a larger power of two");
	size_t const N = (size_t)h->P.n_members;
	size_t const new_stride = (size_t)jdde_ring_stride(new_cap, JD_A);
	std::vector<double> nr(N*new_stride, 0.0);
	if (h->have_past)
		for (size_t m = 0; m < N; m++)
		{
			for (int k = h->first[m]; k <= h->last[m]; k++)
				memcpy(&nr[m*new_stride + (size_t)(k & (new_cap-1))*JD_A], &h->ring[m*h->P.ring_stride + (size_t)(k & (h->P.cap-1))*JD_A], JD_A*sizeof(double));
			if (h->status[m] == JDDE_STATUS_OVERFLOW) h->status[m] = JDDE_STATUS_OK;
		}
	h->ring.swap(nr);
	h->P.cap = new_cap; h->desc.ring_capacity = new_cap; h->P.ring_stride = (long long)new_stride;
	bind(h);
	return JDDE_OK;
}

int jdde_clear_status(jdde_handle * h, int32_t code)
{
	for (auto & s : h->status) if (s == code) s = JDDE_STATUS_OK;
	return JDDE_OK;
}
int jdde_max_live_anchors(jdde_handle * h, int32_t * count)
{
	int most = 0;
	for (long long m = 0; m < h->P.n_members; m++)
		if (h->last[m]-h->first[m]+1 > most)
			most = h->last[m]-h->first[m]+1;
	*count = most;
	return JDDE_OK;
}

int jdde_set_attempt_budget(jdde_handle * h, int64_t attempts) { if (attempts < 1) return fail(h, JDDE_E_INVALID, "budget must be positive"); h->budget = attempts; return JDDE_OK; }
int jdde_host_alloc(size_t bytes, void ** out) { *out = malloc(bytes ? bytes : 1); return *out ? JDDE_OK : JDDE_E_NOMEM; }
int jdde_host_free(void * p) { free(p); return JDDE_OK; }
int jdde_measure_fp64_peak(int32_t, const void *, size_t, double * flops) { if (flops) *flops = 0.0; return JDDE_E_CUDA; }
int64_t jdde_launch_count(const jdde_handle * h) { return h->launches; }
uint64_t jdde_device_result(const jdde_handle *) { return 0; }

} // extern "C"

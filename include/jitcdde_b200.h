/*
 * jitcdde_b200.h — C ABI of the B200 DDE integration engine (libjitcdde_b200.so).
 *
 * Boundary. The reference (neurophysik/jitcdde) crosses from Python into native
 * code through a generated CPython extension type `dde_integrator`
 * (/root/reference/jitcdde/jitced_template.c:1183-1232), 5–7 calls per integration
 * step (/root/reference/jitcdde/_jitcdde.py:830-861). Those calls cannot cross a
 * PCIe boundary per step, so the boundary sits one level higher: every entry point
 * below is a whole-ensemble operation with the meaning of one method of the
 * reference's Python class `jitcdde` / `jitcdde_lyap`
 * (/root/reference/jitcdde/_jitcdde.py); the per-step methods of `dde_integrator`
 * it absorbs are listed with each function.
 *
 * Conventions
 *   - plain C, no CUDA or torch types; all buffers are caller-owned HOST memory,
 *     C-contiguous, float64 / int32 / int64. The handle owns device memory and
 *     the stream.
 *   - an ensemble has n_members independent trajectories ("members") of the same
 *     n-dimensional DDE; arrays with a member axis are [n_members][...].
 *     `per_member = 0` broadcasts one value to all members.
 *   - every function returns 0 on success or a JDDE_E_* code; jdde_last_error()
 *     gives the text. Integration problems of single members are not errors of
 *     the call: they are reported in `status[member]` (JDDE_STATUS_*), the
 *     ensemble counterpart of the reference raising UnsuccessfulIntegration
 *     (_jitcdde.py:135-139, 745-765).
 *   - output pointers may be NULL: the result then stays in device memory and the
 *     call returns without synchronising (see jdde_synchronize).
 *   - a handle is not thread-safe. There is no CPU fallback: without a CUDA
 *     device jdde_create fails with JDDE_E_CUDA.
 */
#ifndef JITCDDE_B200_H
#define JITCDDE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define JDDE_ABI_VERSION 2

/* return codes */
#define JDDE_OK            0
#define JDDE_E_INVALID     1   /* bad argument / wrong call order */
#define JDDE_E_CUDA        2   /* CUDA runtime error (text in jdde_last_error) */
#define JDDE_E_IMAGE       3   /* kernel image does not match the descriptor */
#define JDDE_E_NOMEM       4

/* per-member status */
#define JDDE_STATUS_OK         0
#define JDDE_STATUS_MIN_STEP   1   /* dt < min_step: the reference raises UnsuccessfulIntegration (_jitcdde.py:745-765) */
#define JDDE_STATUS_OVERFLOW   2   /* anchor ring full; call jdde_grow_ring and repeat the call */
#define JDDE_STATUS_NONFINITE  3   /* state became NaN/inf */
#define JDDE_STATUS_BACKWARDS  4   /* integrate_blindly with target < t (reference: ValueError, _jitcdde.py:897) */
#define JDDE_STATUS_BUDGET     5   /* more attempted steps in one call than jdde_set_attempt_budget allows (kernel watchdog) */

typedef struct jdde_handle jdde_handle;

/* What the generated kernel image was compiled for (cf. the Jinja2 variables the
 * reference renders its template with, _jitcdde.py:562-573). */
typedef struct jdde_desc
{
	int32_t n;              /* dimension incl. tangent vectors                         */
	int32_t n_basic;        /* dimension of the DDE proper (n_basic == n: no Lyapunov)  */
	int32_t n_sites;        /* distinct anchors(...) lookup sites (anchor_mem_length)   */
	int32_t n_params;       /* control parameters (control_pars)                        */
	int32_t ring_capacity;  /* anchors kept per member, power of two >= 4               */
	int32_t device;         /* CUDA device ordinal                                      */
	int64_t n_members;      /* ensemble size on this device                             */
} jdde_desc;

/* set_integration_parameters, _jitcdde.py:621-638 (same names, same defaults on the Python side).
 * pws_fuzzy_seed: 0 = pws_fuzzy_increase=False, the deterministic credit rule (_jitcdde.py:733-741); otherwise
 * pws_fuzzy_increase=True (:730-731): after a past-within-step event the step size is increased with the probability
 * the reference computes, drawn from a counter-based generator keyed by this seed, the member index and the member's
 * number of right-hand-side evaluations (csrc/jdde_math.h: jdde_uniform) — reproducible, independent of scheduling. */
typedef struct jdde_int_params
{
	double atol, rtol, first_step, min_step, max_step;
	double decrease_threshold, increase_threshold, safety_factor, max_factor, min_factor;
	double pws_factor, pws_atol, pws_rtol, pws_base_increase_chance;
	int32_t pws_max_iterations;
	int32_t pws_fuzzy_seed;
} jdde_int_params;

/* ---- life cycle ------------------------------------------------------------------- */

/* Replaces dde_integrator_init (jitced_template.c:608-652) minus the past, which is set separately. */
int jdde_create(const jdde_desc * desc, jdde_handle ** out);
int jdde_destroy(jdde_handle * h);                        /* dde_integrator_dealloc, :563-573 */
const char * jdde_last_error(const jdde_handle * h);      /* h may be NULL: error of the last failed jdde_create */

/* Load the kernels generated for this DDE (a cubin for sm_100a produced by
 * jitcdde_b200/_build.py from csrc/jdde_kernels.cuh + the generated right-hand side).
 * Replaces _compile_and_load / module import (_jitcdde.py:575, 585-588). */
int jdde_load_rhs(jdde_handle * h, const void * image, size_t len);

/* Use a caller-provided CUDA stream (a cudaStream_t passed as void*), e.g. torch's
 * current stream, so that callers can bracket calls with their own events. */
int jdde_set_stream(jdde_handle * h, void * cuda_stream);
int jdde_synchronize(jdde_handle * h);

/* ---- definition of the problem ----------------------------------------------------- */

/* Initial past as anchors (time, state[n], diff[n]), n_anchors >= 2; resets integration progress.
 * times[n_anchors], states[n_anchors][n], diffs[n_anchors][n]; with per_member each has a leading
 * member axis. Replaces initiate_past_from_list (jitced_template.c:575-606) as driven by
 * add_past_point(s) / constant_past (_jitcdde.py:282-324). */
int jdde_set_past(jdde_handle * h, int32_t n_anchors, const double * times, const double * states, const double * diffs, int32_t per_member);

/* params[n_params] or [n_members][n_params]. set_parameters, jitced_template.c:167-185, _jitcdde.py:594-614. */
int jdde_set_parameters(jdde_handle * h, const double * params, int32_t per_member);

/* Also resets the step size to first_step and the past-within-step memory, as the reference does
 * (_jitcdde.py:710-728). max_delay is the `forget` horizon (_jitcdde.py:835). */
int jdde_set_integration_parameters(jdde_handle * h, const jdde_int_params * p, double max_delay);

/* Change only the `forget` horizon (attribute jitcdde.max_delay), keeping step size and counters. */
int jdde_set_max_delay(jdde_handle * h, double max_delay);

/* ---- integration -------------------------------------------------------------------- */

/* jitcdde.integrate (_jitcdde.py:807-836): adaptive steps until t >= target, state at target by
 * Hermite interpolation of the last step, then forget(max_delay). Absorbs get_t, get_next_step,
 * past_within_step, get_p, check_new_y_diff, accept_step, get_recent_state, forget
 * (jitced_template.c:187, 421, 474, 502, 526, 287, 534) and try_single_step / _adjust_step_size
 * (_jitcdde.py:775-861). target: 1 value or [n_members]; out_state [n_members][n]; status [n_members]. */
int jdde_integrate(jdde_handle * h, const double * target, int32_t per_member, double * out_state, int32_t * status);

/* jdde_integrate that also reports where every member stands afterwards (get_t, jitced_template.c:187, which the reference's
 * integrate consults before every call, _jitcdde.py:824): t_after[n_members], fetched with the states in the same
 * synchronisation — a sampling loop over single calls then costs one host↔device round trip per call, not three. */
int jdde_integrate_t(jdde_handle * h, const double * target, int32_t per_member, double * out_state, int32_t * status, double * t_after);

/* Pipelined sampling: as jdde_integrate, but returns as soon as the work is queued. If out_state is page-locked
 * memory the GPU can address (jdde_host_alloc, cudaHostAlloc, registered memory), the kernel writes the states
 * straight into it (also in jdde_integrate); otherwise they are copied on a second stream while the next call's
 * kernel already runs, two device result buffers alternating. Up to two results are in flight. out_state is valid and member failures can be asked for (jdde_count_failed) after jdde_synchronize.
 * The sampling loop of the reference (examples/mackey_glass.py:76-79: one integrate() per sample) maps to a
 * sequence of these calls; members that stop with a full ring cannot be resumed in this mode. */
int jdde_integrate_async(jdde_handle * h, const double * target, int32_t per_member, double * out_state);
/* Wait until the out_state of the jdde_integrate_async call `age` calls ago (0: the latest, 1: the one before) has
 * arrived, without waiting for younger work. */
int jdde_wait_result(jdde_handle * h, int32_t age);

/* A series of jitcdde.integrate calls in ONE launch — the user's sampling loop `for T in times: DDE.integrate(T)`
 * (/root/reference/examples/mackey_glass.py:76-78; per call _jitcdde.py:807-836): every member integrates to
 * targets[0], interpolates its state there, forgets, goes on to targets[1], … exactly as n_samples separate calls of
 * jdde_integrate would make it do (same steps, same states, bit for bit), but without waiting for the other members
 * in between and with its controller state loaded and stored once. targets: [n_samples] or, with per_member,
 * [n_members][n_samples] (ascending); out_states: [n_samples][n_members][n]; status: [n_members]. Not available for
 * kernel images that integrate with a group of lanes per member (Lyapunov systems): JDDE_E_INVALID. */
int jdde_integrate_samples(jdde_handle * h, const double * targets, int32_t n_samples, int32_t per_member, double * out_states, int32_t * status);
/* ... queued like jdde_integrate_async (out_states page-locked; jdde_wait_result / jdde_synchronize as there). */
int jdde_integrate_samples_async(jdde_handle * h, const double * targets, int32_t n_samples, int32_t per_member, double * out_states);

/* step_on_discontinuities (_jitcdde.py:985-997): for each of the n_steps sorted step lengths,
 * adaptive steps capped so that one lands on t + step; steps [n_steps] or [n_members][n_steps]
 * (the host side computes them with _propagate_delays, :75-86). */
int jdde_step_on_discontinuities(jdde_handle * h, int32_t n_steps, const double * steps, int32_t per_member, double * out_state, int32_t * status);

/* Repeat the last jdde_integrate / jdde_step_on_discontinuities / jdde_integrate_lyap call after
 * members stopped with JDDE_STATUS_OVERFLOW (then call jdde_grow_ring first) or JDDE_STATUS_BUDGET
 * (jdde_clear_status): stopped members continue where they were, finished members are not stepped
 * again. Outputs as for the original call (lyaps/delta_t may be NULL unless it was integrate_lyap). */
int jdde_resume(jdde_handle * h, double * out_state, double * lyaps, double * delta_t, int32_t * status);

/* integrate_blindly (_jitcdde.py:880-933): fixed steps, no error control; step <= 0 means max_step.
 * target == t performs adjust_diff() as the reference does. out_state = state of the last anchor. */
int jdde_integrate_blindly(jdde_handle * h, const double * target, int32_t per_member, double step, double * out_state, int32_t * status);

/* jump (_jitcdde.py:1002-1044) → apply_jump, truncate_past, extrema (jitced_template.c:1122-1174,
 * 1089-1108, 253-285). amplitude [n] or [n_members][n]; time 1 value or [n_members];
 * minima/maxima [n_members][n]. */
int jdde_jump(jdde_handle * h, const double * amplitude, int32_t amplitude_per_member, const double * time, int32_t time_per_member, double width, int32_t forward, double * minima, double * maxima);

/* adjust_diff (_jitcdde.py:863-878): zero-amplitude backward jump of width shift_ratio·(t_last − t_prev). */
int jdde_adjust_diff(jdde_handle * h, double shift_ratio, double * minima, double * maxima);

/* ---- Lyapunov exponents (n_basic < n) ------------------------------------------------- */

/* orthonormalise (jitced_template.c:846-878 with :657-844): norms [n_members][n_lyap]. */
int jdde_orthonormalise(jdde_handle * h, int32_t n_lyap, double delay, double * norms);

/* jitcdde_lyap.integrate (_jitcdde.py:1141-1172): out_state [n_members][n_basic],
 * lyaps [n_members][n_lyap] = log(norms)/delta_t (0 where delta_t == 0), delta_t [n_members]. */
int jdde_integrate_lyap(jdde_handle * h, const double * target, int32_t per_member, double * out_state, double * lyaps, double * delta_t, int32_t * status);

/* jitcdde_lyap.integrate_blindly (_jitcdde.py:1191-1209): fixed steps, orthonormalising after every
 * step; lyaps = average over the steps of log(norms)/dt; total_time [n_members]. */
int jdde_integrate_blindly_lyap(jdde_handle * h, const double * target, int32_t per_member, double step, double * out_state, double * lyaps, double * total_time, int32_t * status);

/* ---- restricted and transversal Lyapunov exponents ------------------------------------- */

#define JDDE_PROJECT_RESTRICTED 1
#define JDDE_PROJECT_TRANSVERSAL 2

/* What jitcdde_restricted_lyap / jitcdde_transversal_lyap do to the separation function after integrating
 * (the image must have been built with -DJD_PROJECT). The arrays are copied.
 * mode JDDE_PROJECT_RESTRICTED (needs n = n_basic·(2 + 2·n_vectors)): remove_state_component,
 *   remove_diff_component (jitced_template.c:973-1003) for the listed components, then remove_projections
 *   (:885-971) for `vectors` [n_vectors][2][n_basic] (state part, derivative part), as
 *   jitcdde_restricted_lyap.remove_projections calls them (_jitcdde.py:1278-1284);
 * mode JDDE_PROJECT_TRANSVERSAL: normalise_indices (jitced_template.c:1007-1083) over `indices` (the reference
 *   compiles them into the module as tangent_indices, _jitcdde.py:571). */
int jdde_set_projector(jdde_handle * h, int32_t mode, const double * vectors, int32_t n_vectors,
	const int32_t * state_components, int32_t n_state_components, const int32_t * diff_components, int32_t n_diff_components,
	const int32_t * indices, int32_t n_indices);

/* apply the projector once with the given delay (the reference passes max_delay); norms [n_members] = norm of the
 * separation function before its normalisation */
int jdde_project(jdde_handle * h, double delay, double * norms);

/* integrate_blindly of the two classes (_jitcdde.py:1323-1345, 1484-1507): fixed steps, projector after every step;
 * out_state [n_members][n] (state of the last anchor; the caller selects n_basic or main_indices),
 * lyap [n_members] = average over the steps of log(norm)/dt, total_time [n_members]. */
int jdde_integrate_blindly_projected(jdde_handle * h, const double * target, int32_t per_member, double step, double * out_state, double * lyap, double * total_time, int32_t * status);

/* ---- inspection ----------------------------------------------------------------------- */

int jdde_get_t(jdde_handle * h, double * t);                              /* get_t, jitced_template.c:187-190 */
int jdde_get_dt(jdde_handle * h, double * dt);                            /* attribute jitcdde.dt */
int jdde_get_current_state(jdde_handle * h, double * state);              /* get_current_state, :330-334; [n_members][n] */
int jdde_get_status(jdde_handle * h, int32_t * status);
/* Number of members whose status is not JDDE_STATUS_OK, counted on the device (8 bytes back instead of the array). */
int jdde_count_failed(jdde_handle * h, int64_t * count);
/* accepted steps, rejected attempts, right-hand-side evaluations per member (the reference has no counters) */
int jdde_get_counters(jdde_handle * h, int64_t * accepted, int64_t * rejected, int64_t * rhs_evals);
/* Sum over members, computed on the device (for the throughput metric). */
int jdde_sum_counters(jdde_handle * h, int64_t * accepted, int64_t * rejected, int64_t * rhs_evals);

/* get_state (_jitcdde.py:357-371): forget(max_delay) for all members, then all anchors of one member.
 * times[max_anchors], states/diffs[max_anchors][n]; *n_anchors receives the number written
 * (or needed, if larger than max_anchors). get_full_state, jitced_template.c:336-352. */
int jdde_get_state(jdde_handle * h, int64_t member, int32_t max_anchors, int32_t * n_anchors, double * times, double * states, double * diffs);

/* Largest number of anchors any member holds at the moment (last − first + 1): lets the caller size the ring BEFORE calls that
 * cannot be resumed after an overflow — integrate_blindly, jump, adjust_diff (the reference's list just grows,
 * jitced_template.c:56-73). */
int jdde_max_live_anchors(jdde_handle * h, int32_t * count);

/* Double (or more) the ring of every member, keeping all anchors, and clear JDDE_STATUS_OVERFLOW.
 * The reference's list is unbounded (malloc per anchor, jitced_template.c:434). */
int jdde_grow_ring(jdde_handle * h, int32_t new_capacity);

/* Reset every member whose status equals `code` to JDDE_STATUS_OK (e.g. JDDE_STATUS_BUDGET to continue). */
int jdde_clear_status(jdde_handle * h, int32_t code);
/* Upper bound of attempted steps per member and call (default 1<<24); guards against kernels that never end. */
int jdde_set_attempt_budget(jdde_handle * h, int64_t attempts);

/* Page-locked host memory for callers that want asynchronous, full-bandwidth copies of their
 * input/output buffers (bench.py's end-to-end leg); plain malloc'd buffers work everywhere, too. */
int jdde_host_alloc(size_t bytes, void ** out);
int jdde_host_free(void * p);

/* FP64 FMA throughput of the device in FLOP/s, measured with the microbenchmark image `image` (csrc/jdde_peak.cu):
 * the denominator for the FP64 fraction that bench.py reports beside the bandwidth roofline. */
int jdde_measure_fp64_peak(int32_t device, const void * image, size_t len, double * flops);

/* Number of kernels this handle has launched so far (bench.py reports it as gpu_launches). */
int64_t jdde_launch_count(const jdde_handle * h);
/* Device pointer of the last integration result [n_members][n] as an integer address (for zero-copy interop). */
uint64_t jdde_device_result(const jdde_handle * h);

/* ==== one large delay-coupled network, partitioned by node ============================================
 *
 *   dy_i/dt = local(y_i, t, node_par_i) + Σ_{edges e into i} weight_e · coupling(y_{src_e}(t − tau_e), y_i)
 *
 * The structured counterpart of the reference's per-node symbolic expressions
 * (/root/reference/examples/kuramoto_network.py:50-62), which cannot be code-generated for 10^4…10^6 nodes.
 * Same method and controller as above (jitced_template.c:421-498, _jitcdde.py:775-836, 880-933), one scalar state
 * per node, every delay at least as long as a step. With several GPUs each rank owns the nodes [lo, hi), holds the
 * whole history, and after every attempt the ranks all-gather the staged anchor slices and max-reduce the error
 * norm (the caller does this between jdde_net_attempt and jdde_net_control, e.g. with NCCL on the buffers
 * reported by jdde_net_exchange_pointers). */
typedef struct jdde_net jdde_net;

typedef struct jdde_net_desc
{
	int64_t n;               /* nodes of the whole system */
	int64_t lo, hi;          /* nodes [lo, hi) are integrated by this handle */
	int64_t n_edges;         /* edges into the owned nodes */
	int32_t n_node_params;   /* per-node parameters of `local` */
	int32_t ring_capacity;   /* anchors kept, power of two */
	int32_t device;
	int32_t reserved;
} jdde_net_desc;

int jdde_net_create(const jdde_net_desc * desc, jdde_net ** out);
int jdde_net_destroy(jdde_net * h);
const char * jdde_net_last_error(const jdde_net * h);
int jdde_net_load_image(jdde_net * h, const void * image, size_t len);
int jdde_net_set_stream(jdde_net * h, void * cuda_stream);
/* CSR by destination over the owned nodes: row_ptr[hi-lo+1], src/tau/weight[n_edges]; tau > 0. */
int jdde_net_set_graph(jdde_net * h, const int64_t * row_ptr, const int32_t * src, const double * tau, const double * weight);
int jdde_net_set_node_parameters(jdde_net * h, const double * par);          /* [n_node_params][n] */
int jdde_net_set_past(jdde_net * h, int32_t n_anchors, const double * times, const double * states, const double * diffs);   /* states/diffs [n_anchors][n] */
int jdde_net_set_integration_parameters(jdde_net * h, const jdde_int_params * p, double max_delay);
/* pieces of one attempted step, for callers that exchange anchors between ranks in between */
int jdde_net_begin(jdde_net * h, int32_t mode, double target, double step, int64_t steps);   /* mode 0 adaptive, 1 blind */
int jdde_net_attempt(jdde_net * h);
int jdde_net_control(jdde_net * h);
int jdde_net_poll(jdde_net * h, int32_t * done, double * t, int32_t * status);
/* device addresses of the staged anchor ([n][2] doubles: state, diff per node) and of the error norm (IEEE bits of a
 * non-negative double) of the attempt under way */
int jdde_net_exchange_pointers(jdde_net * h, uint64_t * stage, uint64_t * p_bits);

/* Fused peer exchange over NVLink (one process per GPU of one node). jdde_net_exchange_handle writes the CUDA IPC
 * handle (64 bytes) of this rank's exchange buffer; after the ranks have swapped handles (any channel, e.g.
 * torch.distributed.all_gather_object) jdde_net_connect_peers maps the peers' buffers (handles: [world][handle_bytes],
 * this rank's own entry is ignored). From then on jdde_net_attempt stores this rank's part of the tentative anchor
 * directly into every rank's staging buffer from inside the step kernel, merges the error norm with system-scope
 * atomics and ends with a device-side arrival barrier between the GPUs: no collective call between jdde_net_attempt
 * and jdde_net_control any more. All ranks must have connected (host barrier) before any of them integrates. */
int jdde_net_exchange_handle(jdde_net * h, void * handle, size_t handle_bytes);
int jdde_net_connect_peers(jdde_net * h, int32_t rank, int32_t world, const void * handles, size_t handle_bytes);
/* whole calls on a single rank: integrate (_jitcdde.py:807-836), integrate_blindly (:880-933), state [n] */
int jdde_net_integrate(jdde_net * h, double target, double * out_state, int32_t * status);
int jdde_net_integrate_blindly(jdde_net * h, double target, double step, double * out_state, int32_t * status);
/* one entry of the step list of step_on_discontinuities (_jitcdde.py:985-997): adaptive steps to `target`, each attempt capped at
 * the distance to it; otherwise as jdde_net_integrate (status, ring growth and jdde_net_resume included) */
int jdde_net_step_to(jdde_net * h, double target, double * out_state, int32_t * status);
int jdde_net_recent_state(jdde_net * h, double t, int32_t current_only, double * out_state);

/* get_state (_jitcdde.py:357-371): the live anchors in order, for a checkpoint or for switching integrators (they are what
 * jdde_net_set_past takes). *count = their number; with times == NULL nothing else happens; otherwise times [count],
 * states and diffs [count][n] are filled (capacity = anchors the arrays have room for). */
int jdde_net_get_history(jdde_net * h, int32_t * count, double * times, double * states, double * diffs, int32_t capacity);
/* The reference's history is an unbounded list (jitced_template.c:56-73); when the ring is full, a call ends with status
 * JDDE_STATUS_OVERFLOW. jdde_net_grow_ring moves the live anchors into a larger ring (a power of two) and jdde_net_resume
 * continues the call exactly where it stopped (with several GPUs: on every rank; jdde_net_begin(h, 2, …) is the
 * per-attempt form). */
int jdde_net_grow_ring(jdde_net * h, int32_t new_capacity);
int jdde_net_resume(jdde_net * h, double * out_state, int32_t * status);

/* adjust_diff (_jitcdde.py:863-878: a zero-amplitude jump of width shift_ratio·(distance of the last two anchors) at the last
 * anchor, jitced_template.c:1122-1174): the last anchor is replaced by two, the second with the derivative the differential
 * equation gives there. minima/maxima [n] (may be null): extrema of every node over the jump interval. Several ranks: _begin
 * on every rank, then a barrier of the ranks (phase 1 stores into the peers' buffers), then _end on every rank. */
int jdde_net_adjust_diff(jdde_net * h, double shift_ratio, double * minima, double * maxima);
int jdde_net_adjust_diff_begin(jdde_net * h, double shift_ratio);
int jdde_net_adjust_diff_end(jdde_net * h, double * minima, double * maxima);

/* jump (_jitcdde.py:935-974, apply_jump of the C core): amplitude [n] is added to the state over the interval (time, time+width)
 * (forward != 0) or (time−width, time); the time must lie on the last segment of the history — usually it is the time the
 * integration has reached. Several ranks: jdde_net_jump_begin, a barrier of the ranks, jdde_net_adjust_diff_end. */
int jdde_net_jump(jdde_net * h, const double * amplitude, double time, double width, int32_t forward, double * minima, double * maxima);
int jdde_net_jump_begin(jdde_net * h, const double * amplitude, double time, double width, int32_t forward);
int jdde_net_get_counters(jdde_net * h, int64_t * accepted, int64_t * rejected, int64_t * attempts);
int jdde_net_get_dt(jdde_net * h, double * dt);
int64_t jdde_net_launch_count(const jdde_net * h);

#ifdef __cplusplus
}
#endif
#endif

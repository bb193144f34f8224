/*
 * jdde_abi.h — layout shared by the runtime library (jdde_runtime.cu) and the per-DDE
 * kernel image (jdde_kernels.cu). Internal; the public boundary is include/jitcdde_b200.h.
 *
 * Data layout in HBM for an ensemble of n_members trajectories ("members"):
 *
 *   ring   [n_members][cap][A] doubles (+ 16 doubles of skew per member, jdde_ring_stride). One member's anchors are contiguous (member-major):
 *          a trajectory reads and writes its history sequentially in time, so consecutive steps
 *          of a thread hit the same or the next 128-byte line. An anchor is
 *          { time, state[n], diff[n], padding } with A = 1+2n rounded up to a multiple of 4
 *          doubles (32 bytes), so an anchor never straddles a 32-byte sector (n = 1: exactly one
 *          sector). Slot of logical anchor index i = i & (cap-1); first/last/current are logical
 *          indices that only grow (the reference: malloc'd doubly-linked list,
 *          /root/reference/jitcdde/jitced_template.c:25-54).
 *   par    [n_params][n_members]           control parameters (structure of arrays: coalesced)
 *   dt, last_pws, credit, last_start, first, last, current, count, status, cursor[n_sites],
 *   accepted, rejected, attempts, extra_rhs: [n_members] each (structure of arrays), the
 *          per-trajectory controller state that the reference keeps as attributes of its Python
 *          object (/root/reference/jitcdde/_jitcdde.py:710-741) and of dde_integrator
 *          (jitced_template.c:37-54).
 */
#ifndef JDDE_ABI_H
#define JDDE_ABI_H

#include "../../include/jitcdde_b200.h"

#define JDDE_IMAGE_MAGIC 0x4a444445   /* "JDDE" */
#define JDDE_TEAM_FLAG_DOUBLES 64       /* shared memory of jdde_k_orthonormalise_team between the exchange slots and the staged anchors (one flag byte per thread of a block of up to 512) */

typedef struct jdde_problem
{
	double * ring;
	double * par;
	double * dt;
	double * last_pws;
	double * credit;
	double * last_start;
	int * first;
	int * last;
	int * current;
	int * count;
	int * status;
	int * cursor;
	long long * accepted;
	long long * rejected;
	long long * attempts;
	long long * extra_rhs;
	int * resume_index;      /* step_on_discontinuities: index of the step a member stopped in (n_steps: done) */
	double * resume_target;  /* ... and the target time of that step */
	long long n_members;
	int cap;
	int anchor_stride;
	long long ring_stride;   /* doubles between the rings of consecutive members: cap·A plus a skew of one 128-byte
	                            line, so that equal slots of neighbouring members do not fall into the same cache sets */
	double max_delay;
	jdde_int_params P;
} jdde_problem;

/* What follows a step of the restricted / transversal Lyapunov classes (jdde_set_projector; all pointers: device memory) */
typedef struct jdde_projector
{
	int mode;                        /* JDDE_PROJECT_RESTRICTED | JDDE_PROJECT_TRANSVERSAL */
	int n_vectors;                   /* restricted: vectors with more than one non-zero entry … */
	const double * vectors;          /* … [n_vectors][2][n_basic]: state part, derivative part */
	int n_state_components;          /* restricted: unit vectors, handled by zeroing the component */
	const int * state_components;
	int n_diff_components;
	const int * diff_components;
	int n_indices;                   /* transversal: components of the separation function (tangent_indices) */
	const int * indices;
} jdde_projector;

/* `jdde_image_desc` global in every kernel image */
typedef struct jdde_image_desc_t
{
	int magic, abi, n, n_basic, n_sites, n_params, anchor_stride, block, smem_bytes;
	int group_lanes, group_block, group_smem_bytes;   /* jdde_k_integrate_group: lanes per member, block size, dynamic shared memory per block (0 lanes: not in this image) */
} jdde_image_desc_t;

#if defined(__CUDACC__)
#  define JDDE_ABI_FN static __host__ __device__ inline
#else
#  define JDDE_ABI_FN static inline
#endif

JDDE_ABI_FN int jdde_anchor_stride(int n)
{
	return ((1+2*n) + 3) & ~3;
}

JDDE_ABI_FN long long jdde_ring_stride(int cap, int anchor_stride)
{
	return (long long)cap*anchor_stride + 16;
}

#endif

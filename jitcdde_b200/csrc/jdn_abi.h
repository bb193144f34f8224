/*
 * jdn_abi.h — layout of the NETWORK engine: ONE delay-coupled system of n nodes (one scalar state per node),
 * partitioned by node, instead of an ensemble of small systems (jdde_abi.h). Internal; the public boundary
 * is include/jitcdde_b200.h (jdde_net_*).
 *
 *   dy_i/dt = LOCAL(y_i, t, node_par_i) + Σ_{edges e into i} weight_e · COUPLING(y_{src_e}(t − tau_e), y_i)
 *
 * This is the structured, data-driven form of what the reference builds symbolically, one expression per
 * node with literal delays (/root/reference/examples/kuramoto_network.py:50-62) — fine for n = 100, hopeless
 * for a compiler at n = 10^6 (hard part 7).
 *
 * HBM layout:
 *   times  [cap]          anchor times (shared by all nodes)
 *   hist   [cap][n][2]    slot-major, (state, diff) of a node interleaved: the n pairs of one anchor are contiguous →
 *                         the node-parallel reads and writes of a step are coalesced, and a delayed lookup gathers ONE
 *                         16-byte pair per history row (two rows per lookup) instead of four scattered doubles
 *   edges  CSR by destination node: row_ptr[n_owned+1], src[E], tau[E], weight[E], cursor[E]
 *          (cursor = the reference's per-call-site search cursor, jitced_template.c:44-45: one lookup site per edge)
 *   stage  [n][2]         the tentative anchor of the attempt under way. It has a FIXED address, so with several
 *          GPUs the slices computed by the ranks are all-gathered in place without the host knowing the ring index;
 *          a copy kernel commits it to the ring after the controller accepted the step.
 *   ctrl   one jdn_control in device memory: time, step size, indices, error norm, counters — the controller
 *          (/root/reference/jitcdde/_jitcdde.py:775-861) runs on the device in a one-thread kernel between attempts.
 * With several GPUs every rank holds the whole history and the CSR rows of the nodes it owns.
 */
#ifndef JDN_ABI_H
#define JDN_ABI_H

#include "jdde_abi.h"

#define JDN_IMAGE_MAGIC 0x4a444e45   /* "JDNE" */

#define JDN_MODE_ADAPTIVE 0   /* integrate: adaptive steps until t >= target */
#define JDN_MODE_BLIND    1   /* integrate_blindly: `remaining` fixed steps of length h */

typedef struct jdn_control
{
	double t;                 /* time of the current (last accepted) anchor */
	double dt;                /* controller step size (attribute jitcdde.dt) */
	double h;                 /* length of the attempt under way */
	double target;
	double last_start;        /* last_actual_step_start */
	double max_delay;
	unsigned long long p_bits;   /* max_i |err_i|/(atol+rtol|y_i|) of the attempt as IEEE bits (non-negative doubles order like integers) */
	long long accepted, rejected, attempts;
	long long remaining;      /* blind mode: steps still to take */
	int first, current;       /* logical anchor indices; the tentative anchor is current+1 */
	int cap;
	int mode;
	int done;                 /* 1: target reached or failure; attempt kernels return at once */
	int status;               /* JDDE_STATUS_* */
	int pws;                  /* a lookup reached into the step under way (not supported here) */
	int commit_slot;          /* >= 0: the staged anchor was accepted and belongs into this ring slot; -1: nothing to commit */
	jdde_int_params P;
} jdn_control;

typedef struct jdn_problem
{
	jdn_control * ctrl;
	double * times;
	double * hist;            /* [cap][n][2]: (state, diff) pairs */
	double * stage;           /* [n][2] tentative anchor of the attempt under way: written by the attempt kernel (owned nodes),
	                             completed by the all-gather between ranks, copied into the ring when the step is accepted */
	double const * node_par;  /* [n_node_params][n] */
	long long const * row_ptr;/* [n_owned+1], offsets into the edge arrays */
	int const * src;
	double const * tau;
	double const * weight;
	int * cursor;
	long long n;              /* nodes of the whole system */
	long long lo, hi;         /* nodes [lo, hi) are integrated by this rank */
	int cap;
	int n_node_params;
} jdn_problem;

typedef struct jdn_image_desc_t
{
	int magic, abi, n_node_params, block;
} jdn_image_desc_t;

#endif

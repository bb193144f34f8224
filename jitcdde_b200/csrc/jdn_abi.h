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
 *   hist   [cap/2][n][2][2]  (state, diff) of a node interleaved (one 16-byte pair per anchor), and the pairs of ring slots
 *                         2k and 2k+1 of the same node next to each other (jdn_pair_index): a delayed lookup needs the
 *                         pairs of two consecutive anchors of ONE node — half of the time they now share a 32-byte sector,
 *                         and the three stage lookups of an attempt (2–3 consecutive anchors) touch 2 sectors instead of 3.
 *                         Node-parallel reads and writes of a step stride by 32 bytes (every other pair).
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
#define JDN_MODE_CAPPED   3   /* step_on_discontinuities (_jitcdde.py:985-997): adaptive steps, each capped at the distance to the target */
#define JDN_MODE_CONTINUE 2   /* jdde_net_begin only: go on with the call that stopped because the ring was full (after jdde_net_grow_ring) */

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
	int pws;                  /* a lookup reached into the step under way on a path that does not iterate (the two-kernel attempt, warp tiles): the integration stops */
	int commit_slot;          /* >= 0: the staged anchor was accepted and belongs into this ring slot; -1: nothing to commit */
	jdde_int_params P;
	/* peer exchange (several GPUs, jdn_exchange below): number of attempts exchanged so far — never reset; its parity
	 * selects the staging buffer of the attempt under way — and the buffer the last exchange completed */
	unsigned long long epoch;
	int exchanged_buffer;
	int exchange_failed;      /* a peer did not arrive in time */
	/* persistent stepper: tiles beyond a block's first one are handed out in the order in which the blocks ask (blocks on a
	 * slower path to memory take fewer); zero between attempts */
	int next_tile;
	/*
	 * Past within step (a delay shorter than the step: /root/reference/jitcdde/_jitcdde.py:838-861, jitced_template.c:193-217).
	 * `tentative`: an anchor after `current` exists — the result of an attempt that was not accepted (rejected, or one pass of
	 * the iteration below); it is what the reference's list holds as last_anchor != current: lookups beyond `current`
	 * interpolate towards it, the next attempt replaces it. It lives in tentative buffer (epoch & 1) ^ 1, its time in
	 * times[current+1]. `pws_pass` = 0: the attempt under way is the first pass of try_single_step; k > 0: the k-th repetition
	 * with the overlapping lookups (two successive results must agree). pws_bits = max over the lookups of the attempt of
	 * t − t_current as IEEE bits; diff_bits != 0: check_new_y_diff failed for some node. last_pws, count, credit: the
	 * controller's memory (last_pws, count, increase_credit of the reference).
	 */
	int tentative;
	int pws_pass;
	unsigned long long pws_bits;
	unsigned long long diff_bits;
	double last_pws;
	double credit;
	int count;
	int pad_;
} jdn_control;

/*
 * Peer exchange of the tentative anchor between the GPUs that share one network (one process per GPU). Every rank owns
 * one jdn_exchange in device memory that its peers map through CUDA IPC. The step kernel of rank r stores the
 * (state, diff) pairs of the nodes it owns DIRECTLY into stage[b] of every rank (its own included) — 16-byte stores that
 * travel over NVLink while the kernel still computes other nodes — and merges its part of the error norm into every
 * rank's p_bits[b] with a system-scope atomicMax; b = parity of the attempt. A one-thread kernel then tells every peer
 * "my part of attempt k is with you" (release increment of `arrived`) and waits until all peers have told it the same.
 * No collective call, no host involvement, no second pass over the data. Two buffers: a rank that is one attempt ahead
 * writes into the other one, and the arrival counters keep it from getting two ahead.
 */
#define JDN_PWS_SENTINEL 0xffffffffffffffffull   /* in p_bits: a rank saw a delay shorter than the step */
typedef struct jdn_exchange
{
	unsigned long long p_bits[2];
	unsigned long long arrived;
	unsigned long long pws_bits[2];   /* past within step: largest overlap of a lookup with the step, and "two successive results differ" */
	unsigned long long diff_bits[2];
	unsigned long long pad[9];        /* header = 128 bytes; then: double stage[2][n][2] */
} jdn_exchange;

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
	/* peer exchange: world > 1 once jdde_net_connect_peers has mapped the peers' buffers; peers[r] for r = 0…world−1
	 * (device array; peers[rank] is this rank's own buffer) */
	int rank, world;
	jdn_exchange * const * peers;
	/* warp tiles of the persistent stepper: consecutive owned nodes (rows tile_start[k] … tile_start[k+1]−1 of row_ptr) with at
	 * most 32 nodes and at most 32 in-edges together — one lane per edge, one lane per node — or a single node with more
	 * in-edges (handled in rounds); built by jdde_net_set_graph */
	int const * tile_start;
	long long n_warp_tiles;
	/* block tiles of the persistent stepper: consecutive owned nodes per tile (<= JDN_TILE_NODES of the image; 0: that
	 * maximum), chosen by the runtime for the size of the rank's part (jdn_runtime.cu: launch_run) */
	int tile_nodes;
	/* the two tentative buffers [n][2] of ONE GPU (several GPUs: the staging buffers of the exchange serve as such): attempt k
	 * writes the (state, diff) pairs of its result into tentative[k & 1] (k = ctrl->epoch) — and into ring slot current+1 —
	 * and reads the result of attempt k−1, if that is still the last anchor, from the other one. Null: no iteration (the
	 * integration stops when a delay is shorter than the step). */
	double * tentative[2];
	int iterate;              /* 1: this path handles a past within the step as the reference does (persistent stepper with block tiles, host twin) */
	int keep_tentative;       /* 1: a step may be longer than a delay of this rank's in-edges: results are also stored in the tentative buffers */
} jdn_problem;

/* ring slots are grouped by 2^JDN_ROW_GROUP_LOG2 consecutive anchors per node (ring capacities are powers of two >= 4) */
#ifndef JDN_ROW_GROUP_LOG2
#define JDN_ROW_GROUP_LOG2 2
#endif
/* index of the (state, diff) pair of node i in ring slot `slot` (in units of pairs) */
#if defined(__CUDACC__)
static __host__ __device__ inline size_t jdn_pair_index(int slot, long long i, long long n)
#else
static inline size_t jdn_pair_index(int slot, long long i, long long n)
#endif
{
	return (((size_t)(slot >> JDN_ROW_GROUP_LOG2)*(size_t)n + (size_t)i) << JDN_ROW_GROUP_LOG2) + (size_t)(slot & ((1 << JDN_ROW_GROUP_LOG2)-1));
}

#if defined(__CUDACC__)
static __host__ __device__ inline double * jdn_exchange_stage(jdn_exchange * x, int buffer, long long n)
#else
static inline double * jdn_exchange_stage(jdn_exchange * x, int buffer, long long n)
#endif
{
	return (double *)(x+1) + (size_t)buffer*2*(size_t)n;
}

typedef struct jdn_image_desc_t
{
	int magic, abi, n_node_params, block;
	int run_smem_bytes;      /* jdn_k_run: dynamic shared memory per block in addition to the 8·capacity bytes of anchor times */
} jdn_image_desc_t;

#endif

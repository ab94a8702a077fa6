/*
 * chromevol_b200.h -- C ABI of the B200 (sm_100a) chromosome-number likelihood path.
 *
 * The reference (Biols9527/ChromEvol-Enhanced) is a single pure-Python file,
 * src/ancestral_reconstruction.py ("AR" below); it has no FFI of its own.  The entry points
 * here are what a ctypes binding for its hot path binds: each one
 * replaces the numpy/scipy body of one reference method and is called by the Python classes
 * in chromevol_enhanced_b200/ that keep the reference's signatures.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer (torch.Tensor.data_ptr()) unless marked "host";
 *   - `stream` is a cudaStream_t passed as void* (0 = default stream); calls are asynchronous
 *     on that stream unless stated otherwise;
 *   - return value: 0 = ok, negative = error (see cevb_last_error); nothing throws;
 *   - all floating point is IEEE binary64; matrices are row-major;
 *   - n = max_chromosome_number + 1 states (0..N), the reference's matrix order (AR:84).
 *
 * Layouts
 *   params  [B][9]            gain, loss, dupl, demidupl_gain, demidupl_loss,
 *                             base_number_gain, base_number_loss, linear_rate  (AR:22 order)
 *                             then base_number as a double (0 = None)          (AR:44)
 *   pw      [B][7][n][ldq]    slot 0 = Q, slot k = Q^(2k), k = 1..6; ldq = cevb_ldq(n)
 *   norms   [B][16]           see CEVB_NORM_* below
 *   csc     [B][2][CEVB_MAX_SPARSE][n] uint16: slot 0 = row indices of the non-zeros of each Q
 *                             column, slot 1 = column indices of the non-zeros of each Q row,
 *   csc_cnt [B][2][n]         uint8 counts per column / per row (255 = too dense)
 *   t       [E]               branch lengths (>= 0; host applies AR:143-145 / AR:224-227)
 *   P       [B][E][n][ldp]    post-processed transition matrices, ldp = cevb_ldp(n)
 *   info    [B][E]            int32: Pade order m | s << 8 | flags << 16 (CEVB_INFO_*)
 */
#ifndef CHROMEVOL_B200_H
#define CHROMEVOL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CEVB_ABI_VERSION 2
#define CEVB_NPARAM 9
#define CEVB_NPOW 7
#define CEVB_NNORM 16
#define CEVB_MAX_SPARSE 12
#define CEVB_NEVENT 8
#define CEVB_MAX_N 288           /* largest supported matrix order (states 0..287) */
#define CEVB_TAYLOR_KC 16       /* chunks of 4 series coefficients kept per vector: degrees 0..63 (16 = the nibbles of the 64-bit slot-count word) */
#define CEVB_TAYLOR_TRUNC 1e-18 /* absolute truncation bound of the series.  Two orders below the
                                   unit roundoff on purpose: likelihoods of improbable data are
                                   sums of entries near the 1e-12 floor, and logL parity at 1e-9
                                   relative there needs ~1e-21 .. 1e-18 absolute on those entries
                                   (1e-16 fails tests/test_gpu_random.py, 1e-17 passes by 1x) */
#define CEVB_FUSED_INIT 1
#define CEVB_FUSED_APPLY 2
#define CEVB_FUSED_COMBINE 4
#define CEVB_FUSED_LEAVES 8
#define CEVB_FUSED_LEAFVEC 16   /* also write node_vec of the leaves (AR:201-214); output only */
#define CEVB_FUSED_ALL 31

/* norms[b][.] */
#define CEVB_NORM_ONE 0      /* ||Q||_1 */
#define CEVB_NORM_D4 1       /* ||Q^4||_1^(1/4)   */
#define CEVB_NORM_D6 2
#define CEVB_NORM_D8 3
#define CEVB_NORM_D10 4
#define CEVB_NORM_ABS7 5     /* || |Q|^7 ||_1, then ^11, ^15, ^19, ^27 */
#define CEVB_NORM_DENSE 10   /* 1.0 if some column of Q has more than CEVB_MAX_SPARSE non-zeros */
#define CEVB_NORM_SHIFT 11   /* mu = max(0, -min_i Q_ii): the series path sums (Q + mu I)      */
#define CEVB_NORM_SHIFTED 12 /* nu = min(||Q + mu I||_1, ||Q + mu I||_inf)                       */
#define CEVB_NORM_GENERATOR 13 /* 1.0 if Q is a finite generator (off-diagonals >= 0) with nu > 0 */
#define CEVB_NORM_CHUNKS 14  /* series chunks (of 4 degrees) the longest branch t_max can need, 1..CEVB_TAYLOR_KC */

/* info flags (bits 16..) */
#define CEVB_INFO_IDENTITY_T0 1   /* t == 0 -> exact identity (AR:146-148)            */
#define CEVB_INFO_NAN_FALLBACK 2  /* NaN after renormalisation -> identity (AR:156-159) */
#define CEVB_INFO_GROWTH 4        /* element growth monitor tripped in the tile-pivoted solve */
#define CEVB_INFO_REPIVOTED 8     /* matrix was redone with full partial pivoting      */

/* ---- probes -------------------------------------------------------------------------- */
int cevb_abi_version(void);
/* host out-params; returns 0 when a CUDA device with compute capability 10.x is current */
int cevb_device_info(int* sm_count, int* cc_major, int* cc_minor, size_t* smem_optin);
const char* cevb_last_error(void);
int cevb_ldq(int n);   /* row stride (doubles) of Q / power matrices          */
int cevb_ldp(int n);   /* row stride (doubles) of P and of node vectors       */
/* bytes of per-CTA global scratch cevb_expm_branches needs (0 when the working set fits in
 * shared memory), and how many CTAs it launches */
size_t cevb_expm_scratch_bytes(int n);
int cevb_expm_grid(int n, long long n_matrices);

/* ---- (1) rate matrix: replaces ChromEvolutionModel.build_transition_rate_matrix, AR:75-134.
 * Writes slot 0 of pw for every parameter vector, bit-identical to the reference's Q. */
int cevb_build_q(const double* params, int B, int n, double* pw, void* stream);

/* ---- (2a) per-parameter-vector tables for expm: even powers Q^2..Q^12 (DMMA GEMMs), the
 * 1-norms Algorithm 5.1 of Al-Mohy & Higham 2009 needs (what scipy.linalg.expm, AR:151,
 * evaluates per call), and the sparse column structure of Q. */
int cevb_expm_prepare(double* pw, int B, int n, double* norms, uint16_t* csc, uint8_t* csc_cnt,
                      void* stream);

/* ---- (2b) transition matrices: replaces ChromEvolutionModel.get_transition_probabilities,
 * AR:136-166, for every (parameter vector, branch) pair: P = expm(Q t) by variable-order
 * Pade + scaling and squaring, floor 1e-12, row renormalisation, NaN -> identity, t == 0 ->
 * identity.  `scratch` must hold cevb_expm_grid() * cevb_expm_scratch_bytes() bytes (may be
 * NULL when that is 0).  `counter` is one zero-initialised uint32 used as a work queue head. */
int cevb_expm_branches(const double* pw, const double* norms, const uint16_t* csc,
                       const uint8_t* csc_cnt, const double* t, int B, int E, int n, double* P,
                       int32_t* info, void* scratch, unsigned int* counter, int full_pivot,
                       void* stream);

/* Same, for a compact work list built by cevb_expm_plan: queue position q computes matrix
 * worklist[q] (= b * E + e) into P slot q, for q < min(*work_count, work_cap). */
int cevb_expm_branches_list(const double* pw, const double* norms, const uint16_t* csc,
                            const uint8_t* csc_cnt, const double* t, int B, int E, int n, double* P,
                            int32_t* info, void* scratch, unsigned int* counter, int full_pivot,
                            const int32_t* worklist, const unsigned int* work_count,
                            long long work_cap, void* stream);

/* ---- (2c) fused path tables.  All branches of a parameter vector share Q (AR:230-233 with no
 * branch_params), so with the shift mu = max(0, -min_i Q_ii) (uniformisation: Q + mu I is
 * non-negative for a generator, so the series has no cancellation)
 *     expm(Q t) = e^{-mu t} sum_k t^k C_k,   C_k = (Q + mu I)^k / k!
 * with C_k formed once per vector.  cevb_taylor_prepare needs only slot 0 of pw (Q); t_max > 0
 * promises that no branch is longer than t_max, so only the chunks such a branch can need are
 * formed (<= 0: all of them).  It writes
 *   norms[b][CEVB_NORM_ONE / _SHIFT / _SHIFTED / _GENERATOR / _CHUNKS] = ||Q_b||_1, mu_b, nu_b,
 *   "Q_b is a finite generator", chunks the longest branch needs,
 *   the compressed-column structure csc / csc_cnt and
 *   tc: cevb_taylor_doubles(n) doubles per vector -- an opaque table for the other cevb_taylor_*
 *   calls: per (row, pass of 13 column tiles) CEVB_TAYLOR_KC chunks of [13 slots][8 columns]
 *   [4 degrees] (chunk c = degrees 4c..4c+3, tiles in the order in which they first hold a
 *   non-zero, only that slot prefix is written), four 64-bit meta words per (row, pass), and for
 *   generator vectors with n <= 104 a scaled TF32 copy for the observed-leaf launch.
 * The number of terms for a pair is the smallest K + 1 with
 *     x^(K+1)/(K+1)! / (1 - x/(K+2)) e^{-mu t} <= CEVB_TAYLOR_TRUNC,  x = nu t;
 * cevb_expm_plan lists every (vector, branch length) pair for which 4 * CEVB_TAYLOR_KC = 64 terms do not
 * reach that (nu t above about 15.6), which needs scaling and squaring and therefore a materialised
 * matrix: slot [B][E] = position in worklist (-1 = fused path, -2 = list overflow),
 * work_s [cap] = squarings of that position (may be NULL), *count = pairs found (may exceed
 * cap; the caller then re-plans).  It also records the decision per pair for the kernels:
 * plan_nc [B][E] = series terms | kind << 8 (kind 0 fused, 1 identity, 2 materialised),
 * plan_tau [B][E] = t 2^-s, plan_a0 [B][E] = e^{-mu t 2^-s}.
 * cevb_taylor_matrices fills the P slots of a plan: the series at t 2^-s for all listed pairs
 * of a vector at once (tensor cores, t_order = branch-length indices sorted by decreasing t),
 * then s squarings and the AR:153-159 post-processing per matrix in shared memory.
 * counter: one uint32 scratch; nan_count (may be NULL) += matrices that fell back to I;
 * scratch: cevb_taylor_scratch_doubles(n) doubles (may be NULL when that is 0). */
size_t cevb_taylor_doubles(int n);
/* 1 for every supported order.  Up to n = 112 (B200) the squarings run with both matrices in
 * shared memory; larger orders square through global memory / L2 and need
 * cevb_taylor_scratch_doubles(n) doubles of scratch (0 for the shared-memory case). */
int cevb_taylor_matrices_supported(int n);
size_t cevb_taylor_scratch_doubles(int n);
int cevb_taylor_prepare(const double* pw, int B, int n, double t_max, double* norms, uint16_t* csc,
                        uint8_t* csc_cnt, double* tc, void* stream);
int cevb_expm_plan(const double* norms, const double* t, int B, int E, long long cap, int32_t* slot,
                   int32_t* worklist, int32_t* work_s, unsigned int* count, int32_t* plan_nc,
                   double* plan_tau, double* plan_a0, void* stream);
int cevb_taylor_matrices(const double* tc, const double* norms, const int32_t* t_order, int B, int E,
                         int n, const int32_t* slot, const int32_t* work_s,
                         const unsigned int* count, long long cap, const int32_t* plan_nc,
                         const double* plan_tau, const double* plan_a0, double* P,
                         unsigned int* counter, int32_t* nan_count, double* scratch, void* stream);

/* ---- (2d+3) fused transition-probability x child-vector product, AR:136-166 + AR:237 without
 * materialising P: for every listed branch (items[k] = child node id, branch length
 * t_node[child]) and vector b,
 *   w[b][child][i] / wr[b][child][i] = sum_j P'[i][j] L_child[j],  P' = post(expm(Q_b t)),
 * (w = the sum over the floored entries, wr = their row sum, AR:153-154; a non-finite wr[i]
 * means the reference would have found NaN and used the identity, AR:156-159)
 * evaluated on the FP64 tensor cores from tc.  leaf_only = 1 promises that every listed child
 * is an observed leaf (leaf_code >= 0) and selects the variant that stages no child vectors.  L_child is the one-hot / uniform leaf vector
 * (leaf_code >= 0 / == -1) or node_vec[b][child] (leaf_code == -2).  mode [B][n_nodes] must be
 * zero on entry; the kernel sets 1 where P is the identity (t == 0, NaN fallback) and 2 where
 * the pair is on the materialised path -- w is not written for those. */
int cevb_taylor_apply(const double* tc, const double* norms, int B, int n, int n_nodes,
                      const int32_t* items, int n_items, int leaf_only, const double* t_node,
                      const int32_t* leaf_code, const double* node_vec, double* w, double* wr,
                      uint8_t* mode, void* stream);

/* Whole tree with the fused path: same contract and tree arrays as cevb_prune, plus
 *   leaf_child [n_leaf_child] every non-root node that is an observed leaf (leaf_code >= 0):
 *                             their branches need no child vector and go in one launch,
 *   level_child [..]          the remaining child nodes (internal, missing data) grouped by the
 *                             level of their parent (any order inside a level; longest
 *                             branches first packs the work best),
 *   level_child_off [n_levels+1] HOST offsets into level_child,
 *   plan_nc / plan_tau / plan_a0 [B][E] the per-pair plan written by cevb_expm_plan (row
 *                             branch_of[child] belongs to the branch above a child),
 *   P / slot                  materialised matrices and slot map from cevb_expm_plan +
 *                             cevb_expm_branches_list (E = row length of slot),
 *   w, wr [B][n_nodes][ldp], mode [B][n_nodes]   scratch.
 * Levels [level_begin, level_end) are processed; `stages` is a mask of CEVB_FUSED_INIT (zero
 * mode / rescales), CEVB_FUSED_LEAFVEC (write the leaf rows of node_vec; needed only when the
 * caller reads node_vec, not for logL), CEVB_FUSED_LEAVES (the leaf_child launch),
 * CEVB_FUSED_APPLY and CEVB_FUSED_COMBINE.  The whole tree in one call is (0, n_levels,
 * CEVB_FUSED_ALL); the split form lets a caller time the kernels separately. */
int cevb_prune_fused(const double* tc, const double* norms, const double* P, const int32_t* slot,
                     int B, int E, int n, int n_nodes, int root, const int32_t* level_nodes,
                     const int32_t* level_off_host, int n_levels, const int32_t* child_off,
                     const int32_t* child_idx, const int32_t* branch_of, const int32_t* leaf_code,
                     const int32_t* level_child, const int32_t* level_child_off_host,
                     const int32_t* leaf_child, int n_leaf_child, const int32_t* plan_nc,
                     const double* plan_tau, const double* plan_a0, const double* root_freq,
                     double* node_vec, double* w, double* wr,
                     uint8_t* mode, double* logl, int32_t* rescales, int level_begin, int level_end,
                     int stages, void* stream);

/* ---- (3) Felsenstein pruning: replaces ChromEvolLikelihoodCalculator.
 * calculate_likelihood_at_node / calculate_total_log_likelihood, AR:189-272.
 * Tree arrays (int32, device): nodes are numbered 0..n_nodes-1;
 *   level_nodes [n_internal]       internal nodes grouped by level (children before parents),
 *   level_off   [n_levels+1] HOST  offsets into level_nodes,
 *   child_off   [n_nodes+1]        CSR offsets into child_idx (file order, AR:221),
 *   child_idx   [n_branches]       child node of every branch,
 *   branch_of   [n_nodes]          index into the E axis of P for the branch above a node,
 *   leaf_code   [n_nodes]          >= 0 observed state (already clamped, AR:204-208),
 *                                  -1 missing data (uniform, AR:209-211), -2 internal node.
 * node_vec [B][n_nodes][ldp] receives every conditional-likelihood vector, logl [B] the tree
 * log-likelihood (-inf when L <= 0), rescales [B] how many times the AR:242 rescale fired. */
int cevb_prune(const double* P, int B, int E, int n, int n_nodes, int root,
               const int32_t* level_nodes, const int32_t* level_off_host, int n_levels,
               const int32_t* child_off, const int32_t* child_idx, const int32_t* branch_of,
               const int32_t* leaf_code, const double* root_freq, double* node_vec,
               double* logl, int32_t* rescales, void* stream);

/* ---- (4) ancestral states: replaces ancestral_state_reconstruction, AR:298-320.
 * probs may be NULL. */
int cevb_argmax_states(const double* node_vec, int B, int n_nodes, int n, int32_t* ml_count,
                       double* ml_prob, double* probs, void* stream);

/* ---- (4b) marginal and joint reconstruction -- ADDITIVE: the reference has neither (its
 * ancestral_state_reconstruction, AR:274-322, documents the missing down-pass and takes the arg
 * max of the up-pass vector = cevb_argmax_states).  BASELINE.json's north star names both; the
 * CPU restatement is oracle/reconstruction_oracle.py, pinned by exhaustive enumeration over the
 * reference's transition matrices (AR:136-166).  Both read materialised P [B][E][n][ldp] and the
 * tree arrays of cevb_prune, plus
 *   parent    [n_nodes]          parent of every node (-1 for the root),
 *   depth_off [n_depths+1]       nodes are numbered in level order (root 0), so the nodes of
 *                                depth d are the id range [depth_off[d], depth_off[d+1]).
 * cevb_marginal_down: node_vec = the up-pass vectors (cevb_prune / cevb_prune_fused); down
 * [B][n_nodes][ldp] is scratch, post [B][n_nodes][ldp] receives P(state of node | all data).
 * depth_off is a HOST array here (one launch per depth). */
int cevb_marginal_down(const double* P, int B, int E, int n, int n_nodes, const int32_t* parent,
                       const int32_t* child_off, const int32_t* child_idx,
                       const int32_t* branch_of, const int32_t* leaf_code,
                       const int32_t* depth_off_host, int n_depths, const double* root_freq,
                       const double* node_vec, double* down, double* post, void* stream);
/* cevb_joint_reconstruct (Pupko et al. 2000): states [B][n_nodes] of the single most probable
 * assignment, joint_logp [B] = log P(assignment, data).  Scratch: lj [B][n_nodes][ldp],
 * cj [B][n_nodes][n] int32, expo [B][n_nodes] int64.  depth_off is a DEVICE array here. */
int cevb_joint_reconstruct(const double* P, int B, int E, int n, int n_nodes, int root,
                           const int32_t* parent, const int32_t* level_nodes,
                           const int32_t* level_off_host, int n_levels, const int32_t* child_off,
                           const int32_t* child_idx, const int32_t* branch_of,
                           const int32_t* leaf_code, const int32_t* depth_off, int n_depths,
                           const double* root_freq, double* lj, int32_t* cj, long long* expo,
                           int32_t* states, double* joint_logp, void* stream);

/* ---- (5) forward stochastic mapping: replaces StochasticMapper.perform_stochastic_mapping,
 * AR:523-576, with a counter-based Philox4x32-10 generator keyed by (seed, replicate, branch)
 * so any sharding of replicates gives the same histograms.
 *   q [n][ldq] one rate matrix; root_probs [n]; parent [n_nodes] (-1 for root);
 *   preorder [n_nodes] node ids in pre-order (root first); branch_len [n_nodes];
 *   hist [n_nodes][8][n_bins] uint32 += number of replicates with that many events (bin 0 is
 *   NOT written: it is n_reps minus the rest); sums [n_nodes][8][2] uint64 += (sum, sum sq).
 *   state_scratch: n_nodes * rep_chunk uint16.  base_number: 0 = None. */
int cevb_stoch_map(const double* q, int n, const double* root_probs, int n_nodes,
                   const int32_t* parent, const int32_t* preorder, const double* branch_len,
                   int base_number, unsigned long long seed, long long rep_begin,
                   long long rep_end, int n_bins, uint32_t* hist, unsigned long long* sums,
                   uint16_t* state_scratch, long long rep_chunk, unsigned long long* n_events,
                   void* stream);

/* Branch-specific parameter sets (set_branch_parameters, AR:67-73; used by the mapper at AR:569):
 * q [n_sets][...] rate matrices q_stride doubles apart, node_set [n_nodes] the set of the branch
 * above each node, base_numbers [n_sets] (0 = None).  Everything else as cevb_stoch_map; the
 * compact tables of all sets must fit in shared memory (about 15 sets at n = 101). */
int cevb_stoch_map_sets(const double* q, int n_sets, long long q_stride, const int32_t* node_set,
                        const int32_t* base_numbers, int n, const double* root_probs, int n_nodes,
                        const int32_t* parent, const int32_t* preorder, const double* branch_len,
                        unsigned long long seed, long long rep_begin, long long rep_end, int n_bins,
                        uint32_t* hist, unsigned long long* sums, uint16_t* state_scratch,
                        long long rep_chunk, unsigned long long* n_events, void* stream);

/* One branch, one replicate, with every event recorded: replaces
 * StochasticMapper.simulate_events_on_branch, AR:467-521.  ev_* hold max_events entries;
 * out[0] = number of events simulated (may exceed max_events), out[1] = end state. */
int cevb_stoch_branch_trace(const double* q, int n, int start_state, double t, int base_number,
                            unsigned long long seed, unsigned long long replicate, int max_events,
                            double* ev_time, int32_t* ev_from, int32_t* ev_to, int32_t* ev_type,
                            int32_t* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CHROMEVOL_B200_H */

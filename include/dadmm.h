/* dadmm.h -- C ABI of libdadmm_sm100.so: the B200 (sm_100a) hot path of the unfolded D-ADMM
 * distributed-LASSO solver.
 *
 * Each entry point replaces a piece of the reference's Python hot loop (paths relative to the
 * reference checkout; the reference has no FFI of its own -- its boundary is the nn.Module
 * surface, mirrored by the Python modules next to this library):
 *
 *   dadmm_contract      <- unfolded_DLASSO.py:69-71 (AtAy[:,p] = AtA[0,p] @ y[:,p]),
 *                          unfolded_DLASSO.py:120-124 (compute_Atx), and the matmul backward
 *                          autograd derives from them
 *   dadmm_step_fwd      <- unfolded_DLASSO.py:73-99 (grad assembly, clamps, primal update,
 *                          compute_delta :127-140, dual update); model #3 variant
 *                          gnn_dlasso_models_progressive.py:205-232
 *   dadmm_step_bwd      <- the autograd backward of the same lines
 *   dadmm_reduce_hyp    <- the sum-to-size reductions autograd performs for the broadcast
 *                          hyper-parameters alpha,tau,rho,eta (unfolded_DLASSO.py:64-67)
 *   dadmm_unfolded_fwd  <- the whole `for k in range(K)` loop, unfolded_DLASSO.py:53-107
 *   dadmm_unfolded_bwd  <- loss.backward() through that loop (unfolded_train_new.py:79)
 *   dadmm_loss_fwd/bwd  <- gnn_dlasso_utils.py:27-88 (compute_loss) and its backward
 *   dadmm_gcn_epilogue_fwd/bwd <- gnn_dlasso_models_progressive.py:37-72 (the per-sample GCN layer loop of the
 *                          hypernetwork: adjacency mix, bias, LeakyReLU, per-graph BatchNorm, dropout mask)
 *
 * Conventions: all data pointers are DEVICE pointers to contiguous arrays of `dtype`
 * (DADMM_F32 / DADMM_F64) unless a stride is given; state tensors are [B,P,n] with n fastest
 * (the reference's [B,P,n,1]); the caller owns every buffer (workspaces included: nothing is
 * allocated here); calls are asynchronous on `stream` (a cudaStream_t); no global mutable
 * state apart from the thread-local error string, the launch counter / profiling switch and
 * the programmatic-launch switch (dadmm_set_pdl).  Return value: 0 success, <0 invalid
 * argument, >0 a cudaError_t.  Nothing throws.
 */
#ifndef DADMM_H_
#define DADMM_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DADMM_ABI_VERSION 7

typedef void* dadmm_stream_t; /* cudaStream_t */

enum { DADMM_F32 = 0, DADMM_F64 = 1 };

/* contraction algorithms */
enum {
    DADMM_ALGO_AUTO = 0,      /* tcgen05 when the shape allows it (fp32 only), else SIMT          */
    DADMM_ALGO_SIMT = 1,      /* FP32/FP64 FMA pipe, exact IEEE accumulation in k order           */
    DADMM_ALGO_TC_3XTF32 = 2, /* tcgen05.mma kind::tf32, hi/lo split of both operands (3 MMAs)    */
    DADMM_ALGO_TC_3XF16 = 3,  /* tcgen05.mma kind::f16 on scaled fp16 hi/lo pairs (3 MMAs, 2x tf32 rate) */
    DADMM_ALGO_TC_F16X1 = 4   /* FLAGGED reduced precision (1e-2 class): fp16 operands, one MMA, fp32 accumulate;
                                 never chosen by AUTO */
};

/* non-finite bits OR-ed into `flags` by the step kernel (reference guards,
 * unfolded_DLASSO.py:55-61,84-86,102-104) */
enum { DADMM_FLAG_Y = 1, DADMM_FLAG_U = 2, DADMM_FLAG_GRAD = 4, DADMM_FLAG_YNEXT = 8 };

/* Graphs of one batch: `n_graphs` distinct graphs over P nodes, de-duplicated on the host, plus a
 * per-problem index.  For node q of graph g the "event list" ev_idx[ev_ptr[g*P+q] .. ev_ptr[g*P+q+1])
 * holds neighbour ids in the exact order in which the reference's compute_delta
 * (unfolded_DLASSO.py:132-139) accumulates (y_q - y_e) into delta[q]; every undirected edge appears
 * twice per endpoint, so delta = 2*L*y with the reference's rounding. */
typedef struct dadmm_graph {
    int32_t n_graphs;
    int32_t P;
    const int32_t* ev_ptr;   /* device, [n_graphs*P + 1]                                   */
    const int32_t* ev_idx;   /* device, [ev_ptr[n_graphs*P]]                               */
    const int32_t* deg;      /* device, [n_graphs*P]: len(list(graph.neighbors(p)))        */
    const int32_t* graph_id; /* device, [B], or NULL when every problem uses graph 0       */
    /* plain neighbour lists (each neighbour once, self-loops dropped): 2L x = 2(|adj| x_q - sum_j x_j); used
     * where no bit pattern has to be reproduced (backward pass) */
    const int32_t* adj_ptr;  /* device, [n_graphs*P + 1]                                   */
    const int32_t* adj_idx;  /* device, [adj_ptr[n_graphs*P]]                              */
    int32_t max_events;      /* host: largest total event count of one graph (sizes shared-memory staging) */
    int32_t max_adj;         /* host: largest total neighbour count of one graph           */
} dadmm_graph;

/* element-wise clamp bounds of one iteration (unfolded_DLASSO.py:80-81,92-93,99;
 * gnn_dlasso_models_progressive.py:212-232).  D = +inf disables the delta clamp. */
typedef struct dadmm_clamps {
    double G;  /* |grad|  */
    double V;  /* |y|     */
    double D;  /* |delta| */
    double Uc; /* |U|     */
} dadmm_clamps;

/* strided view of the hyper-parameters (alpha,tau,rho,eta) = c 0..3 of agent p of problem b:
 * ptr[b*stride_b + p*stride_p + c*stride_c].  Model #1 table row [P,4]: (0,4,1);
 * model #3 per-sample [B,4,P]: (4P,1,P). */
typedef struct dadmm_hyp {
    const void* ptr;
    int64_t stride_b, stride_p, stride_c;
} dadmm_hyp;

/* Optional low-rank factorisation of the contraction operator of the fused K-iteration path:
 *   W_p = F2_p * F1_p,   F1 [P,m,n],  F2 [P,n,m]   (contiguous, same dtype as the state)
 * For the reference's operator AtA = A^T A (unfolded_DLASSO.py:16; rank A_p = m < n) this is F1 = A[0], F2 = A[0]^T,
 * and W x = F2 (F1 x) costs 4mn instead of 2n^2 flops per agent (F = P*min(2n^2, 4mn)).  The
 * library takes the two-stage route only where it pays (tensor-core path, 8m <= 3n, tile-able m); otherwise the
 * factor is ignored and W is used.  dadmm_unfolded_uses_factor() tells which. */
typedef struct dadmm_factor {
    int32_t m;
    const void* F1;
    const void* F2;
    /* forward only, optional: rhs [B,P,m] with Atb = F2 rhs (the observations b, unfolded_DLASSO.py:45).  The
     * two-stage route then forms the residual as F2 (F1 y - rhs) = AtA y - Atb in its first stage and never
     * streams the n-long Atb.  NULL: Atb is subtracted in the second stage. */
    const void* rhs;
} dadmm_factor;

/* Optional side outputs of dadmm_unfolded_fwd for the per-iteration loss (gnn_dlasso_utils.py:54-66), so that
 * compute_loss need not read Y[K,B,P,n] again.  The label is not known to forward(), hence label-free sums:
 *   agent_sum[k][b][i] = sum_p Y[k][b][p][i]        sumsq[k] = sum_{b,p,i} Y[k][b][p][i]^2
 * and losses[k] = (sumsq[k] - 2 <agent_sum[k], label> + P sum label^2) / (P*B_norm*n)   (dadmm_loss_from_sums).
 * valid[k] (host, written during the call) tells for which iterations the sums were produced: the fused tensor-core
 * path on full tiles, k >= 1; the others must be evaluated from Y with dadmm_loss_fwd. */
typedef struct dadmm_loss_sums {
    void* agent_sum;  /* device [K,B,n], dtype of the state */
    double* sumsq;    /* device [K] */
    int32_t* valid;   /* host   [K] */
} dadmm_loss_sums;

int dadmm_abi_version(void);
const char* dadmm_last_error(void);
/* 0 when the current CUDA device is an sm_100 part this library was built for */
int dadmm_device_check(void);
/* number of kernels launched by this library since load (all threads) */
int64_t dadmm_launch_count(void);

/* Programmatic dependent launch of the K-loop's kernel chain (on by default; DADMM_PDL=0 in the environment starts
 * with it off): each kernel of the chain is scheduled while its predecessor drains and waits on the device
 * (griddepcontrol.wait) before touching anything the predecessor produces.  Results are bit-identical either way;
 * the switch exists for A/B timing and tests.  Returns the previous setting. */
int dadmm_set_pdl(int on);
/* Consensus operator delta = 2 L y of the fused forward levels (reference compute_delta, unfolded_DLASSO.py:127-140).
 * exact = 1: accumulated in the reference's own order of `delta[p] += diff; delta[j] -= diff` events -- bit-identical
 * delta; exact = 0 (default): every neighbour difference (y_q - y_j) taken once and the sum doubled -- the same value to
 * a few ulp with half the shared-memory gathers.  The per-iteration entry point dadmm_step_fwd always uses the exact
 * order.  Returns the previous setting; DADMM_EXACT_ORDER=1 in the environment sets the initial value. */
int dadmm_set_consensus_order(int exact);

/* Per-kernel-kind timing for bench.py's roofline: after dadmm_profile_enable(1) every launch is bracketed
 * by CUDA events on its stream; dadmm_profile_read sums elapsed ms / launch counts per kind
 * (0 contract SIMT, 1 contract tcgen05, 2 step fwd, 3 step bwd, 4 reduce_hyp, 5 loss, 6 operand split,
 * 7 first stage of a two-stage tcgen05 contraction; arrays of 8). */
int dadmm_profile_enable(int on);
int dadmm_profile_read(double* ms_by_kind, int64_t* launches_by_kind);

/* out[b,p,i] (+)= sum_k W[p,i,k] * x[b,p,k]  -- batched over agents, all tensors strided:
 *   W  (p,i,k) -> W  + p*w_sp + i*w_si + k*w_sk      (i < n_out, k < n_in)
 *   x  (b,p,k) -> x  + b*x_sb + p*x_sp + k*x_sk
 *   out(b,p,i) -> out+ b*o_sb + p*o_sp + i*o_si
 * ws/ws_bytes: scratch for the tensor-core algorithms (dadmm_contract_ws_bytes: the fp16 hi/lo copies of both
 * operands for DADMM_ALGO_TC_3XF16), may be NULL when that returns 0. */
int dadmm_contract(int dtype, int algo, int B, int P, int n_out, int n_in,
                   const void* W, int64_t w_sp, int64_t w_si, int64_t w_sk,
                   const void* x, int64_t x_sb, int64_t x_sp, int64_t x_sk,
                   void* out, int64_t o_sb, int64_t o_sp, int64_t o_si,
                   int accumulate, void* ws, size_t ws_bytes, dadmm_stream_t stream);
size_t dadmm_contract_ws_bytes(int dtype, int algo, int B, int P, int n_out, int n_in);
/* dadmm_contract with a workspace the caller keeps between calls (ABI 7): the fp16-pair tensor-core routes leave W's
 * scaled hi/lo copy in the first bytes of `ws`; w_prepared != 0 says that copy, made by an earlier call with the SAME W
 * (contents included), P, n_out, n_in and workspace, is still there, and the call skips W's two split passes.  The
 * reference's operator is a constructor-time constant (unfolded_DLASSO.py:12-16, gnn_dlasso_models_progressive.py:
 * 158-162, where model #3 multiplies by it once per iteration).  Routes without operand copies ignore the flag. */
int dadmm_contract_prepared(int dtype, int algo, int B, int P, int n_out, int n_in,
                            const void* W, int64_t w_sp, int64_t w_si, int64_t w_sk,
                            const void* x, int64_t x_sb, int64_t x_sp, int64_t x_sk,
                            void* out, int64_t o_sb, int64_t o_sp, int64_t o_si,
                            int accumulate, void* ws, size_t ws_bytes, int w_prepared, dadmm_stream_t stream);
/* 1 if dadmm_contract(algo) would run on tcgen05 for this call shape (contiguous [P,n,n]/[B,P,n]) */
int dadmm_contract_uses_tensor_cores(int dtype, int algo, int B, int P, int n_out, int n_in);

/* One D-ADMM iteration (all agents, all problems):
 *   d    = delta ? delta : clampD(2L y)
 *   r    = AtAy - Atb + sign(y)*tau + U*deg + d*rho            (left to right, each op rounded)
 *   g    = clamp(r, +-G);  y+ = clamp(y - alpha*g, +-V)
 *   d+   = clampD(2L y+);  U+ = clamp(U + d+*eta, +-Uc)
 * y_next required; U_next, delta_next, grad_raw (r, saved for backward) and flags may be NULL
 * (U_next == delta_next == NULL skips the consensus phase).  In-place (y_next==y, U_next==U) is allowed. */
int dadmm_step_fwd(int dtype, int B, int P, int n, const dadmm_graph* graph, const dadmm_clamps* clamps,
                   const dadmm_hyp* hyp, const void* y, const void* U, const void* delta,
                   const void* AtAy, const void* Atb,
                   void* y_next, void* U_next, void* delta_next, void* grad_raw,
                   int32_t* flags, dadmm_stream_t stream);

/* Backward of dadmm_step_fwd.  Incoming adjoints (each may be NULL = 0): gy_next_a + gy_next_b
 * (+ loss_coef*(y_next - label[b,:]) when label != NULL) for y+, gU_next for U+, gdelta_next for d+.
 * Outputs: gy = adjoint of y through the direct path only, gAtAy = adjoint of AtAy (the caller adds
 * W^T gAtAy to gy with dadmm_contract), gU, gdelta (adjoint of the `delta` input; when the forward
 * recomputed delta from y, feed it to the previous iteration as gdelta_next), and per-tile partial
 * sums of d/d(alpha,tau,rho,eta) in `ghyp_partials` (dadmm_partials_elems() elements), to be
 * finished by dadmm_reduce_hyp.  Outputs may alias the corresponding *_next inputs. */
int dadmm_step_bwd(int dtype, int B, int P, int n, const dadmm_graph* graph, const dadmm_clamps* clamps,
                   const dadmm_hyp* hyp, const void* y, const void* U, const void* delta,
                   const void* grad_raw, const void* y_next,
                   const void* gy_next_a, const void* gy_next_b, const void* gU_next, const void* gdelta_next,
                   const void* label, double loss_coef,
                   void* gy, void* gAtAy, void* gU, void* gdelta, void* ghyp_partials,
                   dadmm_stream_t stream);
size_t dadmm_partials_elems(int dtype, int B, int P, int n);

/* ghyp(b?,p,c) (+)= sum over tiles (and over b when per_sample == 0) of the partials. */
int dadmm_reduce_hyp(int dtype, int B, int P, int n, const void* ghyp_partials, int per_sample,
                     void* ghyp, int64_t stride_b, int64_t stride_p, int64_t stride_c,
                     int accumulate, dadmm_stream_t stream);

/* Optional persistent home of the operator's tensor-core operand copies (the scaled fp16 hi/lo split of W, or of the
 * factor pair).  The operator of the reference is a constructor-time constant (unfolded_DLASSO.py:12-16 computes AtA
 * once), so its split need not be redone by every forward and every reverse sweep: with ready == 0 the call fills
 * `buf` (dadmm_unfolded_op_split_bytes() bytes, 256-byte aligned) instead of its workspace head, with ready == 1 it
 * reuses the contents.  The split depends on (W | factor F1,F2) and on the route dadmm_unfolded_uses_factor() reports;
 * the forward's and the reverse sweep's splits are the same when Wt and factor_t alias W and factor.  NULL (or a NULL
 * buf, or a shape that does not take the fused tensor-core path: 0 bytes) keeps everything in the workspace. */
typedef struct dadmm_op_split {
    void* buf;
    size_t bytes;
    int ready;
} dadmm_op_split;
size_t dadmm_unfolded_op_split_bytes(int dtype, int algo, int B, int P, int n, int m_factor);

/* K iterations of model #1 (hyp [K,P,4] shared over the batch, clamps[K] on the host):
 * Y[k] = y_{k+1}; U_save[k] = U_{k+1} and R_save[k] = r_k are written when non-NULL (training).
 * W = AtA [P,n,n]; factor (may be NULL) = its factorisation.  Atb [B,P,n] may be NULL only when it is never read:
 * factor->rhs given and dadmm_unfolded_uses_factor() == 1 (the residual is then formed as F2 (F1 y - rhs)).
 * ws >= dadmm_unfolded_ws_bytes(). */
int dadmm_unfolded_fwd(int dtype, int algo, int B, int P, int n, int K, const dadmm_graph* graph,
                       const dadmm_clamps* clamps, const void* hyp, const void* W, const dadmm_factor* factor,
                       const void* Atb,
                       const void* y0, const void* U0, const void* d0,
                       void* Y, void* U_save, void* R_save, void* ws, size_t ws_bytes,
                       int32_t* flags, const dadmm_loss_sums* sums /* may be NULL */,
                       const dadmm_op_split* op_split /* may be NULL */, dadmm_stream_t stream);
/* Reverse sweep: gY [K,B,P,n] dense upstream gradient (may be NULL) and/or the fused loss term
 * coef[k]*(Y[k]-label) with label [B,n] and the K coefficients either on the host (loss_coef) or on the
 * device (loss_coef_dev, doubles; used when loss_coef == NULL -- autograd hands d loss / d losses[k] over
 * as a device tensor, and reading it on the host would drain the stream in the middle of every training
 * step).  A zero coefficient means "no term" either way.  Wt = AtA^T [P,n,n]; factor_t (may be NULL)
 * factorises Wt (for the symmetric AtA: the forward's factor).  Writes ghyp [K,P,4]. */
int dadmm_unfolded_bwd(int dtype, int algo, int B, int P, int n, int K, const dadmm_graph* graph,
                       const dadmm_clamps* clamps, const void* hyp, const void* Wt, const dadmm_factor* factor_t,
                       const void* y0, const void* U0, const void* d0,
                       const void* Y, const void* U_save, const void* R_save,
                       const void* gY, const void* label, const double* loss_coef, const double* loss_coef_dev,
                       void* ghyp, void* ws, size_t ws_bytes, const dadmm_op_split* op_split /* may be NULL */,
                       dadmm_stream_t stream);
/* m_factor = factor->m of the call (0: no factor) */
size_t dadmm_unfolded_ws_bytes(int dtype, int algo, int B, int P, int n, int K, int backward, int m_factor);
/* 1 when the fused path would evaluate the contraction in two stages for this shape and inner dimension m */
int dadmm_unfolded_uses_factor(int dtype, int algo, int B, int P, int n, int m);

/* losses[k] = sum_{b,p,i} (Y[k,b,p,i] - label[b,i])^2 / (P*B_norm*n)  (gnn_dlasso_utils.py:54-66;
 * B_norm = global batch when the batch is sharded over ranks).  ws >= dadmm_loss_ws_bytes(). */
int dadmm_loss_fwd(int dtype, int K, int B, int P, int n, int64_t B_norm, const void* Y, const void* label,
                   void* losses, void* ws, size_t ws_bytes, dadmm_stream_t stream);
/* gY[k] = coef[k] * (Y[k] - label)   (coef on the host, [K]; zero rows are memset) */
int dadmm_loss_bwd(int dtype, int K, int B, int P, int n, const void* Y, const void* label,
                   const double* coef, void* gY, dadmm_stream_t stream);
size_t dadmm_loss_ws_bytes(int dtype, int K, int B, int P, int n);
/* losses[k], k0 <= k < k1, from the side outputs of dadmm_unfolded_fwd (see dadmm_loss_sums); same workspace size.
 * The sums give sum (Y - label)^2 as sum Y^2 - 2 <S, label> + P sum label^2; where that difference cancels more than three
 * digits (a near-converged iterate) and Y is given, the iteration is re-evaluated exactly from Y[k] on the device (no host
 * round trip).  Y may be NULL (no re-evaluation). */
int dadmm_loss_from_sums(int dtype, int k0, int k1, int B, int P, int n, int64_t B_norm, const void* agent_sum,
                         const double* sumsq, const void* label, const void* Y, void* losses, void* ws, size_t ws_bytes,
                         dadmm_stream_t stream);

/* Per-problem epilogue of one graph-convolution layer of the model-#3 hypernetwork (fp32 only) -- replaces, for the whole
 * batch at once, what gnn_dlasso_models_progressive.py:37-72 does sample by sample after the dense product H = x W^T:
 *     Z = A_hat_b H_b + bias;  A = leaky_relu(Z, slope);  BatchNorm over the P nodes of problem b (training: batch
 *     statistics of that one graph; otherwise the running statistics);  out = BN(A) * mask
 * H, out, act, mask: [B*P, C] row-major; adj: [B,P,P]; bias, bn_w, bn_b, run_mean, run_var: [C]; mean, var (biased):
 * [B,C], written in training mode only; mask (may be NULL) already carries the 1/(1-p) scale of dropout.  P <= 64.
 * The backward returns gH = d loss / d H and one row of partial sums per CTA, partials[nblk][3][C] = (d bn_w, d bn_b,
 * d bias), nblk = dadmm_gcn_partial_rows(B, C); the caller sums the rows. */
int dadmm_gcn_epilogue_fwd(int B, int P, int C, const void* H, const void* adj, const void* bias, const void* bn_w,
                           const void* bn_b, const void* run_mean, const void* run_var, int training, double eps,
                           double slope, const void* mask, void* out, void* act, void* mean, void* var,
                           dadmm_stream_t stream);
int dadmm_gcn_epilogue_bwd(int B, int P, int C, const void* gout, const void* adj, const void* bn_w, const void* run_mean,
                           const void* run_var, int training, double eps, double slope, const void* mask,
                           const void* act, const void* mean, const void* var, void* gH, void* partials,
                           dadmm_stream_t stream);
int dadmm_gcn_partial_rows(int B, int C);

#ifdef __cplusplus
}
#endif
#endif /* DADMM_H_ */

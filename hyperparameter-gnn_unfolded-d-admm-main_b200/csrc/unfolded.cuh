// unfolded.cuh -- "level" kernels of the fused K-iteration path (model #1: hyper-parameters shared by the
// batch).  Same tile decomposition as step.cuh (CTA = TB problems x all P agents x 32*VEC unknowns, one warp
// per (problem, agent) row segment), but the recurrence is re-associated so that each launch needs ONE
// shared-memory tile of y_k and every state tensor crosses HBM once:
//
//   forward level k   (reads a_k=AtA y_k, Atb, y_k, U_{k-1}; writes y_{k+1}, U_k, r_k)
//       d_k = k ? clampD(2L y_k) : d_0                 U_k = k ? clamp(U_{k-1} + d_k*eta_{k-1}, Uc_{k-1}) : U_0
//       r_k = a_k - Atb + sign(y_k) tau_k + U_k deg + d_k rho_k
//       y_{k+1} = clamp(y_k - alpha_k clamp(r_k, G_k), V_k)
//     -- the dual update of reference iteration k-1 (unfolded_DLASSO.py:95-99) is evaluated at the start of
//        level k, where 2L y_k is needed anyway for the gradient (:73-77); values are bit-identical
//        (2L y_k is accumulated in the reference's event order).
//
//   backward level k  (reads T=adj(y_{k+1}), C=adj(U_k), y_k, U_{k-1}, r_k; writes gAtAy_k, S, adj(U_{k-1}))
//       zb = T 1[|z_k|<=V_k]            rb = -alpha_k zb 1[|r_k|<=G_k]              -> gAtAy_k = rb
//       Uk = C + deg rb                 Um = Uk 1[|U_{k-1} + d_k eta_{k-1}|<=Uc_{k-1}]  -> adj(U_{k-1}) = Um
//       db = (rho_k rb + eta_{k-1} Um) 1[|2L y_k|<=D]
//       S  = zb + gY[k-1] + coef_{k-1}(y_k - label) + 2L db      (the caller adds W^T rb:  adj(y_k) = S + W^T rb)
//       d alpha_k = -<zb, g_k>, d tau_k = <rb, sign y_k>, d rho_k = <rb, d_k>, d eta_{k-1} = <Um, d_k>
//     -- here 2L x is evaluated as 2(deg x_q - sum_j x_j) over the plain neighbour list (half the terms of the
//        event order; the backward pass has no bit pattern to reproduce).
//
// (In the fused fp16 path a_k already holds AtA y_k - Atb -- the contraction's epilogue, or its first stage
// A^T(A y_k - b), subtracts the observation term -- and the saved stream is that residual; the backward level
// rebuilds r_k from it.)
//
// These kernels are instruction-issue and latency sensitive (ncu, round 1: the first version executed ~1100 warp
// instructions per row segment and sat at 22 % of HBM peak), so: the tile's neighbour lists and per-agent scalars are
// staged in shared memory once per CTA, rows are walked without integer division, the y_k tile arrives by cp.async,
// the four hyper-parameter partial sums share one 6-shuffle reduction, and the fused path's configuration on full
// tiles runs LEAN template instantiations without guards, zero-fills or select chains (see level_fwd_kernel /
// level_bwd_kernel).  Register double-buffering of the next row was measured twice and lost to occupancy both times.
#pragma once
#include <cuda_fp16.h>

#include "step.cuh"

namespace dadmm {

// Optional fused operand split for the fp16 tensor-core contraction (contract_f16.cuh): the producer of a GEMM
// operand writes it as scaled fp16 (hi, lo) pairs, so the GEMM needs no conversion pass.  The scale is a power
// of two derived from a RIGOROUS bound of the tensor's max |v| that every CTA computes identically:
//   forward : |y_{k+1}| <= min(V_k, max|y_k| + max_p(alpha_p) * G_k)
//   backward: |gAtAy_k| <= max_p(alpha_p) * max|adj(y_{k+1})|
// (max|y_k| / max|adj| are tracked with one atomicMax per CTA by the kernels that produce those tensors).
struct SplitOut {
    __half *hi, *lo;          // [B*P][n] fp16 each (n % 8 == 0), nullptr = no split output
    int* exp;                 // device: scale exponent written for the GEMM epilogue
    const unsigned* amax_in;  // device: max |.| bits of the tensor the bound is derived from (nullptr = use the clamp bound)
    unsigned* amax_out;       // device: running max |.| bits of the tensor this kernel produces (nullptr = not tracked)
};

__device__ __forceinline__ int split_exponent(float bound) {
    const unsigned bits = __float_as_uint(bound);
    const int ea = (int)((bits >> 23) & 0xFF) - 127;
    if (ea <= -100 || ea >= 128) return 0;
    return 13 - ea;                                   // bound * 2^e in [2^13, 2^14)
}
__device__ __forceinline__ float pow2_of(int e) { return __uint_as_float((unsigned)(e + 127) << 23); }

template <int VEC>
__device__ __forceinline__ void store_split(const SplitOut& sp, unsigned off, const Vec<float, VEC>& v, float s1, float s2) {
    const float sc = s1 * s2;                      // 2^e with |e| <= 115: a normal float, the product is exact
    if constexpr (VEC % 2 == 0) {
        __half2 h[VEC / 2], l[VEC / 2];
#pragma unroll
        for (int j = 0; j < VEC / 2; ++j) {
            const float x0 = v.v[2 * j] * sc, x1 = v.v[2 * j + 1] * sc;
            h[j] = __floats2half2_rn(x0, x1);                       // one packed conversion per pair
            const float2 hf = __half22float2(h[j]);
            l[j] = __floats2half2_rn(x0 - hf.x, x1 - hf.y);
        }
        if constexpr (VEC == 4) {
            *reinterpret_cast<uint2*>(sp.hi + off) = *reinterpret_cast<const uint2*>(h);
            *reinterpret_cast<uint2*>(sp.lo + off) = *reinterpret_cast<const uint2*>(l);
        } else {
            *reinterpret_cast<__half2*>(sp.hi + off) = h[0];
            *reinterpret_cast<__half2*>(sp.lo + off) = l[0];
        }
    } else {
        const float x = v.v[0] * sc;
        const __half h = __float2half_rn(x);
        sp.hi[off] = h;
        sp.lo[off] = __float2half_rn(x - __half2float(h));
    }
}

// max_p |alpha_p| of one table row, computed by the whole CTA (one load per thread, one barrier) -- a serial
// P-long load chain per thread cost ~1500 cycles at every CTA start in the first version
template <typename T>
__device__ __forceinline__ float block_max_alpha(const T* __restrict__ hyp_row, int P, float* sh /*[32]*/) {
    float v = 0.f;
    for (int q = threadIdx.x; q < P; q += blockDim.x) v = fmaxf(v, fabsf((float)__ldg(hyp_row + q * 4)));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    float m = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) m = fmaxf(m, sh[w]);
    return m;
}

// CTA-wide max of per-thread |.| bits -> at most one atomicMax per CTA
__device__ __forceinline__ void publish_amax(unsigned m, unsigned* out, unsigned* sh /*[32]*/) {
    m = __reduce_max_sync(0xffffffffu, m);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) sh[warp] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned t = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t = max(t, sh[w]);
        if (t > *reinterpret_cast<volatile unsigned*>(out)) atomicMax(out, t);
    }
}

template <typename T>
struct LevelFwdParams {
    int B, P, n, TB, first, list_cap;   // list_cap: shared-memory ints reserved per problem for its neighbour list
    int csplit;                         // CTAs per problem group (each walks nchunks/csplit consecutive chunks)
    const int32_t *lst_ptr, *lst_idx, *deg, *gid;   // event lists (exact order), or plain adjacency lists when !exact_order
    int exact_order;                    // 1: 2L y accumulated in the reference's event order (bit-identical delta); 0: see lean::lap_half
    int prefetch;                       // lean kernels: software prefetch of the warp's next row (0 off, 1 into L1, 2 into L2)
    const T *hyp_k, *hyp_prev;          // rows [P,4] of the table
    T G, V, D, Uc_prev;
    int hasD;
    const T *y, *U_in, *d0, *a, *atb;
    T *y_next, *U_out, *graw;
    int32_t* flags;
    SplitOut sp;                        // fp16 split of y_{k+1} (fp32 only)
    T* agent_sum;                       // optional [B][n]: sum over agents of y_{k+1} (lean form, TB == 1)
    double* sq_part;                    // optional [gridDim.x]: per-CTA sum of y_{k+1}^2
};

template <typename T>
struct LevelBwdParams {
    int B, P, n, TB, first, top, list_cap;
    int csplit;                         // CTAs per problem group: each CTA walks nchunks/csplit consecutive 128-unknown chunks
    const int32_t *lst_ptr, *lst_idx, *deg, *gid;   // plain adjacency lists
    const T *hyp_k, *hyp_prev;
    T G, V, D, Uc_prev;
    int hasD;
    const T *y, *U_prev, *d0, *graw;
    int graw_is_residual;               // 1: `graw` holds AtA y - Atb and r_k is rebuilt here (no saved r_k stream)
    int prefetch;                       // lean kernel: software prefetch of the warp's next row (0 off, 1 into L1, 2 into L2)
    T *Tb, *C, *ga;
    const T *gY_prev, *label;
    T coef_prev;
    const double* coef_dev;             // optional, device: coefficient of this level's loss term (replaces coef_prev; 0 = no term)
    T* partials;
    SplitOut sp;                        // fp16 split of gAtAy_k; sp.amax_in = max|adj(y_{k+1})|
};

// Stage the neighbour lists of the tile's problems: sPtr[bl][0..P] = list bounds relative to sIdx[bl],
// sIdx[bl][e] = neighbour row ids.  When the lists do not fit
// (list_cap == 0) the kernels read them from global memory instead.
template <typename T, int VEC>
__device__ __forceinline__ void stage_lists(int32_t* sPtr, int32_t* sIdx, int TB, int P, int B, int b0, int cap,
                                            const int32_t* __restrict__ lst_ptr, const int32_t* __restrict__ lst_idx,
                                            const int32_t* __restrict__ gid) {
    for (int bl = 0; bl < TB; ++bl) {
        const int b = b0 + bl;
        if (b >= B) break;
        const int node0 = (gid ? __ldg(gid + b) : 0) * P;
        const int e0 = __ldg(lst_ptr + node0);
        for (int q = threadIdx.x; q <= P; q += blockDim.x) sPtr[bl * (P + 1) + q] = __ldg(lst_ptr + node0 + q) - e0;
        const int cnt = __ldg(lst_ptr + node0 + P) - e0;
        for (int e = threadIdx.x; e < cnt; e += blockDim.x) sIdx[bl * cap + e] = __ldg(lst_idx + e0 + e);
    }
}

// 2L x for row q in the reference's event order (sequential (x_q - x_e) accumulation)
template <typename T, int VEC>
__device__ __forceinline__ Vec<T, VEC> lap_events(const unsigned char* tile, const Vec<T, VEC>& xq, const int32_t* sIdx,
                                                  int e0, int e1, int lane_bytes) {
    Vec<T, VEC> acc = vzero<T, VEC>();
#pragma unroll 4
    for (int e = e0; e < e1; ++e) {
        const Vec<T, VEC> xj = *reinterpret_cast<const Vec<T, VEC>*>(tile + sIdx[e] * (32 * VEC * (int)sizeof(T)) + lane_bytes);
#pragma unroll
        for (int v = 0; v < VEC; ++v) acc.v[v] = add_rn(acc.v[v], sub_rn(xq.v[v], xj.v[v]));
    }
    return acc;
}

// 2L x = 2 (deg x_q - sum_j x_j) over the plain neighbour list
template <typename T, int VEC>
__device__ __forceinline__ Vec<T, VEC> lap_adj(const unsigned char* tile, const Vec<T, VEC>& xq, const int32_t* sIdx,
                                               int e0, int e1, int lane_bytes) {
    Vec<T, VEC> acc = vzero<T, VEC>();
#pragma unroll 4
    for (int e = e0; e < e1; ++e) {
        const Vec<T, VEC> xj = *reinterpret_cast<const Vec<T, VEC>*>(tile + sIdx[e] * (32 * VEC * (int)sizeof(T)) + lane_bytes);
#pragma unroll
        for (int v = 0; v < VEC; ++v) acc.v[v] += xj.v[v];
    }
    const T dq = (T)(e1 - e0);
#pragma unroll
    for (int v = 0; v < VEC; ++v) acc.v[v] = (T)2 * (dq * xq.v[v] - acc.v[v]);
    return acc;
}

// sums of four per-lane values in 6 shuffles; results land in lanes 0 (a), 8 (b), 16 (c), 24 (d)
template <typename T>
__device__ __forceinline__ T warp_sum4(T a, T b, T c, T d, int lane) {
    const bool hi = lane & 16;
    T k0 = hi ? c : a, k1 = hi ? d : b;
    k0 += __shfl_xor_sync(0xffffffffu, hi ? a : c, 16);
    k1 += __shfl_xor_sync(0xffffffffu, hi ? b : d, 16);
    const bool h8 = lane & 8;
    T k = h8 ? k1 : k0;
    k += __shfl_xor_sync(0xffffffffu, h8 ? k0 : k1, 8);
    k += __shfl_xor_sync(0xffffffffu, k, 4);
    k += __shfl_xor_sync(0xffffffffu, k, 2);
    k += __shfl_xor_sync(0xffffffffu, k, 1);
    return k;
}

#ifndef DADMM_LEVEL_BWD_PREFETCH
#define DADMM_LEVEL_BWD_PREFETCH 0   // register double-buffering of the next row: 2.24 ms with, 1.78 ms without (B200, cfg4)
#endif
#ifndef DADMM_LEVEL_FWD_PREFETCH
#define DADMM_LEVEL_FWD_PREFETCH 0   // same finding as for the backward kernel: occupancy beats register double-buffering
#endif
#ifndef DADMM_LEVEL_MINB_FWD
#define DADMM_LEVEL_MINB_FWD 5      // round-1 sweep on B200 (cfg4, no prefetch): 3 -> 1.51 ms, 4 -> 1.31, 5 -> 1.24 per level
#endif
#ifndef DADMM_LEVEL_MINB_BWD
#define DADMM_LEVEL_MINB_BWD 3      // no prefetch: 3 -> 1.78 ms (80 regs, no spills), 4 -> 2.10 (spills)
#endif
#ifndef DADMM_LEVEL_BWD_LEAN_PREFETCH
#define DADMM_LEVEL_BWD_LEAN_PREFETCH 0
#endif
#ifndef DADMM_LEVEL_MINB_FWD_LEAN
#define DADMM_LEVEL_MINB_FWD_LEAN DADMM_LEVEL_MINB_FWD
#endif
#ifndef DADMM_LEVEL_MINB_BWD_LEAN
#define DADMM_LEVEL_MINB_BWD_LEAN 4  // lean form: 64 registers, ~55 KB shared memory per CTA; 3 -> 1.49 ms, 4 -> 1.45 ms
#endif

// 16-byte asynchronous global -> shared copy (LDGSTS): the tile loads of a CTA are all in flight at once and hold no
// registers.  ncu (round 1): the register-staged form (load, wait, store, row after row) put 45 % of the backward
// level's stall samples on the tile load and the barrier behind it.
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// sign(y) * tau as torch computes it (sign is 0 for 0 and NaN; +-1 * tau is exact): one compare, one bit merge, one select
template <typename T>
__device__ __forceinline__ T sign_times(T y, T tau) {
    if constexpr (sizeof(T) == 4) return (fabsf(y) > 0.f) ? copysignf(tau, y) : 0.f;
    else return (fabs(y) > 0.0) ? copysign(tau, y) : 0.0;
}

// Per-CTA staging of the per-agent scalars every row needs: sHyp[q] = (alpha_k, tau_k, rho_k, eta_{k-1}) and
// sDeg[bl*P + q] = degree of agent q in the tile's problem bl (one LDS.128 + one LDS per row instead of five
// dependent global loads)
template <typename T>
__device__ __forceinline__ void stage_scalars(T* sHyp, T* sDeg, int TB, int P, int B, int b0, const T* __restrict__ hyp_k,
                                              const T* __restrict__ hyp_prev, const int32_t* __restrict__ deg,
                                              const int32_t* __restrict__ gid) {
    for (int q = threadIdx.x; q < P; q += blockDim.x) {
        sHyp[q * 4 + 0] = __ldg(hyp_k + q * 4);
        sHyp[q * 4 + 1] = __ldg(hyp_k + q * 4 + 1);
        sHyp[q * 4 + 2] = __ldg(hyp_k + q * 4 + 2);
        sHyp[q * 4 + 3] = hyp_prev ? __ldg(hyp_prev + q * 4 + 3) : (T)0;
    }
    for (int r = threadIdx.x; r < TB * P; r += blockDim.x) {
        const int b = b0 + r / P;
        sDeg[r] = (b < B) ? (T)__ldg(deg + (gid ? __ldg(gid + b) : 0) * P + (r % P)) : (T)0;
    }
}

// any non-finite entry in one table row -> all flag bits: a non-finite alpha_k makes y_{k+1} non-finite without touching
// y_k, U_k or r_k, the quantities the level kernels watch (reference guard unfolded_DLASSO.py:102-104)
template <typename T>
__device__ __forceinline__ void flag_nonfinite_row(const T* __restrict__ hyp_row, int P, int32_t* flags) {
    if (flags && blockIdx.x == 0)
        for (int q = threadIdx.x; q < 4 * P; q += blockDim.x)
            if (!isfinite(__ldg(hyp_row + q))) atomicOr(flags, 0xF);
}

// LEAN: the configuration of the fused fp16 training/inference path on full tiles -- a already holds AtA y - Atb
// (no Atb stream), r_k is not written, no delta clamp, n % (32 VEC) == 0 and B % TB == 0 (no lane / problem guards).
// ncu (round 1) showed the generic form executing 366 straight-line instructions per 128-unknown row segment, a
// third of them guards, zero-fills and NaN-propagating select clamps; the lean form keeps the arithmetic (same
// operations, same order, same rounding) and drops the rest.  Non-finite inputs are caught through r_k alone: a NaN/Inf
// in y_k reaches AtA y_k, one in U_k reaches U_k deg (0 * Inf = NaN), and any hit re-runs the batch on the guarded path.
// NTHR: threads per CTA.  One warp owns whole rows, so 8 warps split a 50-row tile (P = 50) as 7,7,6,6,6,6,6,6; ten warps
// (320 threads) split it evenly.  Pays in the backward level (-5 %), not here (the host keeps 256 threads for this kernel).
template <typename T, int VEC, bool LEAN, int NTHR = kStepThreads>
__global__ void __launch_bounds__(NTHR, NTHR > kStepThreads ? DADMM_LEVEL_MINB_FWD_LEAN - 1 : (LEAN ? DADMM_LEVEL_MINB_FWD_LEAN : DADMM_LEVEL_MINB_FWD))
level_fwd_kernel(const LevelFwdParams<T> p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int CH = 32 * VEC;
    using V = Vec<T, VEC>;
    const int P = p.P, R = p.TB * P;
    unsigned char* S0 = smem_raw;                                              // y_k tile [R][CH]
    T* sHyp = reinterpret_cast<T*>(S0 + (p.first ? (size_t)0 : (size_t)R * CH * sizeof(T)));
    T* sDeg = sHyp + 4 * P;
    int32_t* sPtr = reinterpret_cast<int32_t*>(sDeg + R);
    int32_t* sIdx = sPtr + p.TB * (P + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const int lane_bytes = lane * VEC * (int)sizeof(T);
    const int nchunks = (p.n + CH - 1) / CH;
    const int cs = blockIdx.x % p.csplit;               // a CTA walks a contiguous range of chunks of its problems
    const int b0 = (blockIdx.x / p.csplit) * p.TB;
    const int cpc = (nchunks + p.csplit - 1) / p.csplit;
    const int chunk_begin = cs * cpc, chunk_end = min(nchunks, chunk_begin + cpc);
    const bool first = p.first != 0;
    __shared__ unsigned sAmax[32];
    __shared__ float sAlpha[32];

    // call-constant inputs (neighbour lists, table rows, degrees) are staged before the dependency wait: this part of the
    // CTA's life overlaps the tail of the preceding contraction (common.cuh, programmatic dependent launch)
    const bool staged = p.list_cap > 0;
    if (!first && staged) stage_lists<T, VEC>(sPtr, sIdx, p.TB, P, p.B, b0, p.list_cap, p.lst_ptr, p.lst_idx, p.gid);
    stage_scalars<T>(sHyp, sDeg, p.TB, P, p.B, b0, p.hyp_k, p.hyp_prev, p.deg, p.gid);
    flag_nonfinite_row<T>(p.hyp_k, P, p.flags);
    pdl_wait();
    pdl_trigger();

    // scale of the fused fp16 split of y_{k+1} (identical in every CTA)
    float sc = 1.f;
    bool do_split = false;
    if constexpr (sizeof(T) == 4) {
        do_split = p.sp.hi != nullptr;
        if (do_split) {
            float bound = (float)p.V;
            if (p.sp.amax_in) {
                const float amax_alpha = block_max_alpha(p.hyp_k, P, sAlpha);
                bound = fminf(bound, __uint_as_float(__ldg(p.sp.amax_in)) + amax_alpha * (float)p.G);
            }
            const int e = split_exponent(bound);
            sc = pow2_of(e / 2) * pow2_of(e - e / 2);    // |e| <= 115: a normal float, the product is exact
            if (blockIdx.x == 0 && threadIdx.x == 0) *p.sp.exp = e;
        }
    }
    unsigned amax_bits = 0;
    float amax_f = 0.f;

    if (first) __syncthreads();                          // (the tile-load barrier below covers the other levels)
    T nonfinite = (T)0;
    const T nG = -p.G, nV = -p.V;
    const bool sums = LEAN && p.agent_sum != nullptr;
    float sq_tot = 0.f;

    for (int chunk = chunk_begin; chunk < chunk_end; ++chunk) {
    const int i = chunk * CH + lane * VEC;
    const bool act_i = LEAN || i < p.n;
    V vsum = vzero<T, VEC>();
    if (!first) {
        if (chunk != chunk_begin) __syncthreads();      // rows of the previous chunk are still being read
        for (int bl = 0; bl < p.TB; ++bl) {
            const int b = b0 + bl;
            const T* src = p.y + (((unsigned)b * P) * p.n + i);
            for (int pp = warp; pp < P; pp += nwarps) {
                unsigned char* dst = S0 + (unsigned)(bl * P + pp) * (CH * (unsigned)sizeof(T)) + lane_bytes;
                if constexpr (LEAN && sizeof(V) == 16) {
                    cp_async16(dst, src + (unsigned)pp * p.n);
                } else {
                    V v = vzero<T, VEC>();
                    if (act_i && b < p.B) v = ld_vec<T, VEC>(src + (unsigned)pp * p.n);
                    *reinterpret_cast<V*>(dst) = v;
                }
            }
        }
        if constexpr (LEAN && sizeof(V) == 16) cp_async_wait_all();
        __syncthreads();
    }

    for (int bl = 0; bl < p.TB; ++bl) {
        const int b = b0 + bl;
        if (!LEAN && b >= p.B) break;
        const unsigned base = ((unsigned)b * P) * p.n + i;          // 32-bit element offsets: B*P*n < 2^31 (host-checked)
        const unsigned char* tile = S0 + (size_t)bl * P * CH * sizeof(T);
        const int32_t* lptr = staged ? sPtr + bl * (P + 1) : p.lst_ptr + (p.gid ? __ldg(p.gid + b) : 0) * P;
        const int32_t* lidx = staged ? sIdx + bl * p.list_cap : p.lst_idx;
        for (int pp = warp; pp < P; pp += nwarps) {
            const unsigned off = base + (unsigned)pp * p.n;
            V av, atbv, Uv, yv, dv;
            if constexpr (LEAN) {
                av = ld_stream<T, VEC>(p.a + off);
                Uv = ld_stream<T, VEC>(p.U_in + off);
                if (first) {
                    yv = ld_vec<T, VEC>(p.y + off);
                    dv = ld_stream<T, VEC>(p.d0 + off);
                }
            } else {
                av = atbv = Uv = yv = dv = vzero<T, VEC>();
                if (act_i) {
                    av = ld_stream<T, VEC>(p.a + off);
                    if (p.atb) atbv = ld_stream<T, VEC>(p.atb + off);
                    Uv = ld_stream<T, VEC>(p.U_in + off);
                    if (first) {
                        yv = ld_vec<T, VEC>(p.y + off);
                        dv = ld_stream<T, VEC>(p.d0 + off);
                    }
                }
            }
            const T alpha = sHyp[pp * 4], tau = sHyp[pp * 4 + 1], rho = sHyp[pp * 4 + 2];
            const T dg = sDeg[bl * P + pp];
            if (!first) {
                yv = *reinterpret_cast<const V*>(tile + (unsigned)pp * (CH * (unsigned)sizeof(T)) + lane_bytes);
                dv = lap_events<T, VEC>(tile, yv, lidx, lptr[pp], lptr[pp + 1], lane_bytes);
                const T eta_prev = sHyp[pp * 4 + 3];
#pragma unroll
                for (int v = 0; v < VEC; ++v) {
                    if constexpr (!LEAN) {
                        if (p.hasD) dv.v[v] = clamp_sym(dv.v[v], p.D);
                        Uv.v[v] = clamp_sym(add_rn(Uv.v[v], mul_rn(dv.v[v], eta_prev)), p.Uc_prev);
                    } else {
                        Uv.v[v] = fmin(fmax(add_rn(Uv.v[v], mul_rn(dv.v[v], eta_prev)), -p.Uc_prev), p.Uc_prev);
                    }
                }
            }
            V yn, rv;
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                const T y = yv.v[v], U = Uv.v[v];
                T rr = av.v[v];                                            // the fused path's a already holds AtA y - Atb
                if constexpr (!LEAN) {
                    if (p.atb) rr = sub_rn(rr, atbv.v[v]);
                }
                rr = add_rn(rr, sign_times(y, tau));
                rr = add_rn(rr, mul_rn(U, dg));
                rr = add_rn(rr, mul_rn(dv.v[v], rho));
                // min/max clamps: identical to torch.clamp for finite values; a NaN would be swallowed, so the raw
                // gradient joins the non-finite accumulator below (0*x is NaN iff x is Inf/NaN) and any hit sends the
                // batch to the guarded per-iteration path (reference guards :55-61,84-86,102-104)
                const T g = fmin(fmax(rr, nG), p.G);
                const T z = fmin(fmax(sub_rn(y, mul_rn(alpha, g)), nV), p.V);
                rv.v[v] = rr;
                yn.v[v] = z;
                if constexpr (LEAN) nonfinite = fma(rr, (T)0, nonfinite);
                else nonfinite = fma(y, (T)0, fma(U, (T)0, fma(rr, (T)0, nonfinite)));
            }
            if (act_i) {
                st_vec<T, VEC>(p.y_next + off, yn);
                if (p.U_out && !first) st_stream<T, VEC>(p.U_out + off, Uv);
                if constexpr (!LEAN) {
                    if (p.graw) st_stream<T, VEC>(p.graw + off, rv);
                }
                if constexpr (sizeof(T) == 4) {
                    if (do_split) store_split<VEC>(p.sp, off, yn, sc, 1.f);
#pragma unroll
                    for (int v = 0; v < VEC; ++v) amax_f = fmaxf(amax_f, fabsf(yn.v[v]));
                    if constexpr (LEAN) {
                        if (sums) {
#pragma unroll
                            for (int v = 0; v < VEC; ++v) {
                                vsum.v[v] += yn.v[v];
                                sq_tot = fmaf((float)yn.v[v], (float)yn.v[v], sq_tot);
                            }
                        }
                    }
                }
            }
        }
    }
    if constexpr (LEAN && sizeof(T) == 4) {
        if (sums) {      // TB == 1: sum of the tile's rows over the warps -> agent_sum[b0][chunk]
            __syncthreads();                                     // every warp is done with the y_k tile
            float* red = reinterpret_cast<float*>(S0);
            *reinterpret_cast<V*>(red + warp * CH + lane * VEC) = vsum;
            __syncthreads();
            if (threadIdx.x < CH) {
                float a = 0.f;
                for (int wq = 0; wq < nwarps; ++wq) a += red[wq * CH + threadIdx.x];
                p.agent_sum[(unsigned)b0 * p.n + chunk * CH + threadIdx.x] = a;
            }
        }
    }
    }   // chunk loop
    if constexpr (LEAN && sizeof(T) == 4) {
        if (sums) {
            float w = sq_tot;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) w += __shfl_xor_sync(0xffffffffu, w, o);
            __syncthreads();
            if (lane == 0) sAlpha[warp] = w;
            __syncthreads();
            if (threadIdx.x == 0) {
                double t = 0;
                for (int wq = 0; wq < nwarps; ++wq) t += (double)sAlpha[wq];
                p.sq_part[blockIdx.x] = t;
            }
        }
    }
    if constexpr (sizeof(T) == 4) {
        amax_bits = __float_as_uint(amax_f);
        if (p.sp.amax_out) publish_amax(amax_bits, p.sp.amax_out, sAmax);
    }
    if (p.flags) {
        // the fused path only reports "something was not finite" (all bits): the modules then re-run the batch on
        // the guarded per-iteration path, whose step kernel reports the individual conditions
        const unsigned bad = __reduce_or_sync(0xffffffffu, (nonfinite != (T)0) ? 0xFu : 0u);
        if (bad && lane == 0) atomicOr(p.flags, (int)bad);
    }
}

// LEAN: the fused fp16 training path on full tiles, levels k >= 1 -- the saved stream is the residual AtA y - Atb, no
// delta clamp, no dense upstream gradient (the loss term arrives as coef * (y - label)), n % (32 VEC) == 0 and
// B % TB == 0.  Same arithmetic as the generic form; masks are |x| <= c compares (identical to torch's closed
// interval, false for NaN), clamps are min/max, the per-agent scalars come from shared memory.  ncu (round 1): the
// generic form executes ~600 straight-line instructions per 128-unknown row segment, two thirds of them guards,
// zero-fills and select chains.
template <typename T, int VEC, bool LEAN, int NTHR = kStepThreads>
__global__ void __launch_bounds__(NTHR, NTHR > kStepThreads ? DADMM_LEVEL_MINB_BWD_LEAN - 1 : (LEAN ? DADMM_LEVEL_MINB_BWD_LEAN : DADMM_LEVEL_MINB_BWD))
level_bwd_kernel(const LevelBwdParams<T> p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int CH = 32 * VEC;
    using V = Vec<T, VEC>;
    const int P = p.P, R = p.TB * P;
    unsigned char* S0 = smem_raw;                                   // y_k tile
    unsigned char* S1 = S0 + (size_t)R * CH * sizeof(T);            // adjoint of the unclamped 2L y_k
    T* sAcc = reinterpret_cast<T*>(S1 + (size_t)R * CH * sizeof(T));   // [R][4]: per-row sums of d/d (alpha, tau, rho, eta_prev) over the CTA's chunks
    T* sHyp = sAcc + (size_t)R * 4;                                // staged per-agent scalars (stage_scalars)
    T* sDeg = sHyp + 4 * P;
    int32_t* sPtr = reinterpret_cast<int32_t*>(sDeg + R);
    int32_t* sIdx = sPtr + p.TB * (P + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const int lane_bytes = lane * VEC * (int)sizeof(T);
    const int nchunks = (p.n + CH - 1) / CH;
    // A CTA owns TB problems and walks a contiguous range of chunks: the neighbour lists, the split scale and
    // max|alpha| are set up once, and the four hyper-parameter sums of a row are accumulated per lane in shared
    // memory across the chunks (one shuffle tree + one partial-sum row per CTA instead of one per chunk).
    const int cs = blockIdx.x % p.csplit;
    const int b0 = (blockIdx.x / p.csplit) * p.TB;
    const int cpc = (nchunks + p.csplit - 1) / p.csplit;
    const int chunk_begin = cs * cpc, chunk_end = min(nchunks, chunk_begin + cpc);
    const bool first = p.first != 0, top = p.top != 0;
    __shared__ float sAlpha[32];

    // call-constant inputs first, then the dependency wait (see level_fwd_kernel)
    const bool staged = p.list_cap > 0;
    if (!first && staged) stage_lists<T, VEC>(sPtr, sIdx, p.TB, P, p.B, b0, p.list_cap, p.lst_ptr, p.lst_idx, p.gid);
    for (int r = threadIdx.x; r < R * 4; r += blockDim.x) sAcc[r] = (T)0;
    stage_scalars<T>(sHyp, sDeg, p.TB, P, p.B, b0, p.hyp_k, p.hyp_prev, p.deg, p.gid);
    pdl_wait();
    pdl_trigger();

    // loss term coef * (y_k - label): the coefficient may live on the device (no host round trip in autograd's backward)
    const T* label = p.label;
    T coef_prev = p.coef_prev;
    if (p.coef_dev) {
        coef_prev = (T)__ldg(p.coef_dev);
        if (coef_prev == (T)0) label = nullptr;
    }

    // scale of the fused fp16 split of gAtAy_k:  |gAtAy| <= max_p alpha_p * max|adj(y_{k+1})|
    float sc1 = 1.f, sc2 = 1.f;
    bool do_split = false;
    if constexpr (sizeof(T) == 4) {
        do_split = p.sp.hi != nullptr && !first;
        if (do_split) {
            const float amax_alpha = block_max_alpha(p.hyp_k, P, sAlpha);
            const int e = split_exponent(amax_alpha * __uint_as_float(__ldg(p.sp.amax_in)));
            sc1 = pow2_of(e / 2);
            sc2 = pow2_of(e - e / 2);
            if (blockIdx.x == 0 && threadIdx.x == 0) *p.sp.exp = e;
        }
    }

    const T nG = -p.G, nUc = -p.Uc_prev;

    for (int chunk = chunk_begin; chunk < chunk_end; ++chunk) {
    const int i = chunk * CH + lane * VEC;
    const bool act_i = LEAN || i < p.n;
    if (chunk != chunk_begin) __syncthreads();          // the previous chunk's last phase still reads the tiles
    for (int bl = 0; bl < p.TB; ++bl) {
        const int b = b0 + bl;
        const T* src = p.y + (((unsigned)b * P) * p.n + i);
        for (int pp = warp; pp < P; pp += nwarps) {
            unsigned char* dst = S0 + (unsigned)(bl * P + pp) * (CH * (unsigned)sizeof(T)) + lane_bytes;
            if constexpr (LEAN && sizeof(V) == 16) {
                cp_async16(dst, src + (unsigned)pp * p.n);
            } else {
                V v = vzero<T, VEC>();
                if (act_i && b < p.B) v = ld_vec<T, VEC>(src + (unsigned)pp * p.n);
                *reinterpret_cast<V*>(dst) = v;
                if (!first && b >= p.B) *reinterpret_cast<V*>(S1 + (unsigned)(bl * P + pp) * (CH * (unsigned)sizeof(T)) + lane_bytes) = v;
            }
        }
    }
    if constexpr (LEAN && sizeof(V) == 16) cp_async_wait_all();
    __syncthreads();

    if constexpr (LEAN) {
    for (int bl = 0; bl < p.TB; ++bl) {
        const int b = b0 + bl;
        const unsigned base = ((unsigned)b * P) * p.n + i;          // 32-bit element offsets: B*P*n < 2^31 (host-checked)
        const unsigned char* tile = S0 + (size_t)bl * P * CH * sizeof(T);
        unsigned char* tile1 = S1 + (size_t)bl * P * CH * sizeof(T);
        const int32_t* lptr = staged ? sPtr + bl * (P + 1) : p.lst_ptr + (p.gid ? __ldg(p.gid + b) : 0) * P;
        const int32_t* lidx = staged ? sIdx + bl * p.list_cap : p.lst_idx;
        V labv = vzero<T, VEC>();
        if (label) labv = ld_vec<T, VEC>(label + ((unsigned)b * p.n + i));
        struct In { V t, r, u, c; };
        auto issue = [&](int q, In& L) {
            L.c = vzero<T, VEC>();
            if (q < P) {
                const unsigned o = base + (unsigned)q * p.n;
                L.t = ld_vec<T, VEC>(p.Tb + o);
                L.r = ld_stream<T, VEC>(p.graw + o);
                L.u = ld_stream<T, VEC>(p.U_prev + o);
                if (!top) L.c = ld_vec<T, VEC>(p.C + o);
            }
        };
        In cur, nxt;
#if DADMM_LEVEL_BWD_LEAN_PREFETCH
        issue(warp, cur);
#endif
        for (int pp = warp; pp < P; pp += nwarps) {
            const unsigned off = base + (unsigned)pp * p.n;
#if DADMM_LEVEL_BWD_LEAN_PREFETCH
            issue(pp + nwarps, nxt);            // next row's streams fly while this row computes (different rows: no aliasing)
#else
            issue(pp, cur);
#endif
            const V tv = cur.t, rv = cur.r, uv = cur.u, cv = cur.c;
            const T alpha = sHyp[pp * 4], tau = sHyp[pp * 4 + 1], rho = sHyp[pp * 4 + 2], eta_prev = sHyp[pp * 4 + 3];
            const T dg = sDeg[bl * P + pp];
            const V yv = *reinterpret_cast<const V*>(tile + (unsigned)pp * (CH * (unsigned)sizeof(T)) + lane_bytes);
            const V draw = lap_adj<T, VEC>(tile, yv, lidx, lptr[pp], lptr[pp + 1], lane_bytes);
            V o_ga, o_c, o_dir, o_db;
            T pa = (T)0, pt = (T)0, pr = (T)0, pe = (T)0;
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                const T y = yv.v[v], d = draw.v[v];
                // r_k = (AtA y - Atb) + sign(y) tau + U_k deg + d_k rho with U_k = clamp(U_{k-1} + d_k eta_{k-1})
                const T w = add_rn(uv.v[v], mul_rn(d, eta_prev));
                const T Uk = fmin(fmax(w, nUc), p.Uc_prev);
                const T sg = sign_times(y, (T)1);
                T rr = add_rn(rv.v[v], mul_rn(sg, tau));
                rr = add_rn(rr, mul_rn(Uk, dg));
                rr = add_rn(rr, mul_rn(d, rho));
                const T g = fmin(fmax(rr, nG), p.G);
                const T z = sub_rn(y, mul_rn(alpha, g));
                const T zb = (fabs(z) <= p.V) ? tv.v[v] : (T)0;
                pa -= zb * g;
                const T rb = (fabs(rr) <= p.G) ? (-alpha * zb) : (T)0;
                pt += rb * sg;
                pr += rb * d;
                o_ga.v[v] = rb;
                const T uk = cv.v[v] + dg * rb;
                const T um = (fabs(w) <= p.Uc_prev) ? uk : (T)0;
                pe += um * d;
                o_c.v[v] = um;
                o_db.v[v] = rho * rb + eta_prev * um;
                T dir = zb;
                if (label) dir += coef_prev * (y - labv.v[v]);
                o_dir.v[v] = dir;
            }
            *reinterpret_cast<V*>(tile1 + (unsigned)pp * (CH * (unsigned)sizeof(T)) + lane_bytes) = o_db;
            if constexpr (sizeof(T) == 4) {
                if (do_split) store_split<VEC>(p.sp, off, o_ga, sc1, sc2);
                else st_vec<T, VEC>(p.ga + off, o_ga);
            } else {
                st_vec<T, VEC>(p.ga + off, o_ga);
            }
            st_vec<T, VEC>(p.C + off, o_c);
            st_vec<T, VEC>(p.Tb + off, o_dir);     // + 2L db in the last phase (same thread re-reads it)
            {   // lanes 0/8/16/24 end up with the row's four sums (only this warp touches the row's slots)
                const T k4 = warp_sum4(pa, pt, pr, pe, lane);
                if ((lane & 7) == 0) sAcc[(bl * P + pp) * 4 + (lane >> 3)] += k4;
            }
#if DADMM_LEVEL_BWD_LEAN_PREFETCH
            cur = nxt;
#endif
        }
    }
    } else {
    struct Row { V t, c, r, u, d, g, l; };
    for (int bl = 0; bl < p.TB; ++bl) {
        const int b = b0 + bl;
        if (b >= p.B) break;
        const unsigned base = ((unsigned)b * P) * p.n + i;          // 32-bit element offsets: B*P*n < 2^31 (host-checked)
        const int node0 = (p.gid ? __ldg(p.gid + b) : 0) * P;
        const unsigned char* tile = S0 + (size_t)bl * P * CH * sizeof(T);
        unsigned char* tile1 = S1 + (size_t)bl * P * CH * sizeof(T);
        const int32_t* lptr = staged ? sPtr + bl * (P + 1) : p.lst_ptr + (p.gid ? __ldg(p.gid + b) : 0) * P;
        const int32_t* lidx = staged ? sIdx + bl * p.list_cap : p.lst_idx;
        const T* lab = label ? label + ((unsigned)b * p.n + i) : nullptr;
        auto issue = [&](int pp, Row& L) {
            L.t = L.c = L.r = L.u = L.d = L.g = vzero<T, VEC>();
            if (pp < P && act_i) {
                const unsigned off = base + (unsigned)pp * p.n;
                L.t = ld_vec<T, VEC>(p.Tb + off);
                L.r = ld_stream<T, VEC>(p.graw + off);
                if (first) {
                    L.d = ld_stream<T, VEC>(p.d0 + off);
                    if (p.graw_is_residual) L.u = ld_stream<T, VEC>(p.U_prev + off);      // U_0
                } else {
                    if (!top) L.c = ld_vec<T, VEC>(p.C + off);
                    L.u = ld_stream<T, VEC>(p.U_prev + off);
                    if (p.gY_prev) L.g = ld_stream<T, VEC>(p.gY_prev + off);
                }
            }
        };
        V labv = vzero<T, VEC>();
        if (lab && act_i) labv = ld_vec<T, VEC>(lab);
        Row cur, nxt;
#if DADMM_LEVEL_BWD_PREFETCH
        issue(warp, cur);
#endif
        for (int pp = warp; pp < P; pp += nwarps) {
#if DADMM_LEVEL_BWD_PREFETCH
            issue(pp + nwarps, nxt);
#else
            issue(pp, cur);
#endif
            const unsigned off = base + (unsigned)pp * p.n;
            const T alpha = __ldg(p.hyp_k + pp * 4), rho = __ldg(p.hyp_k + pp * 4 + 2);
            const T tau = p.graw_is_residual ? __ldg(p.hyp_k + pp * 4 + 1) : (T)0;
            const T dg = (T)__ldg(p.deg + node0 + pp);
            const V yv = *reinterpret_cast<const V*>(tile + (unsigned)pp * (CH * (unsigned)sizeof(T)) + lane_bytes);
            V draw = cur.d;
            T eta_prev = (T)0;
            if (!first) {
                draw = lap_adj<T, VEC>(tile, yv, lidx, lptr[pp], lptr[pp + 1], lane_bytes);
                eta_prev = __ldg(p.hyp_prev + pp * 4 + 3);
            }
            V o_ga, o_c, o_dir, o_db;
            T pa = (T)0, pt = (T)0, pr = (T)0, pe = (T)0;
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                const T y = yv.v[v];
                const bool mD = (first || !p.hasD) ? true : in_closed(draw.v[v], p.D);
                const T d = (first || !p.hasD) ? draw.v[v] : clamp_sym(draw.v[v], p.D);
                T rr = cur.r.v[v];
                if (p.graw_is_residual) {
                    // r_k = (AtA y - Atb) + sign(y) tau + U_k deg + d_k rho with U_k = clamp(U_{k-1} + d_k eta_{k-1})
                    const T Uk = first ? cur.u.v[v] : clamp_sym(add_rn(cur.u.v[v], mul_rn(d, eta_prev)), p.Uc_prev);
                    rr = add_rn(rr, mul_rn(sign_of(y), tau));
                    rr = add_rn(rr, mul_rn(Uk, dg));
                    rr = add_rn(rr, mul_rn(d, rho));
                }
                const T g = clamp_sym(rr, p.G);
                const T z = sub_rn(y, mul_rn(alpha, g));
                const T zb = in_closed(z, p.V) ? cur.t.v[v] : (T)0;
                pa -= zb * g;
                const T rb = in_closed(rr, p.G) ? (-alpha * zb) : (T)0;
                pt += rb * sign_of(y);
                pr += rb * d;
                o_ga.v[v] = rb;
                if (!first) {
                    const T uk = cur.c.v[v] + dg * rb;
                    const T w = add_rn(cur.u.v[v], mul_rn(d, eta_prev));
                    const T um = in_closed(w, p.Uc_prev) ? uk : (T)0;
                    pe += um * d;
                    o_c.v[v] = um;
                    const T db = rho * rb + eta_prev * um;
                    o_db.v[v] = (mD && act_i) ? db : (T)0;
                    T dir = zb + cur.g.v[v];
                    if (lab) dir += coef_prev * (y - labv.v[v]);
                    o_dir.v[v] = dir;
                }
            }
            if (!first) {
                *reinterpret_cast<V*>(tile1 + (unsigned)pp * (CH * (unsigned)sizeof(T)) + lane_bytes) = o_db;
                if (act_i) {
                    if constexpr (sizeof(T) == 4) {
                        if (do_split) store_split<VEC>(p.sp, off, o_ga, sc1, sc2);
                        else st_vec<T, VEC>(p.ga + off, o_ga);
                    } else {
                        st_vec<T, VEC>(p.ga + off, o_ga);
                    }
                    st_vec<T, VEC>(p.C + off, o_c);
                    st_vec<T, VEC>(p.Tb + off, o_dir);     // + 2L db in the last phase (same thread re-reads it)
                }
            }
            {   // lanes 0/8/16/24 end up with the row's four sums (only this warp touches the row's slots)
                const T k4 = warp_sum4(pa, pt, pr, pe, lane);
                if ((lane & 7) == 0) sAcc[(bl * P + pp) * 4 + (lane >> 3)] += k4;
            }
#if DADMM_LEVEL_BWD_PREFETCH
            cur = nxt;
#endif
        }
    }
    }   // generic form
    if (!first) {
    __syncthreads();
    for (int bl = 0; bl < p.TB; ++bl) {
        const int b = b0 + bl;
        if (!LEAN && b >= p.B) break;
        const unsigned base = ((unsigned)b * P) * p.n + i;          // 32-bit element offsets: B*P*n < 2^31 (host-checked)
        const unsigned char* tile1 = S1 + (size_t)bl * P * CH * sizeof(T);
        const int32_t* lptr = staged ? sPtr + bl * (P + 1) : p.lst_ptr + (p.gid ? __ldg(p.gid + b) : 0) * P;
        const int32_t* lidx = staged ? sIdx + bl * p.list_cap : p.lst_idx;
        for (int pp = warp; pp < P; pp += nwarps) {
            const unsigned off = base + (unsigned)pp * p.n;
            V s = vzero<T, VEC>();
            if (act_i) s = ld_vec<T, VEC>(p.Tb + off);          // issued ahead of the shared-memory gather (an L2 hit: this thread wrote it)
            const V xq = *reinterpret_cast<const V*>(tile1 + (unsigned)pp * (CH * (unsigned)sizeof(T)) + lane_bytes);
            const V lt = lap_adj<T, VEC>(tile1, xq, lidx, lptr[pp], lptr[pp + 1], lane_bytes);
            if (act_i) {
#pragma unroll
                for (int v = 0; v < VEC; ++v) s.v[v] += lt.v[v];
                st_vec<T, VEC>(p.Tb + off, s);
            }
        }
    }
    }   // !first
    }   // chunk loop

    // one partial-sum row per CTA: (d alpha, d tau, d rho, d eta_prev) of every (problem, agent) of the tile
    __syncthreads();
    for (int r = threadIdx.x; r < R * 4; r += blockDim.x) {
        const int b = b0 + (r >> 2) / P, pp = (r >> 2) % P;
        if (b < p.B) p.partials[(((unsigned)cs * p.B + b) * P + pp) * 4 + (r & 3)] = sAcc[r];
    }
}

// hyper-parameter gradient rows from the level partials [K][nchunks][B][P][4] = (d alpha_k, d tau_k, d rho_k, d eta_{k-1}):
// ONE launch after the reverse sweep, CTA (p, k) finishes agent p of level k (the per-level launches this replaces sat
// on the critical path between a backward level and its contraction; the sums and their order are the same)
template <typename T>
__global__ void __launch_bounds__(256) reduce_levels_kernel(const T* __restrict__ part, size_t level_stride, int nchunks, int B,
                                                            int P, T* __restrict__ ghyp) {
    pdl_wait();
    pdl_trigger();
    const int pp = blockIdx.x, k = blockIdx.y;
    part += (size_t)k * level_stride;
    T* row_k = ghyp + (size_t)k * P * 4;
    T* row_prev = k ? row_k - (size_t)P * 4 : nullptr;
    double acc[4] = {0, 0, 0, 0};
    const long long rows = (long long)nchunks * B;
    for (long long rI = threadIdx.x; rI < rows; rI += blockDim.x) {
        const Vec<T, 4> q = ld_vec<T, 4>(part + (rI * P + pp) * 4);
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[c] += (double)q.v[c];
    }
    __shared__ double sh[4][8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const double s = warp_sum(acc[c]);
        if (lane == 0) sh[c][warp] = s;
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        double s = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += sh[threadIdx.x][w];
        if (threadIdx.x < 3) row_k[pp * 4 + threadIdx.x] = (T)s;
        else if (row_prev) row_prev[pp * 4 + 3] = (T)s;
    }
}

// sumsq[k] = sum of the per-CTA partials the forward level k left in part[k][0..grid), for k = k_first + blockIdx.x
__global__ void __launch_bounds__(256) sumsq_final_kernel(const double* __restrict__ part, size_t row_stride, int grid, int k_first,
                                                          double* __restrict__ sumsq) {
    pdl_wait();
    pdl_trigger();
    const int k = k_first + blockIdx.x;
    double a = 0;
    for (int i = threadIdx.x; i < grid; i += blockDim.x) a += part[(size_t)k * row_stride + i];
    __shared__ double sh[8];
    a = warp_sum(a);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = a;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0;
        for (int w = 0; w < 8; ++w) t += sh[w];
        sumsq[k] = t;
    }
}

// partial[k][blk] = (sum S[k] * label, sum label^2) over the block's share of [B*n]
template <typename T>
__global__ void __launch_bounds__(256) loss_sums_partial_kernel(const T* __restrict__ S, const T* __restrict__ label, long long tot,
                                                                double* __restrict__ part) {
    const T* Sk = S + (size_t)blockIdx.y * tot;
    double d = 0, l2 = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < tot; i += (long long)gridDim.x * blockDim.x) {
        const double l = (double)label[i];
        d += (double)Sk[i] * l;
        l2 += l * l;
    }
    __shared__ double sh[2][8];
    d = warp_sum(d);
    l2 = warp_sum(l2);
    if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = d; sh[1][threadIdx.x >> 5] = l2; }
    __syncthreads();
    if (threadIdx.x < 2) {
        double t = 0;
        for (int w = 0; w < 8; ++w) t += sh[threadIdx.x][w];
        part[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 2 + threadIdx.x] = t;
    }
}
// need[k] = 1 where sum Y^2 - 2 <S, label> + P sum label^2 cancels more than three digits (a near-converged iterate: the
// fp32 partial sums behind S and sum Y^2 then no longer carry the loss to fp32 accuracy) -- the caller re-evaluates those
// iterations from Y
template <typename T>
__global__ void __launch_bounds__(256) loss_sums_final_kernel(const double* __restrict__ part, int nblk, const double* __restrict__ sumsq,
                                                              int P, double inv, T* __restrict__ losses, int* __restrict__ need) {
    const int k = blockIdx.x;
    double d = 0, l2 = 0;
    for (int i = threadIdx.x; i < nblk; i += blockDim.x) {
        d += part[((size_t)k * nblk + i) * 2];
        l2 += part[((size_t)k * nblk + i) * 2 + 1];
    }
    __shared__ double sh[2][8];
    d = warp_sum(d);
    l2 = warp_sum(l2);
    if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = d; sh[1][threadIdx.x >> 5] = l2; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double td = 0, tl = 0;
        for (int w = 0; w < 8; ++w) { td += sh[0][w]; tl += sh[1][w]; }
        const double sse = sumsq[k] - 2.0 * td + (double)P * tl;
        losses[k] = (T)(sse * inv);
        if (need) need[k] = (sse < 1e-3 * (sumsq[k] + (double)P * tl)) ? 1 : 0;
    }
}

// T = gY_last + coef * (Y_last - label): adjoint of y_K entering the reverse sweep
template <typename T>
__global__ void __launch_bounds__(256) seed_adjoint_kernel(const T* __restrict__ Ylast, const T* __restrict__ gYlast,
                                                           const T* __restrict__ label, T coef, const double* __restrict__ coef_dev,
                                                           int B, int P, int n, T* __restrict__ out, unsigned* amax_out) {
    pdl_wait();
    pdl_trigger();
    if (coef_dev) {                      // device-side coefficient (see LevelBwdParams::coef_dev)
        coef = (T)__ldg(coef_dev);
        if (coef == (T)0) label = nullptr;
    }
    const long long rows = (long long)B * P;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    __shared__ unsigned sAmax[32];
    unsigned amax_bits = 0;
    for (long long r = (long long)blockIdx.x * nw + warp; r < rows; r += (long long)gridDim.x * nw) {
        const long long base = r * n;
        const T* lr = label ? label + (r / P) * n : nullptr;
        for (int i = lane; i < n; i += 32) {
            T v = gYlast ? gYlast[base + i] : (T)0;
            if (lr) v += coef * (Ylast[base + i] - lr[i]);
            out[base + i] = v;
            amax_bits = max(amax_bits, __float_as_uint(fabsf((float)v)));
        }
    }
    if (amax_out) publish_amax(amax_bits, amax_out, sAmax);
}

}  // namespace dadmm

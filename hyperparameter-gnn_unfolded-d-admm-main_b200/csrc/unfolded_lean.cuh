// unfolded_lean.cuh -- the level kernels of the fused fp32 training / inference path on full tiles, second generation:
// same operations, same order and same rounding as the LEAN instantiations of level_fwd_kernel / level_bwd_kernel
// (unfolded.cuh), written for the instruction issue rate.
//
// ncu on the first generation (round 2, profiles/r02_ncu_levels_v24.txt): the forward level executes 441 warp instructions
// per 128-unknown row segment (723 M per launch), issue slots 65 % busy, DRAM 62 % -- issue and HBM co-limit it, and under
// the 1000 W power cap (1.6 GHz instead of 1.93) the issue side wins: 0.98 ms alone, 1.08 ms inside the step.  Of the 441,
// 166 are the consensus gather (per event: a generic-space index load, two address instructions, LDS.128, 8 FADD) and ~110
// the element-wise update.  Here
//   * arithmetic runs on Blackwell's packed fp32 pipe (add/sub/mul/fma.rn.f32x2 -> FADD2 / FMUL2 / FFMA2): one instruction
//     per PAIR of unknowns, each half rounded exactly as the scalar op it replaces (add.rn is add.rn; nothing is
//     contracted that was not contracted before), so every bit-exactness test of the scalar kernels applies unchanged;
//   * the neighbour lists live in shared memory as BYTE OFFSETS of the neighbour's row in the tile (no index scaling in
//     the loop, LDS instead of generic loads);
//   * everything a row needs arrives as 16-byte shared / global accesses that land directly in register pairs.
// The kernels take the same parameter blocks and the same shared-memory layout as the first generation (the host picks
// them when the tile's lists fit in shared memory, which the first generation's staged path also required).
#pragma once
#include "unfolded.cuh"

namespace dadmm {
namespace lean {

using u64 = unsigned long long;

__device__ __forceinline__ u64 pk2(float lo, float hi) {
    u64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void upk2(u64 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 dup2(float x) { return pk2(x, x); }
__device__ __forceinline__ u64 add2(u64 a, u64 b) {
    u64 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ u64 sub2(u64 a, u64 b) {
    u64 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ u64 mul2(u64 a, u64 b) {
    u64 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
    u64 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
// min(max(x, -c), c) on both halves (identical to torch.clamp for finite x; a NaN is swallowed -- see level_fwd_kernel)
__device__ __forceinline__ u64 clamp2(u64 v, float c) {
    float a, b;
    upk2(v, a, b);
    return pk2(fminf(fmaxf(a, -c), c), fminf(fmaxf(b, -c), c));
}
// sign(y) * tau on both halves (sign_times)
__device__ __forceinline__ u64 sign2(u64 y, float tau) {
    float a, b;
    upk2(y, a, b);
    return pk2(sign_times(a, tau), sign_times(b, tau));
}
// (|x| <= c) ? v : 0 on both halves
__device__ __forceinline__ u64 mask2(u64 x, float c, u64 v) {
    float xa, xb, va, vb;
    upk2(x, xa, xb);
    upk2(v, va, vb);
    return pk2(fabsf(xa) <= c ? va : 0.f, fabsf(xb) <= c ? vb : 0.f);
}
__device__ __forceinline__ float hsum2(u64 v) {
    float a, b;
    upk2(v, a, b);
    return a + b;
}

// software prefetch of a 16-byte-per-lane row access the warp will make one row later (no registers held meanwhile)
__device__ __forceinline__ void prefetch_row(const float* p, int level) {
    if (level == 1) asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
    else asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

struct Q4 { u64 a, b; };      // four consecutive fp32 values as two packed pairs (one 16-byte access)
__device__ __forceinline__ Q4 ldq(const void* p) {
    const ulonglong2 t = *reinterpret_cast<const ulonglong2*>(p);
    return Q4{t.x, t.y};
}
__device__ __forceinline__ Q4 ldq_stream(const float* p) {
    const int4 t = __ldcs(reinterpret_cast<const int4*>(p));
    Q4 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r.a) : "r"(t.x), "r"(t.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(r.b) : "r"(t.z), "r"(t.w));
    return r;
}
__device__ __forceinline__ void stq(void* p, const Q4& v) { *reinterpret_cast<ulonglong2*>(p) = make_ulonglong2(v.a, v.b); }
__device__ __forceinline__ void stq_stream(float* p, const Q4& v) {
    int4 t;
    asm("mov.b64 {%0, %1}, %2;" : "=r"(t.x), "=r"(t.y) : "l"(v.a));
    asm("mov.b64 {%0, %1}, %2;" : "=r"(t.z), "=r"(t.w) : "l"(v.b));
    __stcs(reinterpret_cast<int4*>(p), t);
}

// Neighbour lists of the tile's problems as byte offsets of the neighbour's row inside the problem's tile
// (row = 128 unknowns = 512 bytes): sPtr[bl][0..P] list bounds, sOff[bl][e] = 512 * neighbour id
__device__ __forceinline__ void stage_lists_bytes(int32_t* sPtr, int32_t* sOff, int TB, int P, int b0, int cap,
                                                  const int32_t* __restrict__ lst_ptr, const int32_t* __restrict__ lst_idx,
                                                  const int32_t* __restrict__ gid) {
    for (int bl = 0; bl < TB; ++bl) {
        const int node0 = (gid ? __ldg(gid + b0 + bl) : 0) * P;
        const int e0 = __ldg(lst_ptr + node0);
        for (int q = threadIdx.x; q <= P; q += blockDim.x) sPtr[bl * (P + 1) + q] = __ldg(lst_ptr + node0 + q) - e0;
        const int cnt = __ldg(lst_ptr + node0 + P) - e0;
        for (int e = threadIdx.x; e < cnt; e += blockDim.x) sOff[bl * cap + e] = __ldg(lst_idx + e0 + e) << 9;
    }
}

// 2L x for one row in the reference's event order: acc <- acc + (x_q - x_e), sequentially, each op rounded once
__device__ __forceinline__ Q4 lap_events(const unsigned char* tile_lane, const Q4& xq, const int32_t* sOff, int e0, int e1) {
    u64 a = 0ull, b = 0ull;        // (+0, +0)
#pragma unroll 4
    for (int e = e0; e < e1; ++e) {
        const Q4 xj = ldq(tile_lane + sOff[e]);
        a = add2(a, sub2(xq.a, xj.a));
        b = add2(b, sub2(xq.b, xj.b));
    }
    return Q4{a, b};
}
// 2L x = 2 sum_{j in N(q)} (x_q - x_j) over the plain neighbour list: the same differences the event order accumulates
// (each of them twice, interleaved with the other endpoint's visits), taken once and doubled -- half the shared-memory
// gathers and additions.  Differences first, so nothing cancels that the reference's form does not cancel; the result is
// the reference's delta to rounding (a few ulp), not to the bit.  ncu (round 2) has the forward level's L1 data pipe 74-87 %
// busy, 60 % of it with the event gather: this, not DRAM, is what the level waits for.
__device__ __forceinline__ Q4 lap_half(const unsigned char* tile_lane, const Q4& xq, const int32_t* sOff, int e0, int e1) {
    u64 a = 0ull, b = 0ull;
#pragma unroll 4
    for (int e = e0; e < e1; ++e) {
        const Q4 xj = ldq(tile_lane + sOff[e]);
        a = add2(a, sub2(xq.a, xj.a));
        b = add2(b, sub2(xq.b, xj.b));
    }
    return Q4{add2(a, a), add2(b, b)};
}
// 2L x = 2 (deg x_q - sum_j x_j) over the plain neighbour list (backward: no bit pattern to reproduce)
__device__ __forceinline__ Q4 lap_adj(const unsigned char* tile_lane, const Q4& xq, const int32_t* sOff, int e0, int e1) {
    u64 a = 0ull, b = 0ull;
#pragma unroll 4
    for (int e = e0; e < e1; ++e) {
        const Q4 xj = ldq(tile_lane + sOff[e]);
        a = add2(a, xj.a);
        b = add2(b, xj.b);
    }
    const u64 dq = dup2((float)(e1 - e0)), two = dup2(2.f);
    return Q4{mul2(two, sub2(mul2(dq, xq.a), a)), mul2(two, sub2(mul2(dq, xq.b), b))};
}

// scaled fp16 (hi, lo) split of four values: x = v * sc; hi = rn_fp16(x); lo = rn_fp16(x - hi)   (store_split)
__device__ __forceinline__ void store_split4(const SplitOut& sp, unsigned off, const Q4& v, u64 sc2) {
    const u64 x0 = mul2(v.a, sc2), x1 = mul2(v.b, sc2);
    float a, b, c, d;
    upk2(x0, a, b);
    upk2(x1, c, d);
    __half2 h[2], l[2];
    h[0] = __floats2half2_rn(a, b);
    h[1] = __floats2half2_rn(c, d);
    const float2 f0 = __half22float2(h[0]), f1 = __half22float2(h[1]);
    l[0] = __floats2half2_rn(a - f0.x, b - f0.y);
    l[1] = __floats2half2_rn(c - f1.x, d - f1.y);
    *reinterpret_cast<uint2*>(sp.hi + off) = *reinterpret_cast<const uint2*>(h);
    *reinterpret_cast<uint2*>(sp.lo + off) = *reinterpret_cast<const uint2*>(l);
}

// MINB (CTAs per SM the register allocation is sized for) is a template argument: the host picks it per launch
// (lean_minb() in dadmm_abi.cu; occupancy against spills in the row loop was measured, not guessed)

// ------------------------------------------------------------------------------------------------------------------
// forward level k >= 1 (see unfolded.cuh for the recurrence): reads a_k' = AtA y_k - Atb, y_k (tile), U_{k-1};
// writes y_{k+1} -> Y[k], U_k, the fp16 split of y_{k+1}, optionally the label-free loss sums
// ------------------------------------------------------------------------------------------------------------------
template <int NTHR, int MINB>
__global__ void __launch_bounds__(NTHR, MINB)
level_fwd_lean_kernel(const LevelFwdParams<float> p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int CH = 128, ROWB = 512;
    const int P = p.P, R = p.TB * P;
    unsigned char* S0 = smem_raw;                                              // y_k tile [R][CH]
    float* sHyp = reinterpret_cast<float*>(S0 + (size_t)R * ROWB);
    float* sDeg = sHyp + 4 * P;
    int32_t* sPtr = reinterpret_cast<int32_t*>(sDeg + R);
    int32_t* sOff = sPtr + p.TB * (P + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = NTHR >> 5;
    const int lane_bytes = lane * 16;
    const int nchunks = p.n / CH;
    const int cs = blockIdx.x % p.csplit;
    const int b0 = (blockIdx.x / p.csplit) * p.TB;
    const int cpc = (nchunks + p.csplit - 1) / p.csplit;
    const int chunk_begin = cs * cpc, chunk_end = min(nchunks, chunk_begin + cpc);
    __shared__ unsigned sAmax[32];
    __shared__ float sAlpha[32];

    // call-constant inputs before the dependency wait (common.cuh, programmatic dependent launch)
    stage_lists_bytes(sPtr, sOff, p.TB, P, b0, p.list_cap, p.lst_ptr, p.lst_idx, p.gid);
    stage_scalars<float>(sHyp, sDeg, p.TB, P, p.B, b0, p.hyp_k, p.hyp_prev, p.deg, p.gid);
    flag_nonfinite_row<float>(p.hyp_k, P, p.flags);
    pdl_wait();
    pdl_trigger();

    // scale of the fused fp16 split of y_{k+1} (identical in every CTA)
    u64 sc2 = dup2(1.f);
    const bool do_split = p.sp.hi != nullptr;
    if (do_split) {
        float bound = p.V;
        if (p.sp.amax_in) {
            const float amax_alpha = block_max_alpha(p.hyp_k, P, sAlpha);
            bound = fminf(bound, __uint_as_float(__ldg(p.sp.amax_in)) + amax_alpha * p.G);
        }
        const int e = split_exponent(bound);
        sc2 = dup2(pow2_of(e / 2) * pow2_of(e - e / 2));
        if (blockIdx.x == 0 && threadIdx.x == 0) *p.sp.exp = e;
    }
    float amax_f = 0.f;
    u64 nonfinite = 0ull;
    const u64 zero2 = 0ull;
    const bool sums = p.agent_sum != nullptr;
    u64 sq2 = 0ull;
    const float G = p.G, V = p.V, Uc = p.Uc_prev;

    for (int chunk = chunk_begin; chunk < chunk_end; ++chunk) {
        const int i = chunk * CH + lane * 4;
        if (chunk != chunk_begin) __syncthreads();      // rows of the previous chunk are still being read
        for (int bl = 0; bl < p.TB; ++bl) {
            const float* src = p.y + (((unsigned)(b0 + bl) * P) * p.n + i);
            for (int pp = warp; pp < P; pp += nwarps)
                cp_async16(S0 + (unsigned)(bl * P + pp) * ROWB + lane_bytes, src + (unsigned)pp * p.n);
        }
        cp_async_wait_all();
        __syncthreads();

        for (int bl = 0; bl < p.TB; ++bl) {
            const unsigned base = ((unsigned)(b0 + bl) * P) * p.n + i;          // 32-bit element offsets: B*P*n < 2^31 (host-checked)
            const unsigned char* tile_lane = S0 + (size_t)bl * P * ROWB + lane_bytes;
            const int32_t* lptr = sPtr + bl * (P + 1);
            const int32_t* loff = sOff + bl * p.list_cap;
            u64 vs0 = 0ull, vs1 = 0ull;
            for (int pp = warp; pp < P; pp += nwarps) {
                const unsigned off = base + (unsigned)pp * p.n;
                const Q4 av = ldq_stream(p.a + off);
                Q4 Uv = ldq_stream(p.U_in + off);
                if (p.prefetch) {
                    // next row of this warp; from its last row of the chunk, its first row of the next chunk
                    const bool more = pp + nwarps < P;
                    if (more || (p.TB == 1 && chunk + 1 < chunk_end)) {
                        const unsigned nx = more ? off + (unsigned)nwarps * p.n : base + CH + (unsigned)warp * p.n;
                        prefetch_row(p.a + nx, p.prefetch);
                        prefetch_row(p.U_in + nx, p.prefetch);
                    }
                }
                const float4 h4 = *reinterpret_cast<const float4*>(sHyp + pp * 4);     // alpha_k, tau_k, rho_k, eta_{k-1}
                const u64 dg2 = dup2(sDeg[bl * P + pp]);
                const Q4 yv = ldq(tile_lane + pp * ROWB);
                const Q4 dv = p.exact_order ? lap_events(tile_lane, yv, loff, lptr[pp], lptr[pp + 1])
                                            : lap_half(tile_lane, yv, loff, lptr[pp], lptr[pp + 1]);
                // U_k = clamp(U_{k-1} + d_k eta_{k-1})
                const u64 eta2 = dup2(h4.w);
                Uv.a = clamp2(add2(Uv.a, mul2(dv.a, eta2)), Uc);
                Uv.b = clamp2(add2(Uv.b, mul2(dv.b, eta2)), Uc);
                // r_k = a_k' + sign(y) tau + U_k deg + d_k rho, left to right (unfolded_DLASSO.py:73-77)
                const u64 rho2 = dup2(h4.z), alpha2 = dup2(h4.x);
                u64 r0 = add2(av.a, sign2(yv.a, h4.y)), r1 = add2(av.b, sign2(yv.b, h4.y));
                r0 = add2(r0, mul2(Uv.a, dg2));
                r1 = add2(r1, mul2(Uv.b, dg2));
                r0 = add2(r0, mul2(dv.a, rho2));
                r1 = add2(r1, mul2(dv.b, rho2));
                nonfinite = fma2(r0, zero2, nonfinite);            // 0 * x is NaN iff x is Inf / NaN
                nonfinite = fma2(r1, zero2, nonfinite);
                Q4 yn;
                yn.a = clamp2(sub2(yv.a, mul2(alpha2, clamp2(r0, G))), V);
                yn.b = clamp2(sub2(yv.b, mul2(alpha2, clamp2(r1, G))), V);
                stq(p.y_next + off, yn);
                if (p.U_out) stq_stream(p.U_out + off, Uv);
                if (do_split) store_split4(p.sp, off, yn, sc2);
                {
                    float a, b, c, d;
                    upk2(yn.a, a, b);
                    upk2(yn.b, c, d);
                    amax_f = fmaxf(fmaxf(amax_f, fmaxf(fabsf(a), fabsf(b))), fmaxf(fabsf(c), fabsf(d)));
                }
                if (sums) {
                    vs0 = add2(vs0, yn.a);
                    vs1 = add2(vs1, yn.b);
                    sq2 = fma2(yn.a, yn.a, sq2);
                    sq2 = fma2(yn.b, yn.b, sq2);
                }
            }
            if (sums) {
                // sum of this problem's rows over the warps -> agent_sum[b0 + bl][chunk].  The scratch is the FIRST problem's
                // y_k rows 0 .. nwarps-1 (P >= nwarps, host-checked): after the barrier no warp reads them any more -- the
                // problems of a tile are walked in order, and a problem gathers from its own rows only.
                __syncthreads();
                float* red = reinterpret_cast<float*>(S0);
                stq(red + warp * CH + lane * 4, Q4{vs0, vs1});
                __syncthreads();
                if (threadIdx.x < CH) {
                    float a = 0.f;
                    for (int wq = 0; wq < nwarps; ++wq) a += red[wq * CH + threadIdx.x];
                    p.agent_sum[(unsigned)(b0 + bl) * p.n + chunk * CH + threadIdx.x] = a;
                }
            }
        }
    }   // chunk loop
    if (sums) {
        float w = hsum2(sq2);
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
    if (p.sp.amax_out) publish_amax(__float_as_uint(amax_f), p.sp.amax_out, sAmax);
    if (p.flags) {
        float na, nb;
        upk2(nonfinite, na, nb);
        const unsigned bad = __reduce_or_sync(0xffffffffu, (na != 0.f || nb != 0.f) ? 0xFu : 0u);
        if (bad && lane == 0) atomicOr(p.flags, (int)bad);
    }
}

// ------------------------------------------------------------------------------------------------------------------
// backward level k >= 1 of the fused training path (see unfolded.cuh for the recurrence and the meaning of T, C, S)
// ------------------------------------------------------------------------------------------------------------------
template <int NTHR, int MINB>
__global__ void __launch_bounds__(NTHR, MINB)
level_bwd_lean_kernel(const LevelBwdParams<float> p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int CH = 128, ROWB = 512;
    const int P = p.P, R = p.TB * P;
    unsigned char* S0 = smem_raw;                                   // y_k tile
    unsigned char* S1 = S0 + (size_t)R * ROWB;                      // adjoint of 2L y_k
    float* sAcc = reinterpret_cast<float*>(S1 + (size_t)R * ROWB);  // [R][4]: per-row sums of d/d (alpha, tau, rho, eta_prev)
    float* sHyp = sAcc + (size_t)R * 4;
    float* sDeg = sHyp + 4 * P;
    int32_t* sPtr = reinterpret_cast<int32_t*>(sDeg + R);
    int32_t* sOff = sPtr + p.TB * (P + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = NTHR >> 5;
    const int lane_bytes = lane * 16;
    const int nchunks = p.n / CH;
    const int cs = blockIdx.x % p.csplit;
    const int b0 = (blockIdx.x / p.csplit) * p.TB;
    const int cpc = (nchunks + p.csplit - 1) / p.csplit;
    const int chunk_begin = cs * cpc, chunk_end = min(nchunks, chunk_begin + cpc);
    const bool top = p.top != 0;
    __shared__ float sAlpha[32];

    stage_lists_bytes(sPtr, sOff, p.TB, P, b0, p.list_cap, p.lst_ptr, p.lst_idx, p.gid);
    for (int r = threadIdx.x; r < R * 4; r += NTHR) sAcc[r] = 0.f;
    stage_scalars<float>(sHyp, sDeg, p.TB, P, p.B, b0, p.hyp_k, p.hyp_prev, p.deg, p.gid);
    pdl_wait();
    pdl_trigger();

    const float* label = p.label;
    float coef_prev = p.coef_prev;
    if (p.coef_dev) {
        coef_prev = (float)__ldg(p.coef_dev);
        if (coef_prev == 0.f) label = nullptr;
    }
    const u64 coef2 = dup2(coef_prev);

    // scale of the fused fp16 split of gAtAy_k:  |gAtAy| <= max_p alpha_p * max|adj(y_{k+1})|
    u64 sc2 = dup2(1.f);
    const bool do_split = p.sp.hi != nullptr;
    if (do_split) {
        const float amax_alpha = block_max_alpha(p.hyp_k, P, sAlpha);
        const int e = split_exponent(amax_alpha * __uint_as_float(__ldg(p.sp.amax_in)));
        sc2 = dup2(pow2_of(e / 2) * pow2_of(e - e / 2));       // |e| <= 115: a normal float, the product is exact
        if (blockIdx.x == 0 && threadIdx.x == 0) *p.sp.exp = e;
    }
    const float G = p.G, V = p.V, Uc = p.Uc_prev;

    for (int chunk = chunk_begin; chunk < chunk_end; ++chunk) {
        const int i = chunk * CH + lane * 4;
        if (chunk != chunk_begin) __syncthreads();          // the previous chunk's last phase still reads the tiles
        for (int bl = 0; bl < p.TB; ++bl) {
            const float* src = p.y + (((unsigned)(b0 + bl) * P) * p.n + i);
            for (int pp = warp; pp < P; pp += nwarps)
                cp_async16(S0 + (unsigned)(bl * P + pp) * ROWB + lane_bytes, src + (unsigned)pp * p.n);
        }
        cp_async_wait_all();
        __syncthreads();

        for (int bl = 0; bl < p.TB; ++bl) {
            const unsigned base = ((unsigned)(b0 + bl) * P) * p.n + i;
            const unsigned char* tile_lane = S0 + (size_t)bl * P * ROWB + lane_bytes;
            unsigned char* tile1_lane = S1 + (size_t)bl * P * ROWB + lane_bytes;
            const int32_t* lptr = sPtr + bl * (P + 1);
            const int32_t* loff = sOff + bl * p.list_cap;
            Q4 labv{0ull, 0ull};
            if (label) labv = ldq(label + ((unsigned)(b0 + bl) * p.n + i));
            for (int pp = warp; pp < P; pp += nwarps) {
                const unsigned off = base + (unsigned)pp * p.n;
                const Q4 tv = ldq(p.Tb + off);
                const Q4 rv = ldq_stream(p.graw + off);
                const Q4 uv = ldq_stream(p.U_prev + off);
                Q4 cv{0ull, 0ull};
                if (!top) cv = ldq(p.C + off);
                if (p.prefetch) {
                    // the ncu source view of this kernel has 18 % of all stall samples on the first use of this row's
                    // streams (long scoreboard): the gather in between is shorter than an HBM round trip.  Next row of this
                    // warp; from its last row of the chunk, its first row of the next chunk
                    const bool more = pp + nwarps < P;
                    if (more || (p.TB == 1 && chunk + 1 < chunk_end)) {
                        const unsigned nx = more ? off + (unsigned)nwarps * p.n : base + CH + (unsigned)warp * p.n;
                        prefetch_row(p.Tb + nx, p.prefetch);
                        prefetch_row(p.graw + nx, p.prefetch);
                        prefetch_row(p.U_prev + nx, p.prefetch);
                        if (!top) prefetch_row(p.C + nx, p.prefetch);
                    }
                }
                const float4 h4 = *reinterpret_cast<const float4*>(sHyp + pp * 4);     // alpha_k, tau_k, rho_k, eta_{k-1}
                const u64 dg2 = dup2(sDeg[bl * P + pp]);
                const Q4 yv = ldq(tile_lane + pp * ROWB);
                // 2L y_k with the forward level's own rounding (lean::lap_half is its default form): U_k, r_k and every mask
                // below are then bit for bit the forward's, not a re-evaluation that may fall on the other side of a clamp
                const Q4 dw = lap_half(tile_lane, yv, loff, lptr[pp], lptr[pp + 1]);
                const u64 alpha2 = dup2(h4.x), nalpha2 = dup2(-h4.x), tau2 = dup2(h4.y), rho2 = dup2(h4.z), eta2 = dup2(h4.w);
                Q4 o_ga, o_c, o_dir, o_db;
                u64 pa2 = 0ull, pt2 = 0ull, pr2 = 0ull, pe2 = 0ull;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const u64 y = h ? yv.b : yv.a, d = h ? dw.b : dw.a, t = h ? tv.b : tv.a;
                    // r_k = (AtA y - Atb) + sign(y) tau + U_k deg + d_k rho with U_k = clamp(U_{k-1} + d_k eta_{k-1}): the
                    // forward level's operations, so every mask below is the forward's own decision
                    const u64 w = add2(h ? uv.b : uv.a, mul2(d, eta2));
                    const u64 Uk = clamp2(w, Uc);
                    const u64 sg = sign2(y, 1.f);
                    u64 rr = add2(h ? rv.b : rv.a, mul2(sg, tau2));
                    rr = add2(rr, mul2(Uk, dg2));
                    rr = add2(rr, mul2(d, rho2));
                    const u64 g = clamp2(rr, G);
                    const u64 z = sub2(y, mul2(alpha2, g));
                    const u64 zb = mask2(z, V, t);
                    pa2 = fma2(zb, g, pa2);                                  // d alpha_k = -<zb, g> (negated below)
                    const u64 rb = mask2(rr, G, mul2(nalpha2, zb));
                    pt2 = fma2(rb, sg, pt2);
                    pr2 = fma2(rb, d, pr2);
                    const u64 uk = fma2(dg2, rb, h ? cv.b : cv.a);
                    const u64 um = mask2(w, Uc, uk);
                    pe2 = fma2(um, d, pe2);
                    const u64 db = fma2(eta2, um, mul2(rho2, rb));
                    u64 dir = zb;
                    if (label) dir = fma2(coef2, sub2(y, h ? labv.b : labv.a), zb);
                    if (h) { o_ga.b = rb; o_c.b = um; o_db.b = db; o_dir.b = dir; }
                    else { o_ga.a = rb; o_c.a = um; o_db.a = db; o_dir.a = dir; }
                }
                stq(tile1_lane + pp * ROWB, o_db);
                if (do_split) store_split4(p.sp, off, o_ga, sc2);
                else stq(p.ga + off, o_ga);
                stq(p.C + off, o_c);
                stq(p.Tb + off, o_dir);            // + 2L db in the last phase (same thread re-reads it)
                {   // lanes 0/8/16/24 end up with the row's four sums (only this warp touches the row's slots)
                    const float k4 = warp_sum4(-hsum2(pa2), hsum2(pt2), hsum2(pr2), hsum2(pe2), lane);
                    if ((lane & 7) == 0) sAcc[(bl * P + pp) * 4 + (lane >> 3)] += k4;
                }
            }
        }
        __syncthreads();
        for (int bl = 0; bl < p.TB; ++bl) {
            const unsigned base = ((unsigned)(b0 + bl) * P) * p.n + i;
            const unsigned char* tile1_lane = S1 + (size_t)bl * P * ROWB + lane_bytes;
            const int32_t* lptr = sPtr + bl * (P + 1);
            const int32_t* loff = sOff + bl * p.list_cap;
            for (int pp = warp; pp < P; pp += nwarps) {
                const unsigned off = base + (unsigned)pp * p.n;
                Q4 s = ldq(p.Tb + off);                     // issued ahead of the shared-memory gather (an L2 hit: this thread wrote it)
                const Q4 xq = ldq(tile1_lane + pp * ROWB);
                const Q4 lt = lap_adj(tile1_lane, xq, loff, lptr[pp], lptr[pp + 1]);
                s.a = add2(s.a, lt.a);
                s.b = add2(s.b, lt.b);
                stq(p.Tb + off, s);
            }
        }
    }   // chunk loop

    // one partial-sum row per CTA: (d alpha, d tau, d rho, d eta_prev) of every (problem, agent) of the tile
    __syncthreads();
    for (int r = threadIdx.x; r < R * 4; r += NTHR) {
        const int b = b0 + (r >> 2) / P, pp = (r >> 2) % P;
        p.partials[(((unsigned)cs * p.B + b) * P + pp) * 4 + (r & 3)] = sAcc[r];
    }
}

// ------------------------------------------------------------------------------------------------------------------
// backward level k = 0 of the fused training path: the sweep ends here -- nothing is propagated further, only
// d alpha_0, d tau_0, d rho_0 are summed (delta_0 and U_0 are the given initial state: no gather, no dual update, no tile).
// A streaming reduction over five tensors; the generic kernel that served it ran at 42 % of DRAM bandwidth (1.2 ms at
// config 4, round-2 ncu).
// ------------------------------------------------------------------------------------------------------------------
template <int NTHR>
__global__ void __launch_bounds__(NTHR, 4) level_bwd_first_lean_kernel(const LevelBwdParams<float> p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int CH = 128, ROWB = 512;
    const int P = p.P, R = p.TB * P;
    float* sAcc = reinterpret_cast<float*>(smem_raw + (size_t)2 * R * ROWB);     // same layout as level_bwd_lean_kernel (tiles unused)
    float* sHyp = sAcc + (size_t)R * 4;
    float* sDeg = sHyp + 4 * P;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = NTHR >> 5;
    const int nchunks = p.n / CH;
    const int cs = blockIdx.x % p.csplit;
    const int b0 = (blockIdx.x / p.csplit) * p.TB;
    const int cpc = (nchunks + p.csplit - 1) / p.csplit;
    const int chunk_begin = cs * cpc, chunk_end = min(nchunks, chunk_begin + cpc);
    for (int r = threadIdx.x; r < R * 4; r += NTHR) sAcc[r] = 0.f;
    stage_scalars<float>(sHyp, sDeg, p.TB, P, p.B, b0, p.hyp_k, (const float*)nullptr, p.deg, p.gid);
    __syncthreads();
    pdl_wait();
    pdl_trigger();
    const float G = p.G, V = p.V;
    for (int chunk = chunk_begin; chunk < chunk_end; ++chunk) {
        const int i = chunk * CH + lane * 4;
        for (int bl = 0; bl < p.TB; ++bl) {
            const unsigned base = ((unsigned)(b0 + bl) * P) * p.n + i;
            for (int pp = warp; pp < P; pp += nwarps) {
                const unsigned off = base + (unsigned)pp * p.n;
                const Q4 tv = ldq(p.Tb + off);
                const Q4 rv = ldq_stream(p.graw + off);
                const Q4 uv = ldq_stream(p.U_prev + off);          // U_0
                const Q4 dv = ldq_stream(p.d0 + off);              // delta_0
                const Q4 yv = ldq_stream(p.y + off);               // y_0
                const float4 h4 = *reinterpret_cast<const float4*>(sHyp + pp * 4);
                const u64 dg2 = dup2(sDeg[bl * P + pp]);
                const u64 alpha2 = dup2(h4.x), nalpha2 = dup2(-h4.x), tau2 = dup2(h4.y), rho2 = dup2(h4.z);
                u64 pa2 = 0ull, pt2 = 0ull, pr2 = 0ull;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const u64 y = h ? yv.b : yv.a, d = h ? dv.b : dv.a, t = h ? tv.b : tv.a;
                    const u64 sg = sign2(y, 1.f);
                    u64 rr = add2(h ? rv.b : rv.a, mul2(sg, tau2));
                    rr = add2(rr, mul2(h ? uv.b : uv.a, dg2));
                    rr = add2(rr, mul2(d, rho2));
                    const u64 g = clamp2(rr, G);
                    const u64 z = sub2(y, mul2(alpha2, g));
                    const u64 zb = mask2(z, V, t);
                    pa2 = fma2(zb, g, pa2);
                    const u64 rb = mask2(rr, G, mul2(nalpha2, zb));
                    pt2 = fma2(rb, sg, pt2);
                    pr2 = fma2(rb, d, pr2);
                }
                const float k4 = warp_sum4(-hsum2(pa2), hsum2(pt2), hsum2(pr2), 0.f, lane);
                if ((lane & 7) == 0) sAcc[(bl * P + pp) * 4 + (lane >> 3)] += k4;
            }
        }
    }
    __syncthreads();
    for (int r = threadIdx.x; r < R * 4; r += NTHR) {
        const int b = b0 + (r >> 2) / P, pp = (r >> 2) % P;
        p.partials[(((unsigned)cs * p.B + b) * P + pp) * 4 + (r & 3)] = sAcc[r];
    }
}

}  // namespace lean
}  // namespace dadmm

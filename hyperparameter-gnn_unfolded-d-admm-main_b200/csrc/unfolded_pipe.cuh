// unfolded_pipe.cuh -- forward level kernel of the fused fp32 path as a persistent, warp-specialised TMA pipeline.
//
// STATUS: an experiment that LOST, kept selectable (DADMM_FWD_PIPE=1; parity-green under the full GPU suite) because the
// measurement is the point: 1.34 ms per level with one bulk copy per tile row, 1.16 ms with one tiled tensor-map copy per
// tensor, against 0.84 ms for the occupancy-driven kernel of unfolded_lean.cuh (B200, config 4; DESIGN.md section 4,
// finding 3).  Two 77 KB stages per SM put the reads on HBM in bursts, 17 warps hide less gather latency than 40, and
// shared memory has no room for a third stage at P = 50.  The default path does not launch this kernel.
//
// Why (ncu, round 2, profiles/r02_ncu_levels_gen2.txt): the occupancy-driven level kernels (unfolded.cuh, unfolded_lean.cuh)
// run the forward level at 5.3 TB/s alone and 4.9 TB/s inside the power-capped step; their warps spend 7 of every 12
// stalled cycles on the long scoreboard (the row's `a` and `U` loads, issued by the same warp that needs them a few
// hundred cycles later), DRAM is busy 65 % of the time and the issue slots 53 %: a latency-bound mix, not a bandwidth-
// bound one, and halving the instruction count (packed fp32) moved it by 3 %.  Here the loads leave the compute warps:
//
//   producer warp   per tile: waits for a free stage, then issues one 512-byte bulk copy (cp.async.bulk -> UBLKCP) per
//                   tile row and tensor -- y_k, a_k' = AtA y_k - Atb, U_{k-1} -- completing on the stage's mbarrier;
//   consumer warps  wait on the stage's mbarrier, find every operand of a row in shared memory (the y_k tile is the
//                   gather source of 2L y_k as before), compute with the packed fp32 arithmetic of unfolded_lean.cuh
//                   (same operations, same order, same rounding) and store y_{k+1}, U_k and the fp16 split;
//   one CTA per SM  walks a contiguous range of (problem group, 128-unknown chunk) tiles, two stages deep: the copies of
//                   tile i+1 are in flight while tile i computes, whatever the occupancy or the SM clock.
//
// The label-free loss sums (dadmm_loss_sums) are taken from the tile itself: y_{k+1} overwrites the row's slot of the
// `a` tile, and after the tile's rows are done 128 threads per problem add the P rows of their column -- deterministic,
// and no longer restricted to one problem per tile.
#pragma once
#include "contract_tc.cuh"
#include "unfolded_lean.cuh"

namespace dadmm {
namespace pipe {

using namespace lean;

constexpr int kStages = 2;
constexpr int kRowBytes = 512;                  // 128 fp32 unknowns per tile row

__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void consumer_sync(int nthreads) { asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory"); }

// shared-memory bytes of the kernel for a tile of R rows whose problems' lists hold at most `cap` entries each
inline size_t fwd_smem_bytes(int TB, int P, int cap) {
    const size_t R = (size_t)TB * P;
    return 128 /* alignment slack */ + (size_t)kStages * 3 * R * kRowBytes + 64 /* mbarriers */ + (4 * (size_t)P + R) * 4 +
           (size_t)TB * ((size_t)P + 1 + cap) * 4 + 16;
}

// TMAP: one tiled tensor-map copy per tensor and tile ([R rows] x [128 unknowns] box of the [B*P, n] tensor) instead of one
// 512-byte bulk copy per tile row (measured: the copy engine paces 150 small copies per tile at ~40 ns each)
template <int NCONS, bool TMAP>
__global__ void __launch_bounds__((NCONS + 1) * 32, 1)
level_fwd_pipe_kernel(const LevelFwdParams<float> p, const __grid_constant__ CUtensorMap map_y, const __grid_constant__ CUtensorMap map_a,
                      const __grid_constant__ CUtensorMap map_u) {
    extern __shared__ __align__(128) unsigned char smem_dyn[];
    constexpr int CH = 128, ROWB = kRowBytes, NTHR_C = NCONS * 32;
    const int P = p.P, R = p.TB * P;
    const uint32_t tileB = (uint32_t)R * ROWB;
    unsigned char* smem_raw = smem_dyn + ((128u - (tc::smem_u32(smem_dyn) & 127u)) & 127u);     // tensor-map destinations: 128-byte aligned
    unsigned char* stages = smem_raw;                                   // [kStages][y | a | U][R][512]
    unsigned char* after = stages + (size_t)kStages * 3 * tileB;
    const uint32_t bars = tc::smem_u32(after);                          // full[kStages], empty[kStages]
    float* sHyp = reinterpret_cast<float*>(after + 64);
    float* sDeg = sHyp + 4 * P;
    int32_t* sPtr = reinterpret_cast<int32_t*>(sDeg + R);
    int32_t* sOff = sPtr + p.TB * (P + 1);
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (kStages + s); };
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nchunks = p.n / CH;
    const long long tiles = (long long)(p.B / p.TB) * nchunks;
    const long long t_begin = tiles * blockIdx.x / gridDim.x, t_end = tiles * (blockIdx.x + 1) / gridDim.x;
    __shared__ unsigned sAmax[32];
    __shared__ float sAlpha[32];

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) {
            tc::mbar_init(full_bar(s), 1);
            tc::mbar_init(empty_bar(s), NCONS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    flag_nonfinite_row<float>(p.hyp_k, P, p.flags);
    // per-agent scalars of this level: (alpha_k, tau_k, rho_k, eta_{k-1}) -- the same for every tile
    for (int q = threadIdx.x; q < P; q += blockDim.x) {
        sHyp[q * 4 + 0] = __ldg(p.hyp_k + q * 4);
        sHyp[q * 4 + 1] = __ldg(p.hyp_k + q * 4 + 1);
        sHyp[q * 4 + 2] = __ldg(p.hyp_k + q * 4 + 2);
        sHyp[q * 4 + 3] = __ldg(p.hyp_prev + q * 4 + 3);
    }
    __syncthreads();
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
    u64 nonfinite = 0ull, sq2 = 0ull;

    if (warp == NCONS) {
        // ------------------------------------------------------------------ producer: bulk copies of the next tile
        int it = 0;
        for (long long t = t_begin; t < t_end; ++t, ++it) {
            const int s = it % kStages;
            const uint32_t ph = (uint32_t)(it / kStages) & 1u;
            tc::mbar_wait(empty_bar(s), ph ^ 1u);
            if (lane == 0) tc::mbar_expect_tx(full_bar(s), 3u * tileB);
            __syncwarp();
            const int g = (int)(t / nchunks), chunk = (int)(t % nchunks);
            const uint32_t dst0 = tc::smem_u32(stages) + (uint32_t)s * 3u * tileB;
            if constexpr (TMAP) {
                if (lane == 0) {
                    tma_load_2d(dst0, &map_y, full_bar(s), chunk * CH, g * R);
                    tma_load_2d(dst0 + tileB, &map_a, full_bar(s), chunk * CH, g * R);
                    tma_load_2d(dst0 + 2u * tileB, &map_u, full_bar(s), chunk * CH, g * R);
                }
                continue;
            }
            for (int r = lane; r < R; r += 32) {
                const size_t off = ((size_t)g * R + r) * p.n + (size_t)chunk * CH;       // rows of a group are consecutive (b, p) rows
                const uint32_t d = dst0 + (uint32_t)r * ROWB;
                bulk_load(d, p.y + off, ROWB, full_bar(s));
                bulk_load(d + tileB, p.a + off, ROWB, full_bar(s));
                bulk_load(d + 2u * tileB, p.U_in + off, ROWB, full_bar(s));
            }
        }
    } else {
        // ------------------------------------------------------------------ consumers
        const int lane_bytes = lane * 16;
        const bool sums = p.agent_sum != nullptr;
        const float G = p.G, V = p.V, Uc = p.Uc_prev;
        const u64 zero2 = 0ull;
        int g_staged = -1, it = 0;
        for (long long t = t_begin; t < t_end; ++t, ++it) {
            const int s = it % kStages;
            const uint32_t ph = (uint32_t)(it / kStages) & 1u;
            const int g = (int)(t / nchunks), chunk = (int)(t % nchunks);
            const int b0 = g * p.TB;
            if (g != g_staged) {
                // neighbour lists and degrees of the group's problems (constant inputs of the call)
                if (g_staged >= 0) consumer_sync(NTHR_C);          // every consumer is done with the previous group's lists
                for (int bl = 0; bl < p.TB; ++bl) {
                    const int node0 = (p.gid ? __ldg(p.gid + b0 + bl) : 0) * P;
                    const int e0 = __ldg(p.lst_ptr + node0);
                    for (int q = threadIdx.x; q <= P; q += NTHR_C) sPtr[bl * (P + 1) + q] = __ldg(p.lst_ptr + node0 + q) - e0;
                    const int cnt = __ldg(p.lst_ptr + node0 + P) - e0;
                    for (int e = threadIdx.x; e < cnt; e += NTHR_C) sOff[bl * p.list_cap + e] = __ldg(p.lst_idx + e0 + e) << 9;
                    for (int q = threadIdx.x; q < P; q += NTHR_C) sDeg[bl * P + q] = (float)__ldg(p.deg + node0 + q);
                }
                consumer_sync(NTHR_C);
                g_staged = g;
            }
            tc::mbar_wait(full_bar(s), ph);
            unsigned char* Sy = stages + (size_t)s * 3 * tileB;
            unsigned char* Sa = Sy + tileB;
            const unsigned char* Su = Sa + tileB;
            int bl = 0, pp = warp;
            while (pp >= P) { pp -= P; ++bl; }
            for (int r = warp; r < R; r += NCONS) {
                const unsigned off = ((unsigned)g * R + r) * p.n + chunk * CH + lane * 4;     // 32-bit element offsets: B*P*n < 2^31 (host-checked)
                const unsigned char* tile_lane = Sy + (size_t)bl * P * ROWB + lane_bytes;
                const int32_t* lptr = sPtr + bl * (P + 1);
                const Q4 yv = ldq(tile_lane + pp * ROWB);
                const Q4 av = ldq(Sa + (size_t)r * ROWB + lane_bytes);
                Q4 Uv = ldq(Su + (size_t)r * ROWB + lane_bytes);
                const float4 h4 = *reinterpret_cast<const float4*>(sHyp + pp * 4);     // alpha_k, tau_k, rho_k, eta_{k-1}
                const u64 dg2 = dup2(sDeg[bl * P + pp]);
                const Q4 dv = p.exact_order ? lap_events(tile_lane, yv, sOff + bl * p.list_cap, lptr[pp], lptr[pp + 1])
                                            : lap_half(tile_lane, yv, sOff + bl * p.list_cap, lptr[pp], lptr[pp + 1]);
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
                    stq(Sa + (size_t)r * ROWB + lane_bytes, yn);          // this lane read the slot above: y_{k+1} replaces a_k'
                    sq2 = fma2(yn.a, yn.a, sq2);
                    sq2 = fma2(yn.b, yn.b, sq2);
                }
                pp += NCONS;
                while (pp >= P) { pp -= P; ++bl; }
            }
            if (sums) {      // agent_sum[b][chunk columns] = sum over the problem's P rows of y_{k+1}, in row order
                consumer_sync(NTHR_C);
                for (int j = threadIdx.x; j < p.TB * CH; j += NTHR_C) {
                    const int jb = j / CH, col = j % CH;
                    const float* colp = reinterpret_cast<const float*>(Sa) + (size_t)jb * P * CH + col;
                    float acc = 0.f;
                    for (int q = 0; q < P; ++q) acc += colp[q * CH];
                    p.agent_sum[(unsigned)(b0 + jb) * p.n + chunk * CH + col] = acc;
                }
                tc::fence_proxy_async();      // generic-proxy writes to the stage precede the next bulk copy into it
            }
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(empty_bar(s));
        }
    }

    if (p.sq_part) {
        float w = hsum2(sq2);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) w += __shfl_xor_sync(0xffffffffu, w, o);
        __syncthreads();
        if (lane == 0) sAlpha[warp] = w;
        __syncthreads();
        if (threadIdx.x == 0) {
            double tsum = 0;
            for (int wq = 0; wq < NCONS; ++wq) tsum += (double)sAlpha[wq];
            p.sq_part[blockIdx.x] = tsum;
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

}  // namespace pipe
}  // namespace dadmm

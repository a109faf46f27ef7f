// contract_f16.cuh -- fp32-accurate contraction on the kind::f16 tensor path (twice the tf32 rate).
//
//   out[b,p,:] (+)= W_p x[b,p,:]          (same operation as contract_tc.cuh / contract_simt.cuh)
//
// Every fp32 operand value v is represented as  v * 2^e = hi + lo  with hi = rn_fp16(v 2^e) and
// lo = rn_fp16(v 2^e - hi): 22 significant bits, both halves round-to-nearest (the tf32 split of
// contract_tc.cuh truncates hi), and e is ONE power of two per tensor, chosen from the tensor's max |v| so
// that hi stays in fp16's normal range.  A global scale commutes with the contraction, so the result is
// exact up to the dropped lo*lo term (2^-22) -- measured 7e-8 relative vs fp64 for the products, i.e. the fp32
// rounding of the accumulation dominates.  Three MMAs per k-step (lo*hi, hi*lo, hi*hi) at the f16 rate cost
// half of the 3xTF32 scheme, and because the operands arrive in global memory already split
// (split_f16_kernel; W once per operator, x by its producer), the GEMM kernel has no conversion stage:
// TMA -> shared memory -> tcgen05.mma.
//
// Accumulation is chunked (64 or 128 k per tensor-core partial sum, fp32 RN adds in registers; see KB_PER_CHUNK)
// because the TMEM accumulator truncates.  CTA pairs (cta_group::2), 256 operator rows x NT batch columns per pair
// (NT = 256, or 128 for small batches: Geo<NT>).  The same kernel serves the single-stage contraction AtA y and both
// stages of the two-stage form A^T (A y - b): the first stage writes its result straight as the next stage's fp16
// split (Params::t_hi), optionally seeded with -b; the second (or only) stage writes fp32, optionally `+=` or `- sub`.
#pragma once
#include <cuda_fp16.h>

#include "contract_tc.cuh"

namespace dadmm {
namespace f16 {

#ifndef DADMM_F16_BK64
#define DADMM_F16_BK64 1        // round-1 sweep: 64-k blocks (SWIZZLE_128B, 3 stages) 1.05 ms vs 32-k blocks (SWIZZLE_64B, 6 stages) 1.12 ms
#endif
#if DADMM_F16_BK64
constexpr int BKE = 64;                          // k elements per k-block (128-byte fp16 rows, SWIZZLE_128B)
#else
constexpr int BKE = 32;                          // k elements per k-block (64-byte fp16 rows, SWIZZLE_64B)
#endif
// Tile geometry for a batch tile of NT columns per CTA pair (256: the default; 128: small batches, where 256-wide tiles
// leave the last wave mostly empty -- 512 problems per GPU give 100 first-stage tiles for 74 CTA pairs).  Each CTA holds
// its 128 operator rows and NT/2 batch rows per k-block; TMEM takes 512/NT partial-sum buffers.
template <int NT>
struct Geo {
    static constexpr int XROWS = NT / 2;
    static constexpr int WTILE = 128 * BKE * 2;              // bytes: 128 rows x (64 | 128) B
    static constexpr int XTILE = XROWS * BKE * 2;
    static constexpr int STAGE = 2 * WTILE + 2 * XTILE;      // W_hi | X_hi | W_lo | X_lo
    static constexpr int STAGES = DADMM_F16_BK64 ? (NT == 256 ? 3 : 4) : (NT == 256 ? 6 : 8);
    static constexpr int SMEM = STAGES * STAGE + 1024 + 256;
    static constexpr int NBUF = 512 / NT;
    static constexpr int COLS = NT / 2;                      // accumulator columns per epilogue thread
    static constexpr uint32_t IDESC = (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(NT >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
};
constexpr int THREADS = 384;                     // warps 0-3: TMA, MMA, (2 idle); warps 4-11: accumulate + store
constexpr int EPI_WARP0 = 4;
#ifndef DADMM_F16_KB_PER_CHUNK
#define DADMM_F16_KB_PER_CHUNK (DADMM_F16_BK64 ? 2 : 4)
#endif
// k-blocks per tensor-core partial sum (Params::kbc; kb_per_chunk(): 2 x 64 = 128 k for long contractions, 64 k for
// short ones).  The TMEM accumulator truncates on
// every accumulation, so what matters is how many FULL-MAGNITUDE accumulations a partial sum sees: inside a k-block
// the two correction products (lo*hi, hi*lo: 2^-11 of the main term) of all k-steps are issued before the hi*hi
// products, i.e. into a still-small accumulator where their truncation is negligible.  B200, cfg4 shapes:
// 64-k chunks 0.98 ms per two-stage contraction, 128-k 0.86 ms, 256-k 0.82 ms (the 128 KB TMEM drain per chunk,
// not the MMA, paces the short chunks).
constexpr int KB_PER_CHUNK = DADMM_F16_KB_PER_CHUNK;
// (pairing two TMEM loads per tcgen05.wait::ld was measured and changes nothing: the drain is not latency-paced)

// ---------------------------------------------------------------------------------------------------
// operand preparation
// ---------------------------------------------------------------------------------------------------
// max |x| over a tensor as raw float bits (non-negative floats order like unsigned ints)
__global__ void __launch_bounds__(256) amax_kernel(const float* __restrict__ x, long long n, unsigned* __restrict__ out) {
    pdl_wait();
    pdl_trigger();
    unsigned m = 0;
    const long long stride = (long long)gridDim.x * blockDim.x * 4;
    for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
        if (i + 3 < n) {
            const float4 v = *reinterpret_cast<const float4*>(x + i);
            m = max(max(m, __float_as_uint(fabsf(v.x))), max(__float_as_uint(fabsf(v.y)), max(__float_as_uint(fabsf(v.z)), __float_as_uint(fabsf(v.w)))));
        } else {
            for (long long j = i; j < n; ++j) m = max(m, __float_as_uint(fabsf(x[j])));
        }
    }
    m = __reduce_max_sync(0xffffffffu, m);
    if ((threadIdx.x & 31) == 0 && m) atomicMax(out, m);
}

// exponent e with max|v| * 2^e in [2^13, 2^14): two bits of head-room below fp16's 65504
__device__ __forceinline__ int scale_exponent(unsigned amax_bits) {
    const int ea = (int)((amax_bits >> 23) & 0xFF) - 127;          // floor(log2(amax)); amax == 0 or denormal -> -127
    if (ea <= -100 || ea >= 128) return 0;                          // all zeros / non-finite: leave unscaled
    return 13 - ea;
}
__device__ __forceinline__ float pow2f(int e) { return __uint_as_float((unsigned)(e + 127) << 23); }

// x [rows][n] fp32 (row stride ld) -> hi, lo [rows][n_pad] fp16 with the tensor's scale; writes the exponent and,
// when l1_out is given, max_r sum_i |x[r][i]| as float bits (the operator bound of a two-stage contraction)
__global__ void __launch_bounds__(256) split_f16_kernel(const float* __restrict__ x, long long rows, int n, long long ld,
                                                        int n_pad, const unsigned* __restrict__ amax_bits,
                                                        __half* __restrict__ hi, __half* __restrict__ lo, int* __restrict__ exp_out,
                                                        unsigned* __restrict__ l1_out) {
    pdl_wait();
    pdl_trigger();
    const int e = scale_exponent(*amax_bits);
    // 2^e can exceed the float range for tiny tensors; apply it in two factors
    const float s1 = pow2f(e / 2), s2 = pow2f(e - e / 2);
    if (blockIdx.x == 0 && threadIdx.x == 0) *exp_out = e;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    float l1max = 0.f;
    for (long long r = (long long)blockIdx.x * nw + warp; r < rows; r += (long long)gridDim.x * nw) {
        const float* xr = x + r * ld;
        __half* hr = hi + r * n_pad;
        __half* lr = lo + r * n_pad;
        float l1 = 0.f;
        for (int i = lane * 2; i < n_pad; i += 64) {
            float v0 = (i < n) ? xr[i] : 0.f, v1 = (i + 1 < n) ? xr[i + 1] : 0.f;
            l1 += fabsf(v0) + fabsf(v1);
            v0 = v0 * s1 * s2;
            v1 = v1 * s1 * s2;
            const __half h0 = __float2half_rn(v0), h1 = __float2half_rn(v1);
            const __half l0 = __float2half_rn(v0 - __half2float(h0)), l1h = __float2half_rn(v1 - __half2float(h1));
            *reinterpret_cast<__half2*>(hr + i) = __halves2half2(h0, h1);
            *reinterpret_cast<__half2*>(lr + i) = __halves2half2(l0, l1h);
        }
        if (l1_out) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) l1 += __shfl_xor_sync(0xffffffffu, l1, o);
            l1max = fmaxf(l1max, l1);
        }
    }
    if (l1_out && lane == 0 && l1max > 0.f) atomicMax(l1_out, __float_as_uint(l1max));
}

// ---------------------------------------------------------------------------------------------------
// GEMM kernel
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_load_3d_pair(uint32_t dst, const CUtensorMap* map, uint32_t leader_bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"(map), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
// K-major shared-memory matrix descriptor for this kernel's tile geometry
__device__ __forceinline__ uint64_t tile_desc(uint32_t saddr) {
#if DADMM_F16_BK64
    // SWIZZLE_128B: rows of 128 bytes, 8-row swizzle atoms 1024 bytes apart
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
#else
    return tc::umma_desc(saddr);
#endif
}
__device__ __forceinline__ void umma_f16_pair(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate, uint32_t idesc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}

struct Params {
    int B, P, n_out, n_in;
    float* out;
    long long o_sb;
    int accumulate;
    int m_tiles, n_tiles, k_blocks, total_tiles;
    const int *exp_w, *exp_x;      // device: scale exponents of the two operands
    unsigned* amax_out;            // device, optional: running max |out| bits (feeds the next operand split's bound)
    const float* sub;              // device, optional [B,P,n_out] laid out like `out`: out = W x - sub (the Atb term)
    int fast;                      // 1: flagged reduced-precision mode -- hi*hi only (fp16 operands, 2^-11), one MMA per k-step
    int kbc;                       // k-blocks per tensor-core partial sum
    // first stage of a two-stage contraction  out = F2 (F1 x):  t = F1 x leaves the kernel as the fp16 split the
    // second stage consumes ([B*P][t_ld] hi and lo), scaled with an exponent every thread derives from the rigorous
    // bound |t| <= max|x| * max_row ||F1 row||_1  (t_hi == nullptr: plain fp32 output)
    __half *t_hi, *t_lo;
    int t_ld;
    int* t_exp;                    // device: exponent of the t split (written by one thread)
    const unsigned* w_l1;          // device: max row L1 norm of the operator, float bits
    const unsigned* sub_amax;      // device, with t_hi && sub: max |sub| bits (joins the bound of t = F1 x - sub)
    int seed_prefetch;             // 1: L2 prefetch of the next tile's seed values (`+=` / `- sub` epilogues)
};

template <int NT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
contract_f16_kernel(const __grid_constant__ CUtensorMap map_wh, const __grid_constant__ CUtensorMap map_wl,
                    const __grid_constant__ CUtensorMap map_xh, const __grid_constant__ CUtensorMap map_xl, const Params p) {
    using namespace tc;
    using G = Geo<NT>;
    constexpr int STAGES = G::STAGES, STAGE = G::STAGE, WTILE = G::WTILE, XTILE = G::XTILE, NBUF = G::NBUF;
    constexpr int COLS_PER_THREAD = G::COLS;
    extern __shared__ unsigned char smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bars = base + STAGES * STAGE;
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (STAGES + s); };
    auto tfull_bar = [&](int a) { return bars + 8u * (2 * STAGES + a); };
    auto tempty_bar = [&](int a) { return bars + 8u * (2 * STAGES + NBUF + a); };
    const uint32_t tmem_slot = bars + 8u * (2 * STAGES + 2 * NBUF);
    auto stage_base = [&](int s) { return base + (uint32_t)s * STAGE; };

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full_bar(s), 1);       // leader: one expect_tx arrival, bytes from both CTAs' TMA
            mbar_init(empty_bar(s), 1);
        }
        for (int a = 0; a < NBUF; ++a) {
            mbar_init(tfull_bar(a), 1);
            mbar_init(tempty_bar(a), 16);    // 8 accumulate warps per CTA x 2 CTAs (leader only)
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    cluster_sync_all();
    tcgen05_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    // barrier / TMEM set-up above touched no global memory: it overlaps the tail of the preceding kernel (common.cuh)
    pdl_wait();
    pdl_trigger();

    const int tiles_per_agent = p.m_tiles * p.n_tiles;
    const int n_chunks = (p.k_blocks + p.kbc - 1) / p.kbc;

    if (warp < EPI_WARP0) {
        reg_dec<40>();
        if (warp == 0 && lane == 0) {
            // -------------------------------------------------------------- TMA producer
            asm volatile("prefetch.tensormap [%0];" ::"l"(&map_wh) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&map_wl) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&map_xh) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&map_xl) : "memory");
            int stage = 0;
            uint32_t phase = 0;
            for (int t = cluster_id; t < p.total_tiles; t += num_clusters) {
                const int ag = t / tiles_per_agent, r = t % tiles_per_agent;
                const int i0 = (r % p.m_tiles) * 256 + (int)rank * 128, b0 = (r / p.m_tiles) * NT + (int)rank * G::XROWS;
                for (int kb = 0; kb < p.k_blocks; ++kb) {
                    mbar_wait(empty_bar(stage), phase ^ 1u);
                    const uint32_t lead = full_bar(stage) & 0xFEFFFFFFu;     // same offset in the pair's leader CTA
                    const uint32_t sb = stage_base(stage);
#if defined(DADMM_F16_EXPERIMENT) && DADMM_F16_EXPERIMENT == 2
                    if (rank == 0) mbar_expect_tx(full_bar(stage), STAGE);      // timing experiment: W tiles only
                    tma_load_3d_pair(sb, &map_wh, lead, kb * BKE, i0, ag);
                    tma_load_3d_pair(sb + WTILE + XTILE, &map_wl, lead, kb * BKE, i0, ag);
#elif defined(DADMM_F16_EXPERIMENT) && DADMM_F16_EXPERIMENT == 3
                    if (rank == 0) mbar_arrive(full_bar(stage));                // timing experiment: no TMA at all
#else
                    if (rank == 0) mbar_expect_tx(full_bar(stage), p.fast ? STAGE : 2 * STAGE);
                    tma_load_3d_pair(sb, &map_wh, lead, kb * BKE, i0, ag);
                    tma_load_3d_pair(sb + WTILE, &map_xh, lead, kb * BKE, ag, b0);
                    if (!p.fast) {
                        tma_load_3d_pair(sb + WTILE + XTILE, &map_wl, lead, kb * BKE, i0, ag);
                        tma_load_3d_pair(sb + 2 * WTILE + XTILE, &map_xl, lead, kb * BKE, ag, b0);
                    }
#endif
                    if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                }
            }
        } else if (warp == 1 && lane == 0 && rank == 0) {
            // -------------------------------------------------------------- MMA issuer (leader CTA)
            int stage = 0;
            uint32_t phase = 0;
            int ci = 0;
            for (int t = cluster_id; t < p.total_tiles; t += num_clusters) {
                for (int ch = 0; ch < n_chunks; ++ch, ++ci) {
                    const int buf = ci % NBUF;
                    mbar_wait(tempty_bar(buf), ((uint32_t)(ci / NBUF) & 1u) ^ 1u);
                    tcgen05_fence_after();
                    const uint32_t d_tmem = tmem_base + (uint32_t)buf * NT;
                    const int kb_end = min(p.k_blocks, (ch + 1) * p.kbc);
                    for (int kb = ch * p.kbc; kb < kb_end; ++kb) {
                        mbar_wait(full_bar(stage), phase);
                        tcgen05_fence_after();
                        const uint32_t sa = stage_base(stage), sb = sa + WTILE, sa_lo = sb + XTILE, sb_lo = sa_lo + WTILE;
                        const bool fresh = kb == ch * p.kbc;           // first k-block of the partial sum: accumulator starts at 0
                        if (!p.fast) {
                            // corrections first (see KB_PER_CHUNK): 2 x BKE/16 small accumulations, then BKE/16 full-size ones
#pragma unroll
                            for (int ks = 0; ks < BKE / 16; ++ks) {
                                const uint32_t koff = ks * 32;   // 16 fp16 = 32 bytes along K inside the swizzled row
                                umma_f16_pair(d_tmem, tile_desc(sa_lo + koff), tile_desc(sb + koff), (!fresh || ks != 0) ? 1u : 0u, G::IDESC);
                                umma_f16_pair(d_tmem, tile_desc(sa + koff), tile_desc(sb_lo + koff), 1u, G::IDESC);
                            }
#pragma unroll
                            for (int ks = 0; ks < BKE / 16; ++ks) {
                                const uint32_t koff = ks * 32;
                                umma_f16_pair(d_tmem, tile_desc(sa + koff), tile_desc(sb + koff), 1u, G::IDESC);
                            }
                        } else {
#pragma unroll
                            for (int ks = 0; ks < BKE / 16; ++ks) {
                                const uint32_t koff = ks * 32;
                                umma_f16_pair(d_tmem, tile_desc(sa + koff), tile_desc(sb + koff), (!fresh || ks != 0) ? 1u : 0u, G::IDESC);
                            }
                        }
                        umma_commit_pair(empty_bar(stage));
                        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                    }
                    umma_commit_pair(tfull_bar(buf));
                }
            }
        }
    } else {
        // ------------------------------------------------------------------ accumulate + store (256 threads per CTA)
        reg_inc<216>();
        const int q = warp & 3;
        const int h = (warp - EPI_WARP0) >> 2;
        const int es = -(__ldg(p.exp_w) + __ldg(p.exp_x));           // undo both operand scales (exact: powers of two)
        const float unscale = pow2f(es / 2), unscale2 = pow2f(es - es / 2);
        const float rescale = pow2f(-es / 2), rescale2 = pow2f(-es + es / 2);
        float tscale = 1.f, tscale2 = 1.f;
        if (p.t_hi) {
            // |x| 2^exp_x < 2^14 by construction of every split, so |t| < 2^(14 - exp_x) * L1(F1); 1 % covers the fp32
            // rounding of the L1 sum
            const int ex = max(-112, min(126, 14 - __ldg(p.exp_x)));
            float bound = pow2f(ex) * (__uint_as_float(__ldg(p.w_l1)) * 1.01f);
            if (p.sub) bound += __uint_as_float(__ldg(p.sub_amax));
            const int et = scale_exponent(__float_as_uint(bound));
            const int ts = et + es;                                   // accumulators hold t * 2^(exp_w + exp_x)
            tscale = pow2f(ts / 2);
            tscale2 = pow2f(ts - ts / 2);
            if (blockIdx.x == 0 && threadIdx.x == EPI_WARP0 * 32) *p.t_exp = et;
        }
        int ci = 0;
        for (int t = cluster_id; t < p.total_tiles; t += num_clusters) {
            const int ag = t / tiles_per_agent, r = t % tiles_per_agent;
            const int i0 = (r % p.m_tiles) * 256 + (int)rank * 128, b0 = (r / p.m_tiles) * NT + h * COLS_PER_THREAD;
            float acc[COLS_PER_THREAD];
            const int i = i0 + q * 32 + lane;
            float* orow = p.out ? p.out + (long long)ag * p.n_out + i : nullptr;      // first stage of a two-stage call: no fp32 output
            const bool seeded = p.accumulate || p.sub != nullptr;
            // full tile: every (i, b) of this thread is in range -- no per-element predicates, addresses by pointer stepping
            // (ncu, round 1: the guarded 64-bit index form cost ~16 instructions per load and ~20 per store and made the
            // short-K second stage of a two-stage contraction issue-bound in its epilogue)
            const bool full = (i0 + 128 <= p.n_out) && (b0 + COLS_PER_THREAD <= p.B);
            if (seeded) {
                // out += W x  /  out = W x - sub: the accumulators start from the old values (or -sub).  Only the raw
                // loads are issued here -- 128 independent LDGs per thread that fly while the tensor core works on the
                // first chunk; the (exact, power-of-two) move into the scaled domain is folded into the first
                // accumulation below.  (With the scaling attached to each load, ptxas serialised the loads in small
                // scoreboard batches: +0.7 ms per launch, round-1 measurement.)
                const float* srow = p.accumulate ? orow : p.sub + (long long)ag * p.n_out + i;
                if (full) {
                    const float* sp = srow + (long long)b0 * p.o_sb;
#pragma unroll
                    for (int c = 0; c < COLS_PER_THREAD; ++c, sp += p.o_sb) acc[c] = __ldcs(sp);
                    // The seeds are needed at the first partial sum of the tile, one HBM round trip after these loads issue:
                    // a seeded second stage took 0.42 ms against 0.28 ms unseeded (round 2, ncu: tensor pipe 43 % against
                    // 69 %).  The accumulators hold the whole register budget, so the NEXT tile's seeds are pulled into L2
                    // with prefetches instead (no registers; a tile, ~10 us, ahead; `+=` epilogues only, see launch()).
                    if (p.seed_prefetch) {
                        const int tn = t + num_clusters;
                        if (tn < p.total_tiles) {
                            const int agn = tn / tiles_per_agent, rn = tn % tiles_per_agent;
                            const int i0n = (rn % p.m_tiles) * 256 + (int)rank * 128, b0n = (rn / p.m_tiles) * NT + h * COLS_PER_THREAD;
                            if (i0n + 128 <= p.n_out && b0n + COLS_PER_THREAD <= p.B) {
                                const float* pn = (p.accumulate ? p.out : p.sub) + (long long)agn * p.n_out + (i0n + q * 32 + lane) +
                                                  (long long)b0n * p.o_sb;
#pragma unroll 8
                                for (int c = 0; c < COLS_PER_THREAD; ++c, pn += p.o_sb) asm volatile("prefetch.global.L2 [%0];" ::"l"(pn));
                            }
                        }
                    }
                } else {
#pragma unroll
                    for (int c = 0; c < COLS_PER_THREAD; ++c) {
                        const int b = b0 + c;
                        acc[c] = (i < p.n_out && b < p.B) ? __ldcs(srow + (long long)b * p.o_sb) : 0.0f;
                    }
                }
            }
            const float seed1 = p.accumulate ? rescale : -rescale;
            for (int ch = 0; ch < n_chunks; ++ch, ++ci) {
                const int buf = ci % NBUF;
                mbar_wait(tfull_bar(buf), (uint32_t)(ci / NBUF) & 1u);
                tcgen05_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * NT + h * COLS_PER_THREAD);
#pragma unroll
                for (int j = 0; j < COLS_PER_THREAD / 32; ++j) {
                    uint32_t v[32];
                    tmem_ld32(taddr + 32u * j, v);
                    tmem_ld_wait();
                    if (j == COLS_PER_THREAD / 32 - 1) {
                        // the buffer is free as soon as its last columns sit in registers: release it before the adds
                        // (the MMA of the next-but-one partial sum waits on this arrival)
                        tcgen05_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive_cluster(tempty_bar(buf), 0);
                    }
                    if (ch == 0 && !seeded) {
#pragma unroll
                        for (int c = 0; c < 32; ++c) acc[32 * j + c] = __uint_as_float(v[c]);
                    } else if (ch == 0) {
#pragma unroll
                        for (int c = 0; c < 32; ++c) acc[32 * j + c] = (acc[32 * j + c] * seed1) * rescale2 + __uint_as_float(v[c]);
                    } else {
#pragma unroll
                        for (int c = 0; c < 32; ++c) acc[32 * j + c] += __uint_as_float(v[c]);
                    }
                }
            }
            unsigned amax_bits = 0;
            if (p.t_hi) {
                const long long tstep = (long long)p.P * p.t_ld;
                if (full) {
                    long long o = ((long long)b0 * p.P + ag) * p.t_ld + i;
#pragma unroll
                    for (int c = 0; c < COLS_PER_THREAD; ++c, o += tstep) {
                        const float val = acc[c] * tscale * tscale2;             // t * 2^e_t, exact
                        const __half hv = __float2half_rn(val);
                        p.t_hi[o] = hv;
                        p.t_lo[o] = __float2half_rn(val - __half2float(hv));
                        if ((c & 15) == 15) __syncwarp();                         // scheduling fence: keeps ptxas from batching all conversions (spills)
                    }
                } else if (i < p.n_out) {
#pragma unroll
                    for (int c = 0; c < COLS_PER_THREAD; ++c) {
                        const int b = b0 + c;
                        if (b < p.B) {
                            const float val = acc[c] * tscale * tscale2;
                            const __half hv = __float2half_rn(val);
                            const long long o = ((long long)b * p.P + ag) * p.t_ld + i;
                            p.t_hi[o] = hv;
                            p.t_lo[o] = __float2half_rn(val - __half2float(hv));
                        }
                    }
                }
            } else if (full) {
                float* op = orow + (long long)b0 * p.o_sb;
                float amax = 0.f;
#pragma unroll
                for (int c = 0; c < COLS_PER_THREAD; ++c, op += p.o_sb) {
                    const float val = acc[c] * unscale * unscale2;
                    *op = val;
                    amax = fmaxf(amax, fabsf(val));
                }
                amax_bits = __float_as_uint(amax);
            } else if (i < p.n_out) {
#pragma unroll
                for (int c = 0; c < COLS_PER_THREAD; ++c) {
                    const int b = b0 + c;
                    if (b < p.B) {
                        const float val = acc[c] * unscale * unscale2;
                        orow[(long long)b * p.o_sb] = val;
                        amax_bits = max(amax_bits, __float_as_uint(fabsf(val)));
                    }
                }
            }
            if (p.amax_out) {
                amax_bits = __reduce_max_sync(0xffffffffu, amax_bits);
                if (lane == 0 && amax_bits > *reinterpret_cast<volatile unsigned*>(p.amax_out)) atomicMax(p.amax_out, amax_bits);
            }
        }
    }

    tcgen05_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 1) {
        __syncwarp();
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
inline int pad8(int n) { return (n + 7) / 8 * 8; }

// split layout of a tensor with `rows` rows of n values: [scalars 256 B | hi rows*n_pad fp16 | lo rows*n_pad fp16]
struct Split {
    unsigned* amax;
    int* exp;
    unsigned* l1;             // max row L1 norm (float bits), filled when the split is taken with want_l1
    __half *hi, *lo;
};
inline size_t split_bytes(long long rows, int n) { return 256 + ((size_t)rows * pad8(n) * 2 * 2 + 255) / 256 * 256; }
inline Split split_view(void* base, long long rows, int n) {
    char* c = (char*)base;
    Split s;
    s.amax = (unsigned*)c;
    s.exp = (int*)(c + 16);
    s.l1 = (unsigned*)(c + 32);
    s.hi = (__half*)(c + 256);
    s.lo = s.hi + (size_t)rows * pad8(n);
    return s;
}

// x [rows][n] (row stride ld) -> split buffer (amax pass + split pass)
inline int split_tensor(const float* x, long long rows, int n, long long ld, void* buf, cudaStream_t s, bool want_l1 = false) {
    Split v = split_view(buf, rows, n);
    DADMM_CUDA(cudaMemsetAsync(v.amax, 0, 64, s));
    ProfScope prof(PROF_SPLIT, s);
    if (ld == n) {
        const long long tot = rows * n;
        const int nblk = (int)std::min<long long>(148 * 8, ceil_div64(tot, 1024));
        DADMM_CUDA(launch_chain(amax_kernel, dim3(nblk), dim3(256), 0, s, x, tot, v.amax));
    } else {
        for (long long r = 0; r < rows; ++r) amax_kernel<<<1, 256, 0, s>>>(x + r * ld, n, v.amax);   // strided rows: rare
    }
    DADMM_LAUNCHED();
    const int nblk = (int)std::min<long long>(148 * 8, ceil_div64(rows, 8));
    DADMM_CUDA(launch_chain(split_f16_kernel, dim3(nblk), dim3(256), 0, s, x, rows, n, ld, pad8(n), (const unsigned*)v.amax, v.hi, v.lo,
                            v.exp, want_l1 ? v.l1 : (unsigned*)nullptr));
    DADMM_LAUNCHED();
    return 0;
}

// k-blocks per tensor-core partial sum for a contraction over n_in: the longest chunk whose rounding error stays
// below that of the exact-FMA kernel at the same n_in (rel-L2 vs fp64, random data, B200: FMA loop 2.0e-7 at 256,
// 4.1e-7 at 1024, 5.7e-7 at 2048; this kernel 1.4e-7 with 64-k chunks, 3.7e-7 with 128-k, 8.2e-7 with 256-k).
// DADMM_F16_KBC=<n> overrides it at run time (accuracy / speed sweeps).
inline int kb_per_chunk(int n_in, int wanted = 0) {
    static const int forced = [] {
        const char* e = getenv("DADMM_F16_KBC");
        const int x = e ? atoi(e) : 0;
        return (x >= 1 && x <= 64) ? x : 0;
    }();
    if (forced) return forced;
    if (wanted > 0) return wanted;
    // (round 2 tried 128-k partial sums below n = 1024 as well -- configs[2], n = 256: 11.1 -> 10.5 ms per step -- and took it
    // back: at 3.7e-7 per contraction against the FMA loop's 2.0e-7 at K = 256 the tensor-core trajectory is no longer
    // "no worse than the FMA path" in the expanding regime, which test_unfolded_tc_vs_simt_vs_fp64_oracle asserts)
    return n_in >= 1024 ? KB_PER_CHUNK : 1;
}

// max |x| bits of a contiguous tensor into *out (zeroed here)
inline int amax_tensor(const float* x, long long tot, unsigned* out, cudaStream_t s) {
    DADMM_CUDA(cudaMemsetAsync(out, 0, 4, s));
    ProfScope prof(PROF_SPLIT, s);
    const int nblk = (int)std::min<long long>(148 * 8, ceil_div64(tot, 1024));
    DADMM_CUDA(launch_chain(amax_kernel, dim3(nblk), dim3(256), 0, s, x, tot, out));
    DADMM_LAUNCHED();
    return 0;
}

inline bool dims_supported(int B, int P, int n_out, int n_in) { return B >= 128 && n_out > 128 && n_in >= BKE && P >= 1; }

inline int encode3(tc::EncodeTiledFn enc, CUtensorMap* m, const void* ptr, cuuint64_t d0, cuuint64_t d1, cuuint64_t d2,
                   cuuint64_t s1, cuuint64_t s2, cuuint32_t b0, cuuint32_t b1, cuuint32_t b2) {
    cuuint64_t dims[3] = {d0, d1, d2}, strides[2] = {s1, s2};
    cuuint32_t box[3] = {b0, b1, b2}, es[3] = {1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, (void*)ptr, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     DADMM_F16_BK64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) DADMM_FAIL(-4, "cuTensorMapEncodeTiled(f16) failed: %d", (int)r);
    return 0;
}

// out[b,p,:] (+)= W_p x[b,p,:] from prepared operands: wprep = split of W viewed as [P*n_out][n_in],
// xprep = split of x viewed as [B*P][n_in]
// tprep != nullptr: first stage of a two-stage contraction -- the result leaves as the split buffer `tprep`
// ([B*P][n_out], wprep must have been split with want_l1) and `out` is not written
// batch-tile width: 256 columns per CTA pair unless that leaves too few waves of tiles (DADMM_F16_NT=128|256 forces one,
// DADMM_F16_NT_WAVES the threshold).  Round 1 drew the line at three waves.  Round 2 measured, at 5.4 waves of 256-wide tiles
// (interleaved A/B, profiles/r02_tile_width.txt): the second stage (contraction length m = 256, the tile is mostly seeded
// epilogue) is 9 % FASTER on 128-wide tiles (2.90 -> 2.65 ms at 512 problems), the first stage (contraction length n = 1024,
// the tile is mostly main loop and the wider tile halves the operator re-reads) 5 % SLOWER (8.66 -> 9.17 ms at 2048 problems).
// So short contractions switch at six waves, long ones at three.
inline int batch_tile(int B, int P, int n_in, int n_out, int clusters) {
    static const int forced = [] {
        const char* e = getenv("DADMM_F16_NT");
        const int x = e ? atoi(e) : 0;
        return (x == 128 || x == 256) ? x : 0;
    }();
    if (forced) return forced;
    static const int waves = [] {
        const char* e = getenv("DADMM_F16_NT_WAVES");
        const int x = e ? atoi(e) : 0;
        return x > 0 ? x : 0;
    }();
    const int w = waves ? waves : (n_in <= 256 ? 6 : 3);
    const long long tiles256 = (long long)P * ceil_div(n_out, 256) * ceil_div(B, 256);
    return tiles256 < (long long)w * clusters ? 128 : 256;
}

template <int NT>
inline int launch_nt(Params& p, const Split& x, int n_in, int B, int P, tc::EncodeTiledFn enc, const CUtensorMap& mwh,
                     const CUtensorMap& mwl, int num_sms, bool stage1, cudaStream_t s) {
    using G = Geo<NT>;
    const cuuint64_t np = pad8(n_in);
    CUtensorMap mxh, mxl;
    if (int e = encode3(enc, &mxh, x.hi, n_in, P, B, np * 2, (cuuint64_t)P * np * 2, BKE, 1, G::XROWS)) return e;
    if (int e = encode3(enc, &mxl, x.lo, n_in, P, B, np * 2, (cuuint64_t)P * np * 2, BKE, 1, G::XROWS)) return e;
    p.n_tiles = ceil_div(B, NT);
    p.total_tiles = P * p.m_tiles * p.n_tiles;
    static PerDeviceOnce attr_once;
    if (attr_once.first())
        DADMM_CUDA(cudaFuncSetAttribute(contract_f16_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, G::SMEM));
    const int clusters = std::min(num_sms / 2, p.total_tiles);
    ProfScope prof(stage1 ? PROF_CONTRACT_STAGE1 : PROF_CONTRACT_TC, s);
    DADMM_CUDA(launch_chain(contract_f16_kernel<NT>, dim3(2 * clusters), dim3(THREADS), (size_t)G::SMEM, s, mwh, mwl, mxh, mxl, p));
    DADMM_LAUNCHED();
    return 0;
}

inline int launch(int B, int P, int n_out, int n_in, void* wprep, void* xprep, float* out, int64_t o_sb, int accumulate,
                  cudaStream_t s, unsigned* amax_out = nullptr, const float* sub = nullptr, int fast = 0, void* tprep = nullptr,
                  int kbc = 0, const unsigned* sub_amax = nullptr) {
    tc::EncodeTiledFn enc = tc::encode_fn();
    if (!enc) DADMM_FAIL(-4, "cuTensorMapEncodeTiled unavailable");
    const Split w = split_view(wprep, (long long)P * n_out, n_in), x = split_view(xprep, (long long)B * P, n_in);
    const cuuint64_t np = pad8(n_in);
    CUtensorMap mwh, mwl;
    if (int e = encode3(enc, &mwh, w.hi, n_in, n_out, P, np * 2, (cuuint64_t)n_out * np * 2, BKE, 128, 1)) return e;
    if (int e = encode3(enc, &mwl, w.lo, n_in, n_out, P, np * 2, (cuuint64_t)n_out * np * 2, BKE, 128, 1)) return e;
    Params p;
    p.B = B; p.P = P; p.n_out = n_out; p.n_in = n_in;
    p.out = out; p.o_sb = o_sb; p.accumulate = accumulate;
    p.m_tiles = ceil_div(n_out, 256);
    p.k_blocks = ceil_div(n_in, BKE);
    p.exp_w = w.exp; p.exp_x = x.exp; p.amax_out = amax_out; p.sub = sub; p.fast = fast;
    p.kbc = kb_per_chunk(n_in, kbc);
    p.t_hi = p.t_lo = nullptr; p.t_ld = 0; p.t_exp = nullptr; p.w_l1 = w.l1; p.sub_amax = sub_amax;
    {
        // measured (B200, cfg4, interleaved A/B): `+=` second stage 19.2 -> 18.4 ms per step; the `- sub` first stage gets
        // slower with it (15.6 -> 16.0 ms: its seeds are a quarter the size and the tensor pipe, not the seeds, paces it)
        static const int pf = [] { const char* e = getenv("DADMM_SEED_PREFETCH"); return e ? atoi(e) : 1; }();
        p.seed_prefetch = (pf == 2) ? 1 : (pf == 1 ? (accumulate ? 1 : 0) : 0);
    }
    if (tprep && sub && !sub_amax) DADMM_FAIL(-1, "contract_f16: a first stage with a subtracted term needs its max |.|");
    if (tprep) {
        const Split t = split_view(tprep, (long long)B * P, n_out);
        p.t_hi = t.hi; p.t_lo = t.lo; p.t_ld = pad8(n_out); p.t_exp = t.exp;
    }
    static int num_sms = [] {
        int dev = 0, n = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        return n;
    }();
    if (batch_tile(B, P, n_in, n_out, num_sms / 2) == 128)
        return launch_nt<128>(p, x, n_in, B, P, enc, mwh, mwl, num_sms, tprep != nullptr, s);
    return launch_nt<256>(p, x, n_in, B, P, enc, mwh, mwl, num_sms, tprep != nullptr, s);
}

}  // namespace f16
}  // namespace dadmm

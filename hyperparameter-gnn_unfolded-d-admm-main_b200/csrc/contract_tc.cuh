// contract_tc.cuh -- the per-agent contraction out[b,p,:] (+)= W_p x[b,p,:] on the 5th-generation tensor
// cores (tcgen05 / TMEM / TMA), fp32-accurate through a 3xTF32 split.
//
// The reference evaluates `torch.matmul(AtA[0,p], y[:,p])` per agent (unfolded_DLASSO.py:69-71); AtA is
// shared by the whole batch, so for each agent this is a dense GEMM  D[i,b] = sum_k W_p[i,k] * x[b,p,k]
// with M = n_out (rows of W_p), N = batch, K = n_in, both operands K-major in global memory.
//
// fp32 parity on tf32 tensor cores: every operand value v is used as  v = hi + lo  with hi = the tf32
// truncation the tensor core applies to the raw fp32 bits and lo = rn_tf32(v - hi); the product is
// accumulated as  A_lo*B_hi + A_hi*B_lo + A_hi*B_hi  in the fp32 TMEM accumulator (the dropped lo*lo
// term is <= 2^-20 relative).  The raw fp32 tiles land in shared memory by TMA and double as the "hi"
// operands; four converter warps derive the "lo" tiles element-wise at the same (swizzled) offsets, so
// they never have to know the swizzle.
//
// Accumulation: the tcgen05 accumulator TRUNCATES (round-toward-zero) on every accumulate -- measured on
// B200: a bias of -2^-24 per k-step, i.e. -7e-6 relative at K=1024 (profiles/r01_tc_accumulation_probe.txt),
// 10x the error of an fp32 FMA loop.  So the tensor core only ever sums K_CHUNK = 64 consecutive k (24 MMAs)
// from zero into a TMEM buffer; the epilogue warps drain each chunk and add it to fp32 REGISTER accumulators
// with round-to-nearest.  That restores FMA-loop accuracy and costs no tensor time (two TMEM buffers ping-pong).
//
// Kernel structure (persistent, one CTA per SM, 512 threads = 4 warpgroups, setmaxnreg-rebalanced):
//   warp 0       TMA producer      cp.async.bulk.tensor.3d  -> stage ring (4 x {A 128x16, B 256x16} fp32, SW64)
//   warp 1       MMA issuer        3 x tcgen05.mma.kind::tf32 (M128 N256 K8) per k-step, tcgen05.commit
//   warps 4-7    lo converters     generic-proxy reads/writes + fence.proxy.async
//   warps 8-15   accumulate+store  tcgen05.ld 32x32b.x32 per chunk -> registers (128 per thread); at the end of
//                                  a tile coalesced 128-byte row stores (optionally +=)
#pragma once
#include <cuda.h>
#include <cstdlib>

#include "common.cuh"

namespace dadmm {
namespace tc {

constexpr int BM = 128, BN = 256, BK = 16, STAGES = 4, ACC = 2;
constexpr int A_BYTES = BM * BK * 4;             // 8 KB
constexpr int B_BYTES = BN * BK * 4;             // 16 KB
constexpr int HALF_BYTES = A_BYTES + B_BYTES;    // raw (= hi) tiles of one stage
constexpr int STAGE_BYTES = 2 * HALF_BYTES;      // + lo tiles
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*alignment slack*/ + 256 /*barriers*/;
constexpr int NUM_THREADS = 512;
constexpr int CONV_WARP0 = 4, EPI_WARP0 = 8;
#ifndef DADMM_KB_PER_CHUNK
#define DADMM_KB_PER_CHUNK 4
#endif
constexpr int KB_PER_CHUNK = DADMM_KB_PER_CHUNK;  // 4 k-blocks of 16 = 64 k per tensor-core partial sum
constexpr int EPI_THREADS = 256;
constexpr int COLS_PER_THREAD = BN / 2;          // each accumulate warp owns one lane quarter x one column half
constexpr uint32_t TMEM_COLS = 512;

// tcgen05 instruction descriptor: D=f32, A=B=tf32, both K-major, N=256, M=128 (cute::UMMA::InstrDescriptor)
constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

// ---------------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// Spin with a watchdog: a protocol bug must surface as a trap, never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 27)) __trap();
    }
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tma_prefetch_3d(const CUtensorMap* map, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(map), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(IDESC), "r"(accumulate)
        : "memory");
}
// K-major, SWIZZLE_64B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): rows of 64 bytes,
// 8-row swizzle atoms 512 bytes apart.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(512 >> 4) << 32) | (1ull << 46) | (4ull << 61);
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

template <int N>
__device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N>
__device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }

// lo = rn_tf32(v - trunc_tf32(v)); the tensor core ignores the 13 low mantissa bits of a tf32 operand
__device__ __forceinline__ float tf32_lo(float v) {
    const float hi = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
    float lo = v - hi;
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(lo));
    return __uint_as_float(r);
}

struct Params {
    int B, P, n_out, n_in;
    float* out;
    long long o_sb;   // elements between consecutive problems in `out`
    int accumulate;
    int m_tiles, n_tiles, k_blocks, total_tiles;
};

__global__ void __launch_bounds__(NUM_THREADS, 1)
contract_tc_kernel(const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_x, const Params p) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bars = base + STAGES * STAGE_BYTES;
    // barrier slots (8 bytes each)
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto conv_bar = [&](int s) { return bars + 8u * (STAGES + s); };
    auto empty_bar = [&](int s) { return bars + 8u * (2 * STAGES + s); };
    auto tfull_bar = [&](int a) { return bars + 8u * (3 * STAGES + a); };
    auto tempty_bar = [&](int a) { return bars + 8u * (3 * STAGES + ACC + a); };
    const uint32_t tmem_slot = bars + 8u * (3 * STAGES + 2 * ACC);
    auto stage_base = [&](int s) { return base + (uint32_t)s * STAGE_BYTES; };

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(conv_bar(s), 128);
            mbar_init(empty_bar(s), 1);
        }
        for (int a = 0; a < ACC; ++a) {
            mbar_init(tfull_bar(a), 1);
            mbar_init(tempty_bar(a), EPI_THREADS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    const int tiles_per_agent = p.m_tiles * p.n_tiles;

    const int n_chunks = (p.k_blocks + KB_PER_CHUNK - 1) / KB_PER_CHUNK;

    if (warp < 4) {
        reg_dec<40>();
        if (warp == 0 && lane == 0) {
            // -------------------------------------------------------------- TMA producer
            asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
            int stage = 0;
            uint32_t phase = 0;
            for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
                const int ag = t / tiles_per_agent, r = t % tiles_per_agent;
                const int i0 = (r % p.m_tiles) * BM, b0 = (r / p.m_tiles) * BN;
                for (int kb = 0; kb < p.k_blocks; ++kb) {
                    mbar_wait(empty_bar(stage), phase ^ 1u);
                    mbar_expect_tx(full_bar(stage), HALF_BYTES);
                    tma_load_3d(stage_base(stage), &map_w, full_bar(stage), kb * BK, i0, ag);
                    tma_load_3d(stage_base(stage) + A_BYTES, &map_x, full_bar(stage), kb * BK, ag, b0);
                    if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                }
            }
        } else if (warp == 1 && lane == 0) {
            // -------------------------------------------------------------- MMA issuer
            int stage = 0;
            uint32_t phase = 0;
            int ci = 0;                                   // running chunk index of this CTA
            for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
                for (int ch = 0; ch < n_chunks; ++ch, ++ci) {
                    const int buf = ci & 1;
                    mbar_wait(tempty_bar(buf), ((uint32_t)(ci >> 1) & 1u) ^ 1u);
                    tcgen05_fence_after();
                    const uint32_t d_tmem = tmem_base + (uint32_t)buf * BN;
                    const int kb_end = min(p.k_blocks, (ch + 1) * KB_PER_CHUNK);
                    for (int kb = ch * KB_PER_CHUNK; kb < kb_end; ++kb) {
                        mbar_wait(conv_bar(stage), phase);
                        tcgen05_fence_after();
                        const uint32_t sa = stage_base(stage), sb = sa + A_BYTES;
                        const uint32_t sa_lo = sa + HALF_BYTES, sb_lo = sb + HALF_BYTES;
#pragma unroll
                        for (int ks = 0; ks < BK / 8; ++ks) {
                            const uint32_t koff = ks * 32;   // 8 tf32 = 32 bytes along K inside the 64-byte swizzled row
                            const uint64_t da = umma_desc(sa + koff), db = umma_desc(sb + koff);
                            const uint64_t da_lo = umma_desc(sa_lo + koff), db_lo = umma_desc(sb_lo + koff);
                            umma_tf32(d_tmem, da_lo, db, (kb != ch * KB_PER_CHUNK) || ks != 0);
                            umma_tf32(d_tmem, da, db_lo, 1u);
                            umma_tf32(d_tmem, da, db, 1u);
                        }
                        umma_commit(empty_bar(stage));
                        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                    }
                    umma_commit(tfull_bar(buf));
                }
            }
        }
    } else if (warp < EPI_WARP0) {
        // ------------------------------------------------------------------ lo converters (128 threads)
        reg_dec<56>();
        const int ct = threadIdx.x - CONV_WARP0 * 32;
        int stage = 0;
        uint32_t phase = 0;
        for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
            for (int kb = 0; kb < p.k_blocks; ++kb) {
                mbar_wait(full_bar(stage), phase);
                const uint32_t src = stage_base(stage), dst = src + HALF_BYTES;
#pragma unroll 4
                for (int c = ct; c < HALF_BYTES / 16; c += 128) {
                    float4 v;
                    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(src + 16u * c));
                    v.x = tf32_lo(v.x); v.y = tf32_lo(v.y); v.z = tf32_lo(v.z); v.w = tf32_lo(v.w);
                    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dst + 16u * c), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
                }
                fence_proxy_async();
                mbar_arrive(conv_bar(stage));
                if (++stage == STAGES) { stage = 0; phase ^= 1u; }
            }
        }
    } else {
        // ------------------------------------------------------------------ accumulate + store (256 threads)
        reg_inc<208>();
        const int q = warp & 3;                        // TMEM lane quarter this warp may access
        const int h = (warp - EPI_WARP0) >> 2;         // column half
        int ci = 0;
        for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
            const int ag = t / tiles_per_agent, r = t % tiles_per_agent;
            const int i0 = (r % p.m_tiles) * BM, b0 = (r / p.m_tiles) * BN + h * COLS_PER_THREAD;
            float acc[COLS_PER_THREAD];
            const int i = i0 + q * 32 + lane;
            float* orow = p.out + (long long)ag * p.n_out + i;
            if (p.accumulate) {
                // out += W x: seed the register accumulators with the old values; the loads fly while the
                // tensor core works on the first chunk
#pragma unroll
                for (int c = 0; c < COLS_PER_THREAD; ++c) {
                    const int b = b0 + c;
                    acc[c] = (i < p.n_out && b < p.B) ? __ldcs(orow + (long long)b * p.o_sb) : 0.0f;
                }
            }
            for (int ch = 0; ch < n_chunks; ++ch, ++ci) {
                const int buf = ci & 1;
                mbar_wait(tfull_bar(buf), (uint32_t)(ci >> 1) & 1u);
                tcgen05_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN + h * COLS_PER_THREAD);
                const bool seed = (ch == 0) && !p.accumulate;
#pragma unroll
                for (int j = 0; j < COLS_PER_THREAD / 32; ++j) {
                    uint32_t v[32];
                    tmem_ld32(taddr + 32u * j, v);
                    tmem_ld_wait();
                    if (seed) {
#pragma unroll
                        for (int c = 0; c < 32; ++c) acc[32 * j + c] = __uint_as_float(v[c]);
                    } else {
#pragma unroll
                        for (int c = 0; c < 32; ++c) acc[32 * j + c] += __uint_as_float(v[c]);
                    }
                }
                tcgen05_fence_before();
                mbar_arrive(tempty_bar(buf));
            }
            if (i < p.n_out) {
#pragma unroll
                for (int c = 0; c < COLS_PER_THREAD; ++c) {
                    const int b = b0 + c;
                    if (b < p.B) orow[(long long)b * p.o_sb] = acc[c];
                }
            }
        }
    }

    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// ===================================================================================================
// CTA-pair variant (cta_group::2): one 256 x 256 output tile per 2-CTA cluster.
//
// ncu on the single-CTA kernel (profiles/r01_ncu_summary_v1.txt): per k-block 144 KB cross the shared-memory
// port (72 KB UMMA operand reads + 24 KB TMA fill + 48 KB hi->lo conversion) in ~1200 cycles -- the port, not
// the tensor pipe (768 cycles), is the limiter.  With cta_group::2 each CTA keeps only ITS 128 rows of W_p (A) and
// ITS 128 of the 256 batch rows (half of B; the pair's tensor cores exchange the halves), so fill, conversion
// and operand reads per CTA all drop by a third: 96 KB per k-block.
//
//   CTA r of the pair:  A = W_p[i0+128r .. +128, k-block]   B-half = x[b0+128r .. +128, p, k-block]
//                        D rows i0+128r .. +128 (all 256 columns) live in CTA r's TMEM
//   barriers: full[s]   local   TMA -> converters
//             conv[s]   LEADER  4 local + 4 remote warp arrivals -> MMA issuer (leader CTA only)
//             empty[s]  local   <- tcgen05.commit multicast to both CTAs
//             tfull[b]  local   <- tcgen05.commit multicast
//             tempty[b] LEADER  8 local + 8 remote warp arrivals
// ===================================================================================================
constexpr int P2_STAGES = 6;
constexpr int P2_HALF = 2 * A_BYTES;                 // A tile + B half tile (raw = hi), 16 KB
constexpr int P2_STAGE = 2 * P2_HALF;                // + lo tiles, 32 KB
constexpr int P2_SMEM = P2_STAGES * P2_STAGE + 1024 + 256;
constexpr uint32_t IDESC2 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the barrier at the same shared-memory offset in CTA `rank` of the cluster.  Default (CTA-scope)
// semantics on purpose, as in CUTLASS' ClusterBarrier: an explicit .release.cluster / .acquire.cluster pair makes
// ptxas emit ERRBAR + CCTL.IVALL (L1 invalidate) in every wait loop -- 25 % of all stall samples in the first
// version of this kernel (ncu, round 1).  Operand visibility to the tensor core is carried by
// fence.proxy.async + the mbarrier itself.
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar, uint32_t rank) {
    asm volatile(
        "{\n\t.reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}"
        ::"r"(bar), "r"(rank)
        : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"((uint16_t)3)
                 : "memory");
}
__device__ __forceinline__ void umma_tf32_pair(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(IDESC2), "r"(accumulate)
        : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
contract_tc2_kernel(const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_x, const Params p) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bars = base + P2_STAGES * P2_STAGE;
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto conv_bar = [&](int s) { return bars + 8u * (P2_STAGES + s); };
    auto empty_bar = [&](int s) { return bars + 8u * (2 * P2_STAGES + s); };
    auto tfull_bar = [&](int a) { return bars + 8u * (3 * P2_STAGES + a); };
    auto tempty_bar = [&](int a) { return bars + 8u * (3 * P2_STAGES + ACC + a); };
    const uint32_t tmem_slot = bars + 8u * (3 * P2_STAGES + 2 * ACC);
    auto stage_base = [&](int s) { return base + (uint32_t)s * P2_STAGE; };

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;

    if (threadIdx.x == 0) {
        for (int s = 0; s < P2_STAGES; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(conv_bar(s), 8);       // 4 converter warps per CTA x 2 CTAs (used in the leader only)
            mbar_init(empty_bar(s), 1);
        }
        for (int a = 0; a < ACC; ++a) {
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
    cluster_sync_all();                       // both CTAs' barriers + TMEM are ready before any cross-CTA signal
    tcgen05_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    const int tiles_per_agent = p.m_tiles * p.n_tiles;     // pair tiles (256 x 256)
    const int n_chunks = (p.k_blocks + KB_PER_CHUNK - 1) / KB_PER_CHUNK;

    if (warp < 4) {
        reg_dec<40>();
        if (warp == 0 && lane == 0) {
            // -------------------------------------------------------------- TMA producer (each CTA: its own halves)
            asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
            int stage = 0;
            uint32_t phase = 0;
            for (int t = cluster_id; t < p.total_tiles; t += num_clusters) {
                const int ag = t / tiles_per_agent, r = t % tiles_per_agent;
                const int i0 = (r % p.m_tiles) * 256 + (int)rank * BM, b0 = (r / p.m_tiles) * 256 + (int)rank * 128;
                for (int kb = 0; kb < p.k_blocks; ++kb) {
#ifdef DADMM_TC_L2_PREFETCH
                    if (kb + DADMM_TC_L2_PREFETCH < p.k_blocks) {
                        tma_prefetch_3d(&map_w, (kb + DADMM_TC_L2_PREFETCH) * BK, i0, ag);
                        tma_prefetch_3d(&map_x, (kb + DADMM_TC_L2_PREFETCH) * BK, ag, b0);
                    }
#endif
                    mbar_wait(empty_bar(stage), phase ^ 1u);
#if defined(DADMM_TC_EXPERIMENT) && DADMM_TC_EXPERIMENT >= 2
                    mbar_arrive(full_bar(stage));      // timing experiment: no TMA traffic
#else
                    mbar_expect_tx(full_bar(stage), P2_HALF);
                    tma_load_3d(stage_base(stage), &map_w, full_bar(stage), kb * BK, i0, ag);
                    tma_load_3d(stage_base(stage) + A_BYTES, &map_x, full_bar(stage), kb * BK, ag, b0);
#endif
                    if (++stage == P2_STAGES) { stage = 0; phase ^= 1u; }
                }
            }
        } else if (warp == 1 && lane == 0 && rank == 0) {
            // -------------------------------------------------------------- MMA issuer (leader CTA)
            int stage = 0;
            uint32_t phase = 0;
            int ci = 0;
            for (int t = cluster_id; t < p.total_tiles; t += num_clusters) {
                for (int ch = 0; ch < n_chunks; ++ch, ++ci) {
                    const int buf = ci & 1;
                    mbar_wait(tempty_bar(buf), ((uint32_t)(ci >> 1) & 1u) ^ 1u);
                    tcgen05_fence_after();
                    const uint32_t d_tmem = tmem_base + (uint32_t)buf * 256;
                    const int kb_end = min(p.k_blocks, (ch + 1) * KB_PER_CHUNK);
                    for (int kb = ch * KB_PER_CHUNK; kb < kb_end; ++kb) {
                        mbar_wait(conv_bar(stage), phase);
                        tcgen05_fence_after();
                        const uint32_t sa = stage_base(stage), sb = sa + A_BYTES;
                        const uint32_t sa_lo = sa + P2_HALF, sb_lo = sb + P2_HALF;
#pragma unroll
                        for (int ks = 0; ks < BK / 8; ++ks) {
                            const uint32_t koff = ks * 32;
                            const uint64_t da = umma_desc(sa + koff), db = umma_desc(sb + koff);
                            const uint64_t da_lo = umma_desc(sa_lo + koff), db_lo = umma_desc(sb_lo + koff);
                            umma_tf32_pair(d_tmem, da_lo, db, (kb != ch * KB_PER_CHUNK) || ks != 0);
                            umma_tf32_pair(d_tmem, da, db_lo, 1u);
                            umma_tf32_pair(d_tmem, da, db, 1u);
                        }
                        umma_commit_pair(empty_bar(stage));
                        if (++stage == P2_STAGES) { stage = 0; phase ^= 1u; }
                    }
                    umma_commit_pair(tfull_bar(buf));
                }
            }
        }
    } else if (warp < EPI_WARP0) {
        // ------------------------------------------------------------------ lo converters (128 threads per CTA)
        reg_dec<56>();
        const int ct = threadIdx.x - CONV_WARP0 * 32;
        int stage = 0;
        uint32_t phase = 0;
        for (int t = cluster_id; t < p.total_tiles; t += num_clusters) {
            for (int kb = 0; kb < p.k_blocks; ++kb) {
                mbar_wait(full_bar(stage), phase);
                const uint32_t src = stage_base(stage), dst = src + P2_HALF;
#if defined(DADMM_TC_EXPERIMENT) && DADMM_TC_EXPERIMENT >= 1
                if (p.n_in < 0)      // timing experiment: skip the conversion work
#endif
#pragma unroll 4
                for (int c = ct; c < P2_HALF / 16; c += 128) {
                    float4 v;
                    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(src + 16u * c));
                    v.x = tf32_lo(v.x); v.y = tf32_lo(v.y); v.z = tf32_lo(v.z); v.w = tf32_lo(v.w);
                    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dst + 16u * c), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(conv_bar(stage), 0);
                if (++stage == P2_STAGES) { stage = 0; phase ^= 1u; }
            }
        }
    } else {
        // ------------------------------------------------------------------ accumulate + store (256 threads per CTA)
        reg_inc<208>();
        const int q = warp & 3;
        const int h = (warp - EPI_WARP0) >> 2;
        int ci = 0;
        for (int t = cluster_id; t < p.total_tiles; t += num_clusters) {
            const int ag = t / tiles_per_agent, r = t % tiles_per_agent;
            const int i0 = (r % p.m_tiles) * 256 + (int)rank * BM, b0 = (r / p.m_tiles) * 256 + h * COLS_PER_THREAD;
            float acc[COLS_PER_THREAD];
            const int i = i0 + q * 32 + lane;
            float* orow = p.out + (long long)ag * p.n_out + i;
            if (p.accumulate) {
#pragma unroll
                for (int c = 0; c < COLS_PER_THREAD; ++c) {
                    const int b = b0 + c;
                    acc[c] = (i < p.n_out && b < p.B) ? __ldcs(orow + (long long)b * p.o_sb) : 0.0f;
                }
            }
            for (int ch = 0; ch < n_chunks; ++ch, ++ci) {
                const int buf = ci & 1;
                mbar_wait(tfull_bar(buf), (uint32_t)(ci >> 1) & 1u);
                tcgen05_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * 256 + h * COLS_PER_THREAD);
                const bool seed = (ch == 0) && !p.accumulate;
#if defined(DADMM_TC_EXPERIMENT) && DADMM_TC_EXPERIMENT >= 3
                if (p.n_in < 0 || ch == 0)   // timing experiment: drain only the first chunk
#endif
#pragma unroll
                for (int j = 0; j < COLS_PER_THREAD / 32; ++j) {
                    uint32_t v[32];
                    tmem_ld32(taddr + 32u * j, v);
                    tmem_ld_wait();
                    if (seed) {
#pragma unroll
                        for (int c = 0; c < 32; ++c) acc[32 * j + c] = __uint_as_float(v[c]);
                    } else {
#pragma unroll
                        for (int c = 0; c < 32; ++c) acc[32 * j + c] += __uint_as_float(v[c]);
                    }
                }
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(tempty_bar(buf), 0);
            }
            if (i < p.n_out) {
#pragma unroll
                for (int c = 0; c < COLS_PER_THREAD; ++c) {
                    const int b = b0 + c;
                    if (b < p.B) orow[(long long)b * p.o_sb] = acc[c];
                }
            }
        }
    }

    tcgen05_fence_before();
    __syncthreads();
    cluster_sync_all();                       // the peer may still signal / read this CTA until here
    if (warp == 1) {
        __syncwarp();
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = [] {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            f = nullptr;
        return (EncodeTiledFn)f;
    }();
    return fn;
}

inline bool dims_supported(int B, int P, int n_out, int n_in) {
    return B >= 128 && n_out >= 64 && n_in >= 16 && (n_in % 4) == 0 && P >= 1;
}

inline bool shape_supported(int B, int P, int n_out, int n_in, const void* W, int64_t w_sp, int64_t w_si, int64_t w_sk,
                            const void* x, int64_t x_sb, int64_t x_sp, int64_t x_sk, const void* out, int64_t o_sb,
                            int64_t o_sp, int64_t o_si) {
    if (!dims_supported(B, P, n_out, n_in)) return false;
    const bool lay = w_sk == 1 && w_si == n_in && w_sp == (int64_t)n_out * n_in && x_sk == 1 && x_sp == n_in &&
                     (x_sb % 4) == 0 && x_sb >= (int64_t)P * n_in && o_si == 1 && o_sp == n_out;
    const bool al = (reinterpret_cast<uintptr_t>(W) % 16 == 0) && (reinterpret_cast<uintptr_t>(x) % 16 == 0) &&
                    (reinterpret_cast<uintptr_t>(out) % 4 == 0);
    return lay && al && encode_fn() != nullptr;
}

inline size_t workspace_bytes(int, int, int, int) { return 0; }

inline int launch(int B, int P, int n_out, int n_in, const float* W, const float* x, float* out, int64_t x_sb, int64_t o_sb,
                  int accumulate, void*, size_t, cudaStream_t s) {
    EncodeTiledFn enc = encode_fn();
    if (!enc) DADMM_FAIL(-4, "cuTensorMapEncodeTiled unavailable");
    const char* ev = getenv("DADMM_TC_VARIANT");
    const bool pair_box = (!ev || atoi(ev) == 2) && n_out > BM;
    CUtensorMap mw, mx;
    {
        cuuint64_t dims[3] = {(cuuint64_t)n_in, (cuuint64_t)n_out, (cuuint64_t)P};
        cuuint64_t strides[2] = {(cuuint64_t)n_in * 4, (cuuint64_t)n_out * n_in * 4};
        cuuint32_t box[3] = {BK, BM, 1}, es[3] = {1, 1, 1};
        CUresult r = enc(&mw, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)W, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) DADMM_FAIL(-4, "cuTensorMapEncodeTiled(W) failed: %d", (int)r);
    }
    {
        cuuint64_t dims[3] = {(cuuint64_t)n_in, (cuuint64_t)P, (cuuint64_t)B};
        cuuint64_t strides[2] = {(cuuint64_t)n_in * 4, (cuuint64_t)x_sb * 4};
        cuuint32_t box[3] = {BK, 1, (cuuint32_t)(pair_box ? 128 : BN)}, es[3] = {1, 1, 1};
        CUresult r = enc(&mx, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)x, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) DADMM_FAIL(-4, "cuTensorMapEncodeTiled(x) failed: %d", (int)r);
    }
    Params p;
    p.B = B; p.P = P; p.n_out = n_out; p.n_in = n_in;
    p.out = out; p.o_sb = o_sb; p.accumulate = accumulate;
    static const int variant = [] {                  // DADMM_TC_VARIANT=1 forces the single-CTA kernel
        const char* e = getenv("DADMM_TC_VARIANT");
        return e ? atoi(e) : 2;
    }();
    const bool pair = (variant == 2) && n_out > BM;  // a 256-row pair tile needs more than one 128-row tile of work
    p.m_tiles = ceil_div(n_out, pair ? 256 : BM);
    p.n_tiles = ceil_div(B, BN);
    p.k_blocks = ceil_div(n_in, BK);
    p.total_tiles = P * p.m_tiles * p.n_tiles;
    static int num_sms = [] {
        int dev = 0, n = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        return n;
    }();
    static PerDeviceOnce attr_once;
    if (attr_once.first()) {
        DADMM_CUDA(cudaFuncSetAttribute(contract_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        DADMM_CUDA(cudaFuncSetAttribute(contract_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, P2_SMEM));
    }
    ProfScope prof(PROF_CONTRACT_TC, s);
    if (pair) {
        const int clusters = std::min(num_sms / 2, p.total_tiles);
        contract_tc2_kernel<<<2 * clusters, NUM_THREADS, P2_SMEM, s>>>(mw, mx, p);
    } else {
        const int grid = std::min(num_sms, p.total_tiles);
        contract_tc_kernel<<<grid, NUM_THREADS, SMEM_BYTES, s>>>(mw, mx, p);
    }
    DADMM_LAUNCHED();
    return 0;
}

}  // namespace tc
}  // namespace dadmm

// contract_simt.cuh -- batched, fully strided contraction on the FP32/FP64 FMA pipe:
//
//     out(b,p,i) (+)= sum_k W(p,i,k) * x(b,p,k)
//
// i.e. the reference's per-agent `torch.matmul(AtA[0,p], y[:,p])` loop (unfolded_DLASSO.py:69-71),
// `compute_Atx` (:120-124) and their autograd backward, for all agents in one launch.  Exact IEEE
// FMA accumulation, k ascending -- this is the bit-reproducible path and the fallback for shapes /
// dtypes the tcgen05 kernel (contract_tc.cuh) does not take.  Classic register-tiled SGEMM:
// 128x128x16 CTA tile (64x64x16 in fp64), 8x8 (4x4) outputs per thread, double-buffered shared
// memory with register prefetch.
#pragma once
#include "common.cuh"

namespace dadmm {

template <typename T>
struct GemmParams {
    int B, P, M, Kd;  // M = n_out, Kd = n_in
    const T* W;
    long long w_sp, w_si, w_sk;
    const T* X;
    long long x_sb, x_sp, x_sk;
    T* O;
    long long o_sb, o_sp, o_si;
    int accumulate;
};

// operand-tile load modes
enum { LD_VEC_K = 0, LD_GENERIC = 1, LD_ROW_CONTIG = 2 };

template <typename T, int ROWS, int BK, int MODE, int NT>
struct TileLoader {
    static constexpr int VECK = (MODE == LD_VEC_K) ? 4 : 1;
    static constexpr int PER_THREAD = ROWS * BK / (NT * VECK);
    static_assert(ROWS * BK % (NT * VECK) == 0, "tile must divide evenly");
    T reg[PER_THREAD][VECK];

    __device__ __forceinline__ void load(const T* base, long long row_stride, long long k_stride, int row0, int nrows,
                                         int k0, int K, int tid) {
#pragma unroll
        for (int l = 0; l < PER_THREAD; ++l) {
            const int idx = tid + l * NT;
            int row, k;
            if constexpr (MODE == LD_VEC_K) {
                row = idx / (BK / 4);
                k = (idx % (BK / 4)) * 4;
            } else if constexpr (MODE == LD_GENERIC) {
                row = idx / BK;
                k = idx % BK;
            } else {
                k = idx / ROWS;
                row = idx % ROWS;
            }
            const bool ok = (row0 + row < nrows) && (k0 + k < K);
            if constexpr (MODE == LD_VEC_K) {
                Vec<T, 4> v;
                if (ok) v = ld_vec<T, 4>(base + (long long)(row0 + row) * row_stride + (k0 + k));
                else v.v[0] = v.v[1] = v.v[2] = v.v[3] = (T)0;
#pragma unroll
                for (int j = 0; j < 4; ++j) reg[l][j] = v.v[j];
            } else {
                reg[l][0] = ok ? base[(long long)(row0 + row) * row_stride + (long long)(k0 + k) * k_stride] : (T)0;
            }
        }
    }
    template <int LDS>
    __device__ __forceinline__ void store(T (*Ts)[LDS], int tid) const {
#pragma unroll
        for (int l = 0; l < PER_THREAD; ++l) {
            const int idx = tid + l * NT;
            if constexpr (MODE == LD_VEC_K) {
                const int row = idx / (BK / 4), k = (idx % (BK / 4)) * 4;
#pragma unroll
                for (int j = 0; j < 4; ++j) Ts[k + j][row] = reg[l][j];
            } else if constexpr (MODE == LD_GENERIC) {
                Ts[idx % BK][idx / BK] = reg[l][0];
            } else {
                Ts[idx / ROWS][idx % ROWS] = reg[l][0];
            }
        }
    }
};

template <typename T, int BM, int BN, int BK, int TM, int TN, int WMODE, int XMODE>
__global__ void __launch_bounds__((BM / TM) * (BN / TN)) contract_simt_kernel(const GemmParams<T> p) {
    constexpr int NT = (BM / TM) * (BN / TN);
    constexpr int PAD = 4;
    constexpr int GM = TM / 4, GN = TN / 4;  // groups of 4 consecutive outputs per thread
    __shared__ __align__(32) T As[2][BK][BM + PAD];
    __shared__ __align__(32) T Bs[2][BK][BN + PAD];
    pdl_wait();            // chain kernel (common.cuh): nothing above touches global memory
    pdl_trigger();
    const int tid = threadIdx.x;
    const int tx = tid % (BM / TM), ty = tid / (BM / TM);
    const int i0 = blockIdx.x * BM, b0 = blockIdx.y * BN, ag = blockIdx.z;
    const T* Wp = p.W + (long long)ag * p.w_sp;
    const T* Xp = p.X + (long long)ag * p.x_sp;

    TileLoader<T, BM, BK, WMODE, NT> lw;
    TileLoader<T, BN, BK, XMODE, NT> lx;
    T acc[TM][TN];
#pragma unroll
    for (int a = 0; a < TM; ++a)
#pragma unroll
        for (int c = 0; c < TN; ++c) acc[a][c] = (T)0;

    const int nk = (p.Kd + BK - 1) / BK;
    lw.load(Wp, p.w_si, p.w_sk, i0, p.M, 0, p.Kd, tid);
    lx.load(Xp, p.x_sb, p.x_sk, b0, p.B, 0, p.Kd, tid);
    lw.store(As[0], tid);
    lx.store(Bs[0], tid);
    __syncthreads();
    for (int kb = 0; kb < nk; ++kb) {
        const int cur = kb & 1;
        if (kb + 1 < nk) {
            lw.load(Wp, p.w_si, p.w_sk, i0, p.M, (kb + 1) * BK, p.Kd, tid);
            lx.load(Xp, p.x_sb, p.x_sk, b0, p.B, (kb + 1) * BK, p.Kd, tid);
        }
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            T a[TM], b[TN];
#pragma unroll
            for (int g = 0; g < GM; ++g) {
                const Vec<T, 4> v = ld_vec<T, 4>(&As[cur][kk][g * (BM / GM) + tx * 4]);
#pragma unroll
                for (int j = 0; j < 4; ++j) a[g * 4 + j] = v.v[j];
            }
#pragma unroll
            for (int g = 0; g < GN; ++g) {
                const Vec<T, 4> v = ld_vec<T, 4>(&Bs[cur][kk][g * (BN / GN) + ty * 4]);
#pragma unroll
                for (int j = 0; j < 4; ++j) b[g * 4 + j] = v.v[j];
            }
#pragma unroll
            for (int x = 0; x < TM; ++x)
#pragma unroll
                for (int z = 0; z < TN; ++z) acc[x][z] = fma(a[x], b[z], acc[x][z]);
        }
        if (kb + 1 < nk) {
            lw.store(As[cur ^ 1], tid);
            lx.store(Bs[cur ^ 1], tid);
        }
        __syncthreads();
    }

    T* Op = p.O + (long long)ag * p.o_sp;
#pragma unroll
    for (int gn = 0; gn < GN; ++gn)
#pragma unroll
        for (int jn = 0; jn < 4; ++jn) {
            const int b = b0 + gn * (BN / GN) + ty * 4 + jn;
            if (b >= p.B) continue;
#pragma unroll
            for (int gm = 0; gm < GM; ++gm) {
                const int i = i0 + gm * (BM / GM) + tx * 4;
                T* q = Op + (long long)b * p.o_sb + (long long)i * p.o_si;
                const bool vec_ok = (p.o_si == 1) && (i + 3 < p.M) && ((reinterpret_cast<uintptr_t>(q) & (sizeof(T) * 4 - 1)) == 0);
                if (vec_ok) {
                    Vec<T, 4> v;
#pragma unroll
                    for (int j = 0; j < 4; ++j) v.v[j] = acc[gm * 4 + j][gn * 4 + jn];
                    if (p.accumulate) {
                        const Vec<T, 4> o = ld_vec<T, 4>(q);
#pragma unroll
                        for (int j = 0; j < 4; ++j) v.v[j] += o.v[j];
                    }
                    st_vec<T, 4>(q, v);
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        if (i + j < p.M) {
                            T* qq = q + (long long)j * p.o_si;
                            const T r = acc[gm * 4 + j][gn * 4 + jn];
                            *qq = p.accumulate ? (*qq + r) : r;
                        }
                    }
                }
            }
        }
}

template <typename T>
struct SimtCfg;
template <>
struct SimtCfg<float> {
    static constexpr int BM = 128, BN = 128, BK = 16, TM = 8, TN = 8;
};
template <>
struct SimtCfg<double> {
    static constexpr int BM = 64, BN = 64, BK = 16, TM = 4, TN = 4;
};
// small batches (the reference's default B = 16..64): a 128-wide batch tile would idle 75 % of the CTA and leave
// most SMs without work, so use 64 x 32 tiles / 128 threads
template <typename T>
struct SimtCfgSmall {
    static constexpr int BM = 64, BN = 32, BK = 16, TM = 4, TN = 4;
};

template <typename T, typename C, int WMODE, int XMODE>
int launch_contract_simt_cfg(const GemmParams<T>& p, cudaStream_t s) {
    dim3 grid(ceil_div(p.M, C::BM), ceil_div(p.B, C::BN), p.P);
    ProfScope prof(PROF_CONTRACT_SIMT, s);
    DADMM_CUDA(launch_chain(contract_simt_kernel<T, C::BM, C::BN, C::BK, C::TM, C::TN, WMODE, XMODE>, grid,
                            dim3((C::BM / C::TM) * (C::BN / C::TN)), 0, s, p));
    DADMM_LAUNCHED();
    return 0;
}

// ---- skinny batches (B <= 64: the reference's own runs use 16..64 problems per step) -------------------------------------
// At B = 32, P = 5, n = 500 the tiled kernel above has 40 CTAs walking a 32-step k loop with a barrier per step: 36 us per
// launch, 56 % of the GPU time of a configs[0] training step (ncu launch list, round 2).  The work is 80 MFLOP against a
// 5 MB operator that lives in L2 -- a few microseconds if every SM takes part.  Here a CTA owns 16 output rows x 32
// problems of one agent and its 8 warps SPLIT the contraction index between them (warp w takes the 16-byte groups
// w, w + 8, ...; the 8 warps together read 128 contiguous bytes of each row per step); a lane holds a 4 x 4 block of the
// outputs and feeds it from 16-byte loads straight from global memory / L1 (no shared-memory staging, no barrier in the
// loop); the eight partial sums are added in warp order through shared memory at the end: deterministic, and a fixed
// summation tree instead of the k-ascending chain of the tiled kernel (the reference's torch.matmul promises neither).
constexpr int kSkinnyRows = 16, kSkinnyBatch = 32, kSkinnyWarps = 8;

template <typename T>
__global__ void __launch_bounds__(kSkinnyWarps * 32) contract_skinny_kernel(const GemmParams<T> p) {
    __shared__ __align__(16) T red[kSkinnyWarps][kSkinnyRows * kSkinnyBatch];
    pdl_wait();
    pdl_trigger();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rg = lane & 3, bg = lane >> 2;                       // 4 row groups x 8 batch groups of 4
    const int i0 = blockIdx.x * kSkinnyRows, b0 = blockIdx.y * kSkinnyBatch, ag = blockIdx.z;
    const T* wrow[4];
    const T* xrow[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {                                   // out-of-range rows / problems read the last valid one
        const int i = min(i0 + rg * 4 + j, p.M - 1), b = min(b0 + bg * 4 + j, p.B - 1);
        wrow[j] = p.W + (long long)ag * p.w_sp + (long long)i * p.w_si;
        xrow[j] = p.X + (long long)ag * p.x_sp + (long long)b * p.x_sb;
    }
    T acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[a][c] = (T)0;
    const int nk4 = p.Kd >> 2;
#pragma unroll 4
    for (int k4 = warp; k4 < nk4; k4 += kSkinnyWarps) {
        Vec<T, 4> w[4], x[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) w[j] = ld_vec<T, 4>(wrow[j] + 4 * k4);
#pragma unroll
        for (int j = 0; j < 4; ++j) x[j] = ld_vec<T, 4>(xrow[j] + 4 * k4);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[a][c] = fma(w[a].v[kk], x[c].v[kk], acc[a][c]);
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {                                   // red[warp][b][i]: 4 consecutive rows per store
        Vec<T, 4> v;
#pragma unroll
        for (int a = 0; a < 4; ++a) v.v[a] = acc[a][c];
        st_vec<T, 4>(&red[warp][(bg * 4 + c) * kSkinnyRows + rg * 4], v);
    }
    __syncthreads();
    T* Op = p.O + (long long)ag * p.o_sp;
    for (int o = threadIdx.x; o < kSkinnyRows * kSkinnyBatch; o += kSkinnyWarps * 32) {
        const int i = i0 + o % kSkinnyRows, b = b0 + o / kSkinnyRows;
        if (i >= p.M || b >= p.B) continue;
        T sum = red[0][o];
#pragma unroll
        for (int w = 1; w < kSkinnyWarps; ++w) sum += red[w][o];
        T* q = Op + (long long)b * p.o_sb + (long long)i * p.o_si;
        *q = p.accumulate ? (*q + sum) : sum;
    }
}

template <typename T>
int launch_contract_skinny(const GemmParams<T>& p, cudaStream_t s) {
    dim3 grid(ceil_div(p.M, kSkinnyRows), ceil_div(p.B, kSkinnyBatch), p.P);
    ProfScope prof(PROF_CONTRACT_SIMT, s);
    DADMM_CUDA(launch_chain(contract_skinny_kernel<T>, grid, dim3(kSkinnyWarps * 32), 0, s, p));
    DADMM_LAUNCHED();
    return 0;
}

inline bool skinny_enabled() {
    static const bool on = [] {
        const char* e = getenv("DADMM_SKINNY");
        return !(e && atoi(e) == 0);
    }();
    return on;
}

template <typename T, int WMODE, int XMODE>
int launch_contract_simt_modes(const GemmParams<T>& p, cudaStream_t s) {
    if constexpr (WMODE == LD_VEC_K && XMODE == LD_VEC_K)
        if (p.B <= 64 && skinny_enabled()) return launch_contract_skinny<T>(p, s);
    if (p.B <= 64) return launch_contract_simt_cfg<T, SimtCfgSmall<T>, WMODE, XMODE>(p, s);
    return launch_contract_simt_cfg<T, SimtCfg<T>, WMODE, XMODE>(p, s);
}

template <typename T>
int operand_mode(const T* base, long long row_stride, long long k_stride, long long batch_stride, int K) {
    if (k_stride == 1) {
        const bool aligned = (reinterpret_cast<uintptr_t>(base) % (4 * sizeof(T)) == 0) && (row_stride % 4 == 0) &&
                             (batch_stride % 4 == 0) && (K % 4 == 0);
        return aligned ? LD_VEC_K : LD_GENERIC;
    }
    if (row_stride == 1) return LD_ROW_CONTIG;
    return LD_GENERIC;
}

template <typename T>
int launch_contract_simt(const GemmParams<T>& p, cudaStream_t s) {
    const int wm = operand_mode(p.W, p.w_si, p.w_sk, p.w_sp, p.Kd);
    const int xm = operand_mode(p.X, p.x_sb, p.x_sk, p.x_sp, p.Kd);
#define DADMM_SIMT_CASE(WM, XM) \
    if (wm == WM && xm == XM) return launch_contract_simt_modes<T, WM, XM>(p, s);
    DADMM_SIMT_CASE(LD_VEC_K, LD_VEC_K)
    DADMM_SIMT_CASE(LD_VEC_K, LD_GENERIC)
    DADMM_SIMT_CASE(LD_VEC_K, LD_ROW_CONTIG)
    DADMM_SIMT_CASE(LD_GENERIC, LD_VEC_K)
    DADMM_SIMT_CASE(LD_GENERIC, LD_GENERIC)
    DADMM_SIMT_CASE(LD_GENERIC, LD_ROW_CONTIG)
    DADMM_SIMT_CASE(LD_ROW_CONTIG, LD_VEC_K)
    DADMM_SIMT_CASE(LD_ROW_CONTIG, LD_GENERIC)
    DADMM_SIMT_CASE(LD_ROW_CONTIG, LD_ROW_CONTIG)
#undef DADMM_SIMT_CASE
    DADMM_FAIL(-1, "contract_simt: no kernel for operand modes %d/%d", wm, xm);
}

}  // namespace dadmm

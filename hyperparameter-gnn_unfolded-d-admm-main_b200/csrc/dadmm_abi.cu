// dadmm_abi.cu -- extern "C" entry points of libdadmm_sm100.so (see include/dadmm.h).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared -Xcompiler -fPIC
#include "../../include/dadmm.h"

#include <cmath>
#include <cstring>
#include <limits>
#include <mutex>
#include <vector>

#include "common.cuh"
#include "contract_simt.cuh"
#include "contract_tc.cuh"
#include "contract_f16.cuh"
#include "step.cuh"
#include "unfolded.cuh"
#include "unfolded_lean.cuh"
#include "unfolded_pipe.cuh"
#include "gcn.cuh"

namespace dadmm {
thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};
std::atomic<int> g_pdl{-1};
std::atomic<int> g_exact_order{-1};      // consensus accumulation order of the lean forward level (dadmm_set_consensus_order)

// ------------------------------------------------------------------------------------------
// per-kind kernel timing
// ------------------------------------------------------------------------------------------
struct ProfRec { int kind; cudaEvent_t a, b; };
static std::mutex g_prof_mu;
static bool g_prof_on = false;
static std::vector<ProfRec> g_prof;
static thread_local int g_prof_open = -1;

void prof_begin(int kind, cudaStream_t s) {
    if (!g_prof_on) return;
    ProfRec r{kind, nullptr, nullptr};
    if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) return;
    cudaEventRecord(r.a, s);
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof.push_back(r);
    g_prof_open = (int)g_prof.size() - 1;
}
void prof_end(cudaStream_t s) {
    if (!g_prof_on || g_prof_open < 0) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    if (g_prof_open < (int)g_prof.size()) cudaEventRecord(g_prof[g_prof_open].b, s);
    g_prof_open = -1;
}

// ------------------------------------------------------------------------------------------
// tile configuration of the step kernels
// ------------------------------------------------------------------------------------------
struct StepCfg {
    int vec, TB, nchunks, grid;
    size_t smem_fwd, smem_bwd;
};

static bool aligned_to(const void* p, size_t a) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) % a) == 0; }

static int forced_tb() {
    static const int forced = [] { const char* e = getenv("DADMM_STEP_TB"); return e ? atoi(e) : 0; }();   // experiment knob
    return forced;
}

// narr_fwd: shared-memory tile arrays the forward needs (0..2); the backward always needs 2 (+R scalars)
static int step_cfg(int dtype, int B, int P, int n, int narr_fwd, int max_vec, StepCfg* c, bool fwd_only = false) {
    const size_t es = dtype == DADMM_F64 ? 8 : 4;
    const size_t budget = 100 * 1024;  // two CTAs per SM
    int TB = std::max(1, std::min(B, (16 + P - 1) / P));
    // 16..31 agents: two problems per tile (BASELINE configs[2], P = 20: a 20-row tile paid the per-CTA staging and barriers
    // for too little work -- forward levels 2.71 -> 2.45 ms, backward 4.12 -> 3.87 ms per step; four problems: 2.75 / 5.33).
    // Even batches only: the lean level kernels take whole tiles.
    if (P >= 16 && P < 32 && B % 2 == 0) TB = 2;
    if (forced_tb() > 0) TB = std::max(1, std::min(B, forced_tb()));
    for (int vec = max_vec; vec >= 1; vec >>= 1) {
        if (n % vec) continue;
        for (int tb = TB; tb >= 1; --tb) {
            const size_t R = (size_t)tb * P, CH = 32 * vec;
            const size_t sf = (size_t)narr_fwd * R * CH * es;
            const size_t sb = fwd_only ? 0 : 2 * R * CH * es + R * es;
            if (std::max(sf, sb) <= budget || (vec == 1 && tb == 1 && std::max(sf, sb) <= 227 * 1024)) {
                c->nchunks = (n + (int)CH - 1) / (int)CH;
                // skinny batches (the reference's own B = 16..64): fewer problems per tile until there is a CTA per SM
                // (configs[0], B = 32, P = 5: TB = 4 left 16 / 32 CTAs, 20 / 17 us per forward / backward level)
                while (tb > 1 && !forced_tb() && (long long)((B + tb - 1) / tb) * c->nchunks < 148) --tb;
                c->vec = vec;
                c->TB = tb;
                c->grid = c->nchunks * ((B + tb - 1) / tb);
                c->smem_fwd = sf;
                c->smem_bwd = sb;
                return 0;
            }
        }
    }
    DADMM_FAIL(-2, "step kernels: P=%d does not fit in shared memory", P);
}

// Vector width of the step kernels: the widest of {4,2,1} (fp32) / {2,1} (fp64) that divides n.  It
// depends on n only, so that dadmm_reduce_hyp can recompute the chunking; every tensor pointer must
// then be aligned to vec*sizeof(T) (torch allocations and [k] slices of [K,B,P,n] tensors are).
static int max_vec_for(int dtype, int n) {
    int vec = dtype == DADMM_F64 ? 2 : 4;
    while (vec > 1 && (n % vec)) vec >>= 1;
    return vec;
}
static int check_aligned(int dtype, int vec, std::initializer_list<const void*> ptrs) {
    const size_t es = dtype == DADMM_F64 ? 8 : 4;
    for (const void* p : ptrs)
        if (!aligned_to(p, es * vec)) DADMM_FAIL(-5, "step kernels: tensor pointer %p is not %zu-byte aligned", p, es * vec);
    return 0;
}

template <typename K>
static int allow_smem(K kernel, size_t bytes) {
    if (bytes > 48 * 1024) DADMM_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return 0;
}

static int check_graph(const dadmm_graph* g, int P) {
    if (!g || !g->ev_ptr || !g->ev_idx || !g->deg || !g->adj_ptr || !g->adj_idx) DADMM_FAIL(-3, "graph: null CSR pointers");
    if (g->P != P) DADMM_FAIL(-3, "graph: P mismatch (%d vs %d)", g->P, P);
    if (g->n_graphs < 1) DADMM_FAIL(-3, "graph: n_graphs < 1");
    return 0;
}

template <typename T, int VEC>
static int launch_step_fwd_t(const StepFwdParams<T>& p, const StepCfg& c, cudaStream_t s) {
    if (int e = allow_smem(step_fwd_kernel<T, VEC>, c.smem_fwd)) return e;
    ProfScope prof(PROF_STEP_FWD, s);
    step_fwd_kernel<T, VEC><<<c.grid, kStepThreads, c.smem_fwd, s>>>(p);
    DADMM_LAUNCHED();
    return 0;
}
template <typename T, int VEC>
static int launch_step_bwd_t(const StepBwdParams<T>& p, const StepCfg& c, cudaStream_t s) {
    if (int e = allow_smem(step_bwd_kernel<T, VEC>, c.smem_bwd)) return e;
    ProfScope prof(PROF_STEP_BWD, s);
    step_bwd_kernel<T, VEC><<<c.grid, kStepThreads, c.smem_bwd, s>>>(p);
    DADMM_LAUNCHED();
    return 0;
}

template <typename T>
static int step_fwd_impl(int dtype, int B, int P, int n, const dadmm_graph* g, const dadmm_clamps* cl,
                         const dadmm_hyp* h, const void* y, const void* U, const void* delta, const void* a,
                         const void* atb, void* y_next, void* U_next, void* delta_next, void* graw, int32_t* flags,
                         cudaStream_t s) {
    StepFwdParams<T> p;
    p.B = B; p.P = P; p.n = n;
    p.ev_ptr = g->ev_ptr; p.ev_idx = g->ev_idx; p.deg = g->deg; p.gid = g->graph_id;
    p.hyp = (const T*)h->ptr; p.hsb = h->stride_b; p.hsp = h->stride_p; p.hsc = h->stride_c;
    p.G = (T)cl->G; p.V = (T)cl->V; p.Uc = (T)cl->Uc;
    p.hasD = std::isfinite(cl->D) ? 1 : 0;
    p.D = p.hasD ? (T)cl->D : std::numeric_limits<T>::infinity();
    p.y = (const T*)y; p.U = (const T*)U; p.delta = (const T*)delta; p.a = (const T*)a; p.atb = (const T*)atb;
    p.y_next = (T*)y_next; p.U_next = (T*)U_next; p.delta_next = (T*)delta_next; p.graw = (T*)graw;
    p.flags = flags;
    const int narr = (delta ? 0 : 1) + ((U_next || delta_next) ? 1 : 0);
    StepCfg c;
    if (int e = step_cfg(dtype, B, P, n, narr, max_vec_for(dtype, n), &c)) return e;
    if (int e = check_aligned(dtype, c.vec, {y, U, delta, a, atb, y_next, U_next, delta_next, graw})) return e;
    p.TB = c.TB;
    if constexpr (sizeof(T) == 4) {
        if (c.vec == 4) return launch_step_fwd_t<T, 4>(p, c, s);
    }
    if (c.vec == 2) return launch_step_fwd_t<T, 2>(p, c, s);
    return launch_step_fwd_t<T, 1>(p, c, s);
}

template <typename T>
static int step_bwd_impl(int dtype, int B, int P, int n, const dadmm_graph* g, const dadmm_clamps* cl,
                         const dadmm_hyp* h, const void* y, const void* U, const void* delta, const void* graw,
                         const void* y_next, const void* gy_a, const void* gy_b, const void* gU_next,
                         const void* gd_next, const void* label, double loss_coef, void* gy, void* ga, void* gU,
                         void* gd, void* partials, cudaStream_t s) {
    StepBwdParams<T> p;
    p.B = B; p.P = P; p.n = n;
    p.ev_ptr = g->ev_ptr; p.ev_idx = g->ev_idx; p.deg = g->deg; p.gid = g->graph_id;
    p.hyp = (const T*)h->ptr; p.hsb = h->stride_b; p.hsp = h->stride_p; p.hsc = h->stride_c;
    p.G = (T)cl->G; p.V = (T)cl->V; p.Uc = (T)cl->Uc;
    p.hasD = std::isfinite(cl->D) ? 1 : 0;
    p.D = p.hasD ? (T)cl->D : std::numeric_limits<T>::infinity();
    p.y = (const T*)y; p.U = (const T*)U; p.delta = (const T*)delta; p.graw = (const T*)graw;
    p.y_next = (const T*)y_next;
    p.gy_a = (const T*)gy_a; p.gy_b = (const T*)gy_b; p.gU_next = (const T*)gU_next; p.gd_next = (const T*)gd_next;
    p.label = (loss_coef != 0.0) ? (const T*)label : nullptr;
    p.loss_coef = (T)loss_coef;
    p.gy = (T*)gy; p.ga = (T*)ga; p.gU = (T*)gU; p.gd = (T*)gd; p.partials = (T*)partials;
    StepCfg c;
    if (int e = step_cfg(dtype, B, P, n, 2, max_vec_for(dtype, n), &c)) return e;
    if (int e = check_aligned(dtype, c.vec, {y, U, delta, graw, y_next, gy_a, gy_b, gU_next, gd_next, label, gy, ga, gU, gd}))
        return e;
    p.TB = c.TB;
    if constexpr (sizeof(T) == 4) {
        if (c.vec == 4) return launch_step_bwd_t<T, 4>(p, c, s);
    }
    if (c.vec == 2) return launch_step_bwd_t<T, 2>(p, c, s);
    return launch_step_bwd_t<T, 1>(p, c, s);
}

// The partial-sum buffer is laid out for the coarsest chunking the backward may pick (vec=1).
static size_t partials_elems(int B, int P, int n) { return (size_t)((n + 31) / 32) * B * P * 4; }

template <typename T>
static int reduce_hyp_impl(int dtype, int B, int P, int n, const void* partials, int per_sample, void* ghyp,
                           int64_t sb, int64_t sp, int64_t sc, int accumulate, int nchunks, cudaStream_t s) {
    ProfScope prof(PROF_REDUCE_HYP, s);
    if (per_sample) {
        const long long tot = (long long)B * P;
        reduce_hyp_sample_kernel<T><<<(unsigned)ceil_div64(tot, 256), 256, 0, s>>>((const T*)partials, nchunks, B, P,
                                                                                   (T*)ghyp, sb, sp, sc, accumulate);
    } else {
        reduce_hyp_table_kernel<T><<<P, 256, 0, s>>>((const T*)partials, nchunks, B, P, (T*)ghyp, sp, sc, accumulate);
    }
    DADMM_LAUNCHED();
    return 0;
}

// chunk count the backward kernel uses for (dtype, B, P, n) -- same derivation as step_bwd_impl
static int bwd_nchunks(int dtype, int B, int P, int n, int* out) {
    StepCfg c;
    if (int e = step_cfg(dtype, B, P, n, 2, max_vec_for(dtype, n), &c)) return e;
    *out = c.nchunks;
    return 0;
}

// Which kernel serves a contraction call (AUTO: f16 pairs > tf32 > FMA, by shape / layout support).
static int resolve_algo(int dtype, int algo, int B, int P, int n_out, int n_in, bool tc_layout) {
    if (dtype != DADMM_F32 || algo == DADMM_ALGO_SIMT) return DADMM_ALGO_SIMT;
    const bool f16_ok = tc_layout && f16::dims_supported(B, P, n_out, n_in);
    const bool tf32_ok = tc_layout && tc::dims_supported(B, P, n_out, n_in);
    if (algo == DADMM_ALGO_TC_3XF16 || algo == DADMM_ALGO_TC_F16X1) return f16_ok ? algo : -1;
    if (algo == DADMM_ALGO_TC_3XTF32) return tf32_ok ? DADMM_ALGO_TC_3XTF32 : -1;
    static const int prefer = [] {                       // DADMM_TC_PREFER=tf32 keeps AUTO on the 3xTF32 kernels
        const char* e = getenv("DADMM_TC_PREFER");
        return (e && !strcmp(e, "tf32")) ? DADMM_ALGO_TC_3XTF32 : DADMM_ALGO_TC_3XF16;
    }();
    if (f16_ok && prefer == DADMM_ALGO_TC_3XF16) return DADMM_ALGO_TC_3XF16;
    if (tf32_ok) return DADMM_ALGO_TC_3XTF32;
    return DADMM_ALGO_SIMT;
}

static size_t contract_ws_bytes(int dtype, int algo, int B, int P, int n_out, int n_in) {
    const int ra = resolve_algo(dtype, algo, B, P, n_out, n_in, true);
    if (ra != DADMM_ALGO_TC_3XF16 && ra != DADMM_ALGO_TC_F16X1) return 0;
    return f16::split_bytes((long long)P * n_out, n_in) + f16::split_bytes((long long)B * P, n_in);
}

// w_prepared: the fp16 split of W already sits at the head of `ws` (K-loop drivers split the operator once)
static int contract_impl(int dtype, int algo, int B, int P, int n_out, int n_in, const void* W, int64_t w_sp,
                         int64_t w_si, int64_t w_sk, const void* x, int64_t x_sb, int64_t x_sp, int64_t x_sk,
                         void* out, int64_t o_sb, int64_t o_sp, int64_t o_si, int accumulate, void* ws,
                         size_t ws_bytes, cudaStream_t s, bool w_prepared = false) {
    if (B <= 0 || P <= 0 || n_out <= 0 || n_in <= 0) DADMM_FAIL(-1, "contract: bad dims");
    if (!W || !x || !out) DADMM_FAIL(-1, "contract: null pointer");
    if (dtype != DADMM_F32 && dtype != DADMM_F64) DADMM_FAIL(-1, "contract: unknown dtype %d", dtype);
    if (dtype == DADMM_F64) {
        if (algo >= DADMM_ALGO_TC_3XTF32) DADMM_FAIL(-4, "contract: tensor-core paths are fp32 only");
        GemmParams<double> p{B, P, n_out, n_in, (const double*)W, w_sp, w_si, w_sk, (const double*)x, x_sb, x_sp, x_sk,
                             (double*)out, o_sb, o_sp, o_si, accumulate};
        return launch_contract_simt<double>(p, s);
    }
    const bool tc_layout = tc::shape_supported(std::max(B, 128), P, std::max(n_out, 64), std::max(n_in, 16), W, w_sp, w_si, w_sk, x,
                                               x_sb, x_sp, x_sk, out, o_sb, o_sp, o_si);
    const int ra = resolve_algo(dtype, algo, B, P, n_out, n_in, tc_layout && x_sb == (int64_t)P * n_in);
    if (ra < 0) DADMM_FAIL(-4, "contract: shape/layout not supported by the requested tensor-core kernel");
    if (ra == DADMM_ALGO_TC_3XF16 || ra == DADMM_ALGO_TC_F16X1) {
        const size_t wb = f16::split_bytes((long long)P * n_out, n_in), xb = f16::split_bytes((long long)B * P, n_in);
        if (!ws || ws_bytes < wb + xb) DADMM_FAIL(-1, "contract: workspace too small for the fp16 operand copies");
        char* c = (char*)ws;
        if (!w_prepared)
            if (int e = f16::split_tensor((const float*)W, (long long)P * n_out, n_in, n_in, c, s)) return e;
        if (int e = f16::split_tensor((const float*)x, (long long)B * P, n_in, n_in, c + wb, s)) return e;
        return f16::launch(B, P, n_out, n_in, c, c + wb, (float*)out, o_sb, accumulate, s, nullptr, nullptr, ra == DADMM_ALGO_TC_F16X1);
    }
    if (ra == DADMM_ALGO_TC_3XTF32)
        return tc::launch(B, P, n_out, n_in, (const float*)W, (const float*)x, (float*)out, x_sb, o_sb, accumulate, ws, ws_bytes, s);
    GemmParams<float> p{B, P, n_out, n_in, (const float*)W, w_sp, w_si, w_sk, (const float*)x, x_sb, x_sp, x_sk,
                        (float*)out, o_sb, o_sp, o_si, accumulate};
    return launch_contract_simt<float>(p, s);
}

// ------------------------------------------------------------------------------------------
// level kernels of the fused K-iteration path (unfolded.cuh)
// ------------------------------------------------------------------------------------------
// Tile configuration of a level kernel: `narr` [R][CH] tiles plus (when it fits) the staged neighbour lists of
// the tile's TB problems.  *list_cap == 0 means "lists stay in global memory".
static int level_cfg(int dtype, int B, int P, int n, int narr, int max_list, StepCfg* c, size_t* smem, int* list_cap,
                     int acc_rows = 0, bool fwd_only = false) {
    const size_t es = dtype == DADMM_F64 ? 8 : 4;
    // backward: same vec/TB/nchunks as step_bwd (the partial-sum layout follows them).  The forward level holds ONE tile and
    // leaves no partials, so it keeps 128-unknown chunks up to twice the agents (P = 100, BASELINE configs[4]: the two-tile
    // rule had it on 64-unknown chunks and the generic kernel, 3.7 ms per level instead of ~2)
    if (int e = fwd_only ? step_cfg(dtype, B, P, n, std::max(narr, 1), max_vec_for(dtype, n), c, true)
                         : step_cfg(dtype, B, P, n, 2, max_vec_for(dtype, n), c))
        return e;
    const size_t scal = ((size_t)4 * P + (size_t)c->TB * P) * es;       // staged per-agent scalars (stage_scalars)
    const size_t tiles = (size_t)narr * c->TB * P * 32 * c->vec * es + (size_t)acc_rows * c->TB * P * 4 * es + scal;
    const size_t lists = (size_t)c->TB * ((size_t)P + 1 + (size_t)std::max(max_list, 1)) * 4;
    if (narr > 0 && tiles + lists <= 110 * 1024) {
        *list_cap = std::max(max_list, 1);
        *smem = tiles + lists;
    } else {
        *list_cap = 0;
        *smem = tiles + (size_t)c->TB * ((size_t)P + 1) * 4;           // sPtr slot stays addressable
    }
    return 0;
}

// CTAs per problem group of the level kernels: 1 (a CTA walks all chunks of its problems) unless the
// batch is too small to fill the machine
static int level_bwd_csplit(int B, int TB, int nchunks) {
    const int groups = (B + TB - 1) / TB;
    static const int forced = [] { const char* e = getenv("DADMM_BWD_CSPLIT"); return e ? atoi(e) : 0; }();   // experiment knob
    if (forced > 0) return std::max(1, std::min(nchunks, forced));
    return std::max(1, std::min(nchunks, (148 * 12 + groups - 1) / groups));
}

// backward level: ten warps per CTA when that splits the tile's rows evenly and eight do not (P = 50: 1.361 -> 1.287 ms on
// B200; DADMM_LEVEL_WARPS=8 keeps eight)
static bool level_ten_warps(int rows) {
    static const int forced = [] { const char* e = getenv("DADMM_LEVEL_WARPS"); return e ? atoi(e) : 0; }();
    if (forced == 8) return false;
    if (forced == 10) return true;
    return (rows % 8) != 0 && (rows % 10) == 0;
}

// 1: the lean forward level accumulates 2L y in the reference's event order (bit-identical delta, twice the gathers);
// 0 (default): each neighbour difference once, doubled.  DADMM_EXACT_ORDER=1 / dadmm_set_consensus_order(1) select the former.
static bool exact_order() {
    int v = g_exact_order.load(std::memory_order_relaxed);
    if (v < 0) {
        const char* e = getenv("DADMM_EXACT_ORDER");
        v = (e && atoi(e) == 1) ? 1 : 0;
        g_exact_order.store(v, std::memory_order_relaxed);
    }
    return v != 0;
}

// second-generation lean level kernels (unfolded_lean.cuh: packed fp32 arithmetic, byte-offset lists); DADMM_LEVEL_GEN=1
// keeps the first generation (A/B measurements)
static bool lean_gen2() {
    static const bool on = [] { const char* e = getenv("DADMM_LEVEL_GEN"); return !(e && atoi(e) == 1); }();
    return on;
}

// CTAs per SM the lean kernels' registers are budgeted for: DADMM_LEAN_MINB_FWD / _BWD override the defaults
// (forward, B200, cfg4, ms per training step in the forward levels: event-order gather 24.5 at 4 CTAs x 64 registers / 25.8
// at 5 x 48 with spills; half-length gather 23.6 at 4 / 22.7 at 5 -- with half the gathers in flight the extra warps win)
static int lean_minb(bool fwd, bool exact = false) {
    static const int f = [] { const char* e = getenv("DADMM_LEAN_MINB_FWD"); return e ? atoi(e) : 0; }();
    static const int b = [] { const char* e = getenv("DADMM_LEAN_MINB_BWD"); return e ? atoi(e) : 0; }();
    return fwd ? (f ? f : (exact ? 4 : 5)) : (b ? b : 4);
}
template <typename K, typename Prm>
static int launch_level(K kernel, int grid, int threads, size_t smem, cudaStream_t s, const Prm& p) {
    if (int e = allow_smem(kernel, smem)) return e;
    DADMM_CUDA(launch_chain(kernel, dim3(grid), dim3(threads), smem, s, p));
    return 0;
}

// software prefetch of the warp's next row in the lean level kernels (0 off, 1 into L1, 2 into L2):
// DADMM_LEVEL_PREFETCH_FWD / _BWD override the defaults
static int level_prefetch(bool fwd) {
    static const int f = [] { const char* e = getenv("DADMM_LEVEL_PREFETCH_FWD"); return e ? atoi(e) : -1; }();
    static const int b = [] { const char* e = getenv("DADMM_LEVEL_PREFETCH_BWD"); return e ? atoi(e) : -1; }();
    return fwd ? (f >= 0 ? f : 2) : (b >= 0 ? b : 0);
}
static int sm_count() {
    static const int n = [] {
        int dev = 0, v = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
        return v;
    }();
    return n;
}
// forward level as a persistent TMA pipeline (unfolded_pipe.cuh); DADMM_FWD_PIPE=0 keeps the occupancy-driven kernels
// (measured with one 512-byte bulk copy per tile row: 1.34 ms per level against 0.95 -- the copy engine, not HBM, paces
// 150 small copies per tile; off by default, DADMM_FWD_PIPE=1 selects it)
static bool fwd_pipe_enabled() {
    static const bool on = [] { const char* e = getenv("DADMM_FWD_PIPE"); return e && atoi(e) == 1; }();
    return on;
}
// problems per tile of the pipelined kernel: as many as keep the tile at or below 64 rows, dividing the batch
static int pipe_tb(int B, int P) {
    static const int forced = [] { const char* e = getenv("DADMM_PIPE_TB"); return e ? atoi(e) : 0; }();
    int best = 1;
    for (int d = 1; d * P <= 64 && d <= B; ++d)
        if (B % d == 0) best = d;
    if (forced > 0 && B % forced == 0) best = forced;
    return best;
}

template <typename T>
static int level_fwd_impl(int dtype, int B, int P, int n, const dadmm_graph* g, const dadmm_clamps* cl_k,
                          const dadmm_clamps* cl_prev, const void* hyp_k, const void* hyp_prev, const void* y,
                          const void* U_in, const void* d0, const void* a, const void* atb, void* y_next, void* U_out,
                          void* graw, int32_t* flags, const SplitOut& sp, cudaStream_t s, void* agent_sum = nullptr,
                          double* sq_part = nullptr, int* sums_grid = nullptr) {
    LevelFwdParams<T> p;
    p.sp = sp;
    p.B = B; p.P = P; p.n = n; p.first = (hyp_prev == nullptr);
    p.lst_ptr = g->ev_ptr; p.lst_idx = g->ev_idx; p.deg = g->deg; p.gid = g->graph_id;
    p.exact_order = 1;
    p.prefetch = level_prefetch(true);
    p.hyp_k = (const T*)hyp_k; p.hyp_prev = (const T*)hyp_prev;
    p.G = (T)cl_k->G; p.V = (T)cl_k->V;
    p.hasD = std::isfinite(cl_k->D) ? 1 : 0;
    p.D = p.hasD ? (T)cl_k->D : std::numeric_limits<T>::infinity();
    p.Uc_prev = cl_prev ? (T)cl_prev->Uc : (T)0;
    p.y = (const T*)y; p.U_in = (const T*)U_in; p.d0 = (const T*)d0; p.a = (const T*)a; p.atb = (const T*)atb;
    p.y_next = (T*)y_next; p.U_out = (T*)U_out; p.graw = (T*)graw; p.flags = flags;
    p.agent_sum = nullptr; p.sq_part = nullptr;
    StepCfg c;
    size_t smem = 0;
    if (int e = level_cfg(dtype, B, P, n, p.first ? 0 : 1, g->max_events, &c, &smem, &p.list_cap, 0, true)) return e;
    if (int e = check_aligned(dtype, c.vec, {y, U_in, d0, a, atb, y_next, U_out, graw})) return e;
    p.TB = c.TB;
    // forward: two 128-unknown chunks per CTA (B200, cfg4, lean kernel: 1 chunk 0.940 ms, 2 -> 0.916, 4 -> 0.919, 8 -> 0.925;
    // the neighbour lists and per-agent scalars are staged once per CTA; register prefetch of the next row: 0.98 ms).
    // DADMM_FWD_CHUNKS_PER_CTA overrides.
    {
        static const int forced = [] { const char* e = getenv("DADMM_FWD_CHUNKS_PER_CTA"); return e ? atoi(e) : 0; }();
        const int groups = (B + c.TB - 1) / c.TB;
        // (one chunk per CTA while two would leave SMs without a CTA)
        const int cpc = forced > 0 ? forced : ((long long)groups * ((c.nchunks + 1) / 2) < 148 ? 1 : 2);
        p.csplit = std::max(1, (c.nchunks + cpc - 1) / cpc);
    }
    c.grid = ((B + c.TB - 1) / c.TB) * p.csplit;
    ProfScope prof(PROF_STEP_FWD, s);
#define DADMM_LAUNCH_LFWD(VEC, LEAN, NTHR)                                                      \
    {                                                                                           \
        if (int e = allow_smem(level_fwd_kernel<T, VEC, LEAN, NTHR>, smem)) return e;           \
        DADMM_CUDA(launch_chain(level_fwd_kernel<T, VEC, LEAN, NTHR>, dim3(c.grid), dim3(NTHR), smem, s, p)); \
    }
    bool launched = false;
    if constexpr (sizeof(T) == 4) {
        if (c.vec == 4) {
            // lean form: the fused fp16 path's configuration on full tiles (see level_fwd_kernel)
            const bool lean = !atb && !graw && !p.hasD && (n % 128) == 0 && (B % c.TB) == 0;
            // kernels of unfolded_lean.cuh / unfolded_pipe.cuh: which one can serve this launch
            const bool gen2 = lean && !p.first && lean_gen2();
            const int pipe_tbv = pipe_tb(B, P);
            const bool fast = gen2 && !exact_order();      // half-length consensus gather (lean::lap_half): plain adjacency lists
            const int list_max = fast ? g->max_adj : g->max_events;
            const bool can_pipe = gen2 && fwd_pipe_enabled() && pipe::fwd_smem_bytes(pipe_tbv, P, std::max(list_max, 1)) <= 227 * 1024;
            const bool can_lean2 = gen2 && p.list_cap > 0;
            // label-free loss sums (dadmm_loss_sums): lean form, problems at least as tall as the CTA has warps; several problems
            // per tile in the second-generation kernel only (level_fwd_kernel<LEAN> sums the whole tile as one problem)
            if (lean && agent_sum && sq_part && !p.first && P >= kStepThreads / 32 && (c.TB == 1 || (can_lean2 && !can_pipe))) {
                p.agent_sum = (T*)agent_sum;
                p.sq_part = sq_part;
                if (sums_grid) *sums_grid = c.grid;
            }
            if (fast && (can_pipe || can_lean2)) {
                p.exact_order = 0;
                p.lst_ptr = g->adj_ptr; p.lst_idx = g->adj_idx;
                if (p.list_cap > 0) p.list_cap = std::max(list_max, 1);      // never larger than the event lists the tile was sized for
            }
            if (can_pipe) {
                const int tb = pipe_tbv;
                const size_t psmem = pipe::fwd_smem_bytes(tb, P, std::max(list_max, 1));
                {
                    static const int ncons = [] { const char* e = getenv("DADMM_PIPE_WARPS"); return e ? atoi(e) : 16; }();
                    p.TB = tb;
                    p.list_cap = std::max(list_max, 1);
                    p.csplit = 1;
                    const long long tiles = (long long)(B / tb) * (n / 128);
                    const int grid = (int)std::min<long long>(sm_count(), tiles);
                    if (agent_sum && sq_part) {
                        p.agent_sum = (T*)agent_sum;
                        p.sq_part = sq_part;
                        if (sums_grid) *sums_grid = grid;
                    }
                    // tiled tensor maps of the three input tensors ([B*P, n] fp32, box [tile rows] x [128 unknowns])
                    static const bool tmap = [] { const char* e = getenv("DADMM_PIPE_TMAP"); return !(e && atoi(e) == 0); }();
                    CUtensorMap maps[3];
                    memset(maps, 0, sizeof(maps));
                    bool use_tmap = tmap && tb * P <= 256 && tc::encode_fn() != nullptr;
                    if (use_tmap) {
                        const void* src[3] = {y, a, U_in};
                        for (int q = 0; q < 3 && use_tmap; ++q) {
                            cuuint64_t dims[2] = {(cuuint64_t)n, (cuuint64_t)B * P}, strides[1] = {(cuuint64_t)n * 4};
                            cuuint32_t box[2] = {128, (cuuint32_t)(tb * P)}, es[2] = {1, 1};
                            use_tmap = tc::encode_fn()(&maps[q], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)src[q], dims, strides, box, es,
                                                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                                       CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
                        }
                    }
#define DADMM_LAUNCH_PIPE(NC, TM)                                                                                         \
    {                                                                                                                      \
        if (int e = allow_smem(pipe::level_fwd_pipe_kernel<NC, TM>, psmem)) return e;                                     \
        DADMM_CUDA(launch_chain(pipe::level_fwd_pipe_kernel<NC, TM>, dim3(grid), dim3((NC + 1) * 32), psmem, s, p, maps[0], maps[1], maps[2])); \
    }
                    if (use_tmap) {
                        if (ncons == 8) DADMM_LAUNCH_PIPE(8, true)
                        else if (ncons == 12) DADMM_LAUNCH_PIPE(12, true)
                        else DADMM_LAUNCH_PIPE(16, true)
                    } else {
                        if (ncons == 8) DADMM_LAUNCH_PIPE(8, false)
                        else if (ncons == 12) DADMM_LAUNCH_PIPE(12, false)
                        else DADMM_LAUNCH_PIPE(16, false)
                    }
#undef DADMM_LAUNCH_PIPE
                    DADMM_LAUNCHED();
                    return 0;
                }
            }
            if (can_lean2) {
                const int mb = lean_minb(true, p.exact_order != 0);
                if (int e = mb == 3   ? launch_level(lean::level_fwd_lean_kernel<kStepThreads, 3>, c.grid, kStepThreads, smem, s, p)
                            : mb == 4 ? launch_level(lean::level_fwd_lean_kernel<kStepThreads, 4>, c.grid, kStepThreads, smem, s, p)
                                      : launch_level(lean::level_fwd_lean_kernel<kStepThreads, 5>, c.grid, kStepThreads, smem, s, p))
                    return e;
            }
            else if (lean) DADMM_LAUNCH_LFWD(4, true, kStepThreads)      // (ten warps: 0.922 vs 0.913 ms -- no gain in the forward level; the same for the second-generation kernel, round 2 call AG)
            else DADMM_LAUNCH_LFWD(4, false, kStepThreads)
            launched = true;
        }
    }
    if (!launched) {
        if (c.vec == 2) DADMM_LAUNCH_LFWD(2, false, kStepThreads)
        else DADMM_LAUNCH_LFWD(1, false, kStepThreads)
    }
#undef DADMM_LAUNCH_LFWD
    DADMM_LAUNCHED();
    return 0;
}

template <typename T>
static int level_bwd_impl(int dtype, int B, int P, int n, const dadmm_graph* g, const dadmm_clamps* cl_k,
                          const dadmm_clamps* cl_prev, const void* hyp_k, const void* hyp_prev, int top, const void* y,
                          const void* U_prev, const void* d0, const void* graw, void* Tb, void* C, void* ga,
                          const void* gY_prev, const void* label, double coef_prev, void* partials, const SplitOut& sp,
                          int graw_is_residual, cudaStream_t s, const double* coef_dev = nullptr) {
    LevelBwdParams<T> p;
    p.sp = sp;
    p.graw_is_residual = graw_is_residual;
    p.prefetch = level_prefetch(false);
    p.B = B; p.P = P; p.n = n; p.first = (hyp_prev == nullptr); p.top = top;
    p.lst_ptr = g->adj_ptr; p.lst_idx = g->adj_idx; p.deg = g->deg; p.gid = g->graph_id;
    p.hyp_k = (const T*)hyp_k; p.hyp_prev = (const T*)hyp_prev;
    p.G = (T)cl_k->G; p.V = (T)cl_k->V;
    p.hasD = std::isfinite(cl_k->D) ? 1 : 0;
    p.D = p.hasD ? (T)cl_k->D : std::numeric_limits<T>::infinity();
    p.Uc_prev = cl_prev ? (T)cl_prev->Uc : (T)0;
    p.y = (const T*)y; p.U_prev = (const T*)U_prev; p.d0 = (const T*)d0; p.graw = (const T*)graw;
    p.Tb = (T*)Tb; p.C = (T*)C; p.ga = (T*)ga;
    p.gY_prev = (const T*)gY_prev;
    p.label = (coef_prev != 0.0 || coef_dev) ? (const T*)label : nullptr;
    p.coef_prev = (T)coef_prev;
    p.coef_dev = p.label ? coef_dev : nullptr;
    p.partials = (T*)partials;
    StepCfg c;
    size_t smem = 0;
    if (int e = level_cfg(dtype, B, P, n, 2, g->max_adj, &c, &smem, &p.list_cap, 1)) return e;
    if (int e = check_aligned(dtype, c.vec, {y, U_prev, d0, graw, Tb, C, ga, gY_prev, label})) return e;
    p.TB = c.TB;
    p.csplit = level_bwd_csplit(B, c.TB, c.nchunks);
    c.grid = ((B + c.TB - 1) / c.TB) * p.csplit;
    ProfScope prof(PROF_STEP_BWD, s);
#define DADMM_LAUNCH_LBWD(VEC, LEAN, NTHR)                                                      \
    {                                                                                           \
        if (int e = allow_smem(level_bwd_kernel<T, VEC, LEAN, NTHR>, smem)) return e;           \
        DADMM_CUDA(launch_chain(level_bwd_kernel<T, VEC, LEAN, NTHR>, dim3(c.grid), dim3(NTHR), smem, s, p)); \
    }
    bool launched = false;
    if constexpr (sizeof(T) == 4) {
        if (c.vec == 4) {
            // lean form: the fused fp16 training path on full tiles, levels k >= 1 (see level_bwd_kernel)
            const bool lean = !p.first && graw_is_residual && !p.hasD && !gY_prev && (n % 128) == 0 && (B % c.TB) == 0;
            static const bool bwd_gen1 = [] { const char* e = getenv("DADMM_BWD_GEN"); return e && atoi(e) == 1; }();
            // level k = 0 of the fused training path: a streaming reduction (no gather, nothing propagated)
            const bool lean_first = p.first && graw_is_residual && !p.hasD && !gY_prev && (n % 128) == 0 && (B % c.TB) == 0 && lean_gen2() && !bwd_gen1;
            if (lean_first) {
                if (int e = launch_level(lean::level_bwd_first_lean_kernel<kStepThreads>, c.grid, kStepThreads, smem, s, p)) return e;
            }
            else if (lean && p.list_cap > 0 && lean_gen2() && !bwd_gen1) {
                const int mb = lean_minb(false);
                if (level_ten_warps(c.TB * P)) {
                    if (int e = mb <= 2 ? launch_level(lean::level_bwd_lean_kernel<320, 2>, c.grid, 320, smem, s, p)
                                        : launch_level(lean::level_bwd_lean_kernel<320, 3>, c.grid, 320, smem, s, p))
                        return e;
                } else {
                    if (int e = mb <= 3 ? launch_level(lean::level_bwd_lean_kernel<kStepThreads, 3>, c.grid, kStepThreads, smem, s, p)
                                        : launch_level(lean::level_bwd_lean_kernel<kStepThreads, 4>, c.grid, kStepThreads, smem, s, p))
                        return e;
                }
            }
            else if (lean && level_ten_warps(c.TB * P)) DADMM_LAUNCH_LBWD(4, true, 320)
            else if (lean) DADMM_LAUNCH_LBWD(4, true, kStepThreads)
            else DADMM_LAUNCH_LBWD(4, false, kStepThreads)
            launched = true;
        }
    }
    if (!launched) {
        if (c.vec == 2) DADMM_LAUNCH_LBWD(2, false, kStepThreads)
        else DADMM_LAUNCH_LBWD(1, false, kStepThreads)
    }
#undef DADMM_LAUNCH_LBWD
    DADMM_LAUNCHED();
    return 0;
}

// partial-sum rows of the reverse sweep: K levels x [csplit][B][P][4] (the chunking level_bwd_impl picks), reduced by one
// launch after the sweep
static size_t sweep_part_elems(int dtype, int B, int P, int n, int K, size_t* level_stride = nullptr, int* nchunks = nullptr) {
    StepCfg c;
    if (step_cfg(dtype, B, P, n, 2, max_vec_for(dtype, n), &c)) return 0;
    const int cs = level_bwd_csplit(B, c.TB, c.nchunks);
    const size_t stride = (size_t)cs * B * P * 4;
    if (level_stride) *level_stride = stride;
    if (nchunks) *nchunks = cs;
    return stride * (size_t)std::max(K, 1);
}

static size_t amax_slots_bytes(int K) { return ((size_t)(K + 1) * 4 + 255) / 256 * 256; }

// per-CTA sum-of-squares partials of the forward levels (dadmm_loss_sums): K rows of one double per CTA
static size_t sq_part_bytes(int dtype, int B, int P, int n, int K) {
    StepCfg c;
    if (dtype != DADMM_F32 || step_cfg(dtype, B, P, n, 2, max_vec_for(dtype, n), &c)) return 0;
    const size_t grid = std::max<size_t>((size_t)((B + c.TB - 1) / c.TB) * c.nchunks, 192);   // upper bound of the level grid
    return ((size_t)K * grid * sizeof(double) + 255) / 256 * 256;
}

// the fused fp16 path: operands of the contraction are produced already split by the level kernels
static bool fused_f16(int dtype, int algo, int B, int P, int n) {
    if (dtype != DADMM_F32 || (n % 8) != 0) return false;
    const int ra = resolve_algo(dtype, algo, B, P, n, n, true);
    return ra == DADMM_ALGO_TC_3XF16 || ra == DADMM_ALGO_TC_F16X1;
}

// two-stage contraction W x = F2 (F1 x) of the fused path: only where it removes a quarter of the flops or more
// and both stages tile (DADMM_TWO_STAGE=0 forces the single-stage route, for A/B measurements)
static bool use_factor(int dtype, int algo, int B, int P, int n, int m) {
    static const bool enabled = [] {
        const char* e = getenv("DADMM_TWO_STAGE");
        return !(e && !strcmp(e, "0"));
    }();
    if (!enabled || m <= 0 || (m % 8) != 0 || 8LL * m > 3LL * n) return false;
    return fused_f16(dtype, algo, B, P, n) && f16::dims_supported(B, P, m, n) && f16::dims_supported(B, P, n, m);
}

// operand-copy region at the head of the fused path's workspace
struct FusedWs {
    bool two;
    size_t w1, w2, xb, tb;       // split(W) | split(F1), split(F2); split(x); split(t)
    size_t total() const { return w1 + w2 + xb + tb; }
};
static FusedWs fused_ws(int dtype, int algo, int B, int P, int n, int m) {
    FusedWs f{};
    if (!fused_f16(dtype, algo, B, P, n)) return f;
    f.two = use_factor(dtype, algo, B, P, n, m);
    f.xb = f16::split_bytes((long long)B * P, n);
    if (f.two) {
        f.w1 = f16::split_bytes((long long)P * m, n);
        f.w2 = f16::split_bytes((long long)P * n, m);
        f.tb = f16::split_bytes((long long)B * P, m);
    } else {
        f.w1 = f16::split_bytes((long long)P * n, n);
    }
    return f;
}
static size_t unfolded_cw(int dtype, int algo, int B, int P, int n, int m) {
    const FusedWs f = fused_ws(dtype, algo, B, P, n, m);
    const size_t c = f.total() ? f.total() : contract_ws_bytes(dtype, algo, B, P, n, n);
    return (c + 255) / 256 * 256;
}

// one contraction of the fused path from prepared operands: out (+)= W x (- sub); with `rhs` (two-stage only) the
// subtracted term enters the first stage instead: out = F2 (F1 x - rhs)
// (w8 = operator split(s): head of the workspace, or the caller's persistent dadmm_op_split buffer; xsp = split of x)
static int fused_contract(const FusedWs& f, char* w8, char* xsp, int B, int P, int n, int m, float* out, int accumulate, cudaStream_t s,
                          unsigned* amax_out, const float* sub, int fast, const float* rhs = nullptr) {
    if (!f.two) return f16::launch(B, P, n, n, w8, xsp, out, (int64_t)P * n, accumulate, s, amax_out, sub, fast);
    char* tsp = xsp + f.xb;
    const unsigned* rhs_amax = (const unsigned*)(tsp + 48);       // spare scalar slot of the t split header
    // partial-sum lengths: 128 k in both stages (3.7e-7 rel-L2 each, composite 5.2e-7 against 4.1e-7 for the exact-FMA
    // kernel at n = 1024 and a 1e-5 tolerance).  Round 1 ran the first stage with 64-k partial sums (composite 4.0e-7); the
    // 128 KB TMEM drain per partial sum paces that stage, and halving the number of drains takes it from 0.33 to 0.29 ms
    // (16.0 -> 14.5 ms per training step at cfg4; DADMM_F16_KBC overrides the length for both stages)
    if (int e = f16::launch(B, P, m, n, w8, xsp, nullptr, (int64_t)P * m, 0, s, nullptr, rhs, fast, tsp, 2, rhs ? rhs_amax : nullptr))
        return e;
    return f16::launch(B, P, n, m, w8 + f.w1, tsp, out, (int64_t)P * n, accumulate, s, amax_out, rhs ? nullptr : sub, fast, nullptr, 2);
}
static int fused_prepare(const FusedWs& f, char* w8, int P, int n, int m, const void* W, const dadmm_factor* fac, cudaStream_t s) {
    if (!f.two) return f16::split_tensor((const float*)W, (long long)P * n, n, n, w8, s);
    if (int e = f16::split_tensor((const float*)fac->F1, (long long)P * m, n, n, w8, s, true)) return e;
    return f16::split_tensor((const float*)fac->F2, (long long)P * n, m, m, w8 + f.w1, s);
}

static int check_op_split(const dadmm_op_split* ops, int dtype, int algo, int B, int P, int n, int m) {
    if (!ops || !ops->buf) return 0;
    const FusedWs f = fused_ws(dtype, algo, B, P, n, m);
    if (ops->bytes < f.w1 + f.w2) DADMM_FAIL(-1, "op_split: buffer of %zu bytes, %zu needed", ops->bytes, f.w1 + f.w2);
    if (!aligned_to(ops->buf, 256)) DADMM_FAIL(-5, "op_split: buffer must be 256-byte aligned");
    return 0;
}

template <typename T>
static int unfolded_fwd_impl(int dtype, int algo, int B, int P, int n, int K, const dadmm_graph* graph,
                             const dadmm_clamps* clamps, const void* hyp, const void* W, const dadmm_factor* fac,
                             const void* Atb, const void* y0,
                             const void* U0, const void* d0, void* Y, void* U_save, void* R_save, void* ws, int32_t* flags,
                             const dadmm_loss_sums* sums, const dadmm_op_split* ops, cudaStream_t s) {
    const size_t es = sizeof(T);
    const size_t N = (size_t)B * P * n, NB = N * es;
    if (N >= (1ull << 31)) DADMM_FAIL(-1, "unfolded: B*P*n must stay below 2^31 per device (shard the batch)");
    const int mf = fac ? fac->m : 0;
    const size_t cw = unfolded_cw(dtype, algo, B, P, n, mf);
    const FusedWs fw = fused_ws(dtype, algo, B, P, n, mf);
    const size_t NBa = (NB + 255) / 256 * 256;      // 256-byte aligned workspace regions
    char* w8 = (char*)ws;
    char* a = w8 + cw;
    char* pp[2] = {a + NBa, a + 2 * NBa};
    unsigned* slots = (unsigned*)(a + 3 * NBa);     // max|y_k| bits, k = 0..K
    double* sq_part = (double*)((char*)slots + amax_slots_bytes(K));         // [K][level grid] (dadmm_loss_sums)
    const size_t sq_row = sq_part_bytes(dtype, B, P, n, K) / sizeof(double) / (size_t)std::max(K, 1);
    int sums_grid = 0, sums_first = K;
    if (sums)
        for (int k = 0; k < K; ++k) sums->valid[k] = 0;
    const int64_t sn = n, sPn = (int64_t)P * n;
    const size_t row = (size_t)P * 4 * es;
    const bool fused = fused_f16(dtype, algo, B, P, n);
    const size_t wb = fw.w1 + fw.w2;
    f16::Split xs{};
    char* wsp = (fused && ops && ops->buf) ? (char*)ops->buf : w8;       // operator split: persistent buffer or workspace head
    if (fused) {
        xs = f16::split_view(w8 + wb, (long long)B * P, n);
        DADMM_CUDA(cudaMemsetAsync(slots, 0, amax_slots_bytes(K), s));
        if (!(ops && ops->buf && ops->ready))
            if (int e = fused_prepare(fw, wsp, P, n, mf, W, fac, s)) return e;
        if (int e = f16::split_tensor((const float*)y0, (long long)B * P, n, n, w8 + wb, s)) return e;
        if (fw.two && fac->rhs)
            if (int e = f16::amax_tensor((const float*)fac->rhs, (long long)B * P * mf, (unsigned*)(w8 + wb + fw.xb + 48), s)) return e;
    }
    const float* rhs = (fused && fw.two) ? (const float*)fac->rhs : nullptr;
    for (int k = 0; k < K; ++k) {
        const char* yk = k ? (const char*)Y + (size_t)(k - 1) * NB : (const char*)y0;
        // U_j for j >= 1 lives in U_save[j-1] (training) or in a ping-pong buffer (inference)
        auto Uslot = [&](int j) -> char* { return U_save ? (char*)U_save + (size_t)(j - 1) * NB : pp[j & 1]; };
        const char* Uin = (k <= 1) ? (const char*)U0 : Uslot(k - 1);
        char* Uout = (k == 0 || k == K - 1) ? nullptr : Uslot(k);   // U_{K-1} is consumed in-register only
        SplitOut sp{};
        // fused path: the contraction epilogue subtracts Atb and writes a_k' = AtA y_k - Atb straight into the saved
        // stream (R_save[k]) -- the forward level neither reads Atb nor writes r_k; the backward rebuilds r_k from a_k'
        char* ak = (fused && R_save) ? (char*)R_save + (size_t)k * NB : a;
        if (fused) {
            if (int e = fused_contract(fw, wsp, w8 + wb, B, P, n, mf, (float*)ak, 0, s, nullptr, (const float*)Atb, algo == DADMM_ALGO_TC_F16X1, rhs))
                return e;
            if (k < K - 1) sp = SplitOut{xs.hi, xs.lo, xs.exp, k ? slots + k : nullptr, slots + k + 1};
        } else {
            if (int e = contract_impl(dtype, algo, B, P, n, n, W, (int64_t)n * n, sn, 1, yk, sPn, sn, 1, a, sPn, sn, 1, 0, w8, cw, s, k > 0))
                return e;
        }
        if (int e = level_fwd_impl<T>(dtype, B, P, n, graph, clamps + k, k ? clamps + k - 1 : nullptr, (const char*)hyp + k * row,
                                      k ? (const char*)hyp + (k - 1) * row : nullptr, yk, Uin, d0, ak, fused ? nullptr : Atb,
                                      (char*)Y + (size_t)k * NB, Uout,
                                      (R_save && !fused) ? (char*)R_save + (size_t)k * NB : nullptr,
                                      flags ? flags + k : nullptr, sp, s,
                                      (sums && fused && sq_row) ? (char*)sums->agent_sum + (size_t)k * B * n * es : nullptr,
                                      (sums && fused && sq_row) ? sq_part + (size_t)k * sq_row : nullptr, &sums_grid))
            return e;
        if (sums && sums_grid > 0 && k >= 1) {        // the level produced the sums (its grid is the same for every k >= 1)
            sums->valid[k] = 1;
            sums_first = std::min(sums_first, k);
        }
    }
    if (sums && sums_first < K) {
        DADMM_CUDA(cudaMemsetAsync(sums->sumsq, 0, (size_t)K * sizeof(double), s));
        ProfScope prof(PROF_LOSS, s);
        DADMM_CUDA(launch_chain(sumsq_final_kernel, dim3(K - sums_first), dim3(256), 0, s, (const double*)sq_part, sq_row, sums_grid,
                                sums_first, sums->sumsq));
        DADMM_LAUNCHED();
    }
    return 0;
}

template <typename T>
static int unfolded_bwd_impl(int dtype, int algo, int B, int P, int n, int K, const dadmm_graph* graph,
                             const dadmm_clamps* clamps, const void* hyp, const void* Wt, const dadmm_factor* fac,
                             const void* y0, const void* U0,
                             const void* d0, const void* Y, const void* U_save, const void* R_save, const void* gY,
                             const void* label, const double* loss_coef, const double* loss_coef_dev, void* ghyp, void* ws,
                             const dadmm_op_split* ops, cudaStream_t s) {
    const size_t es = sizeof(T);
    const size_t N = (size_t)B * P * n, NB = N * es;
    const int mf = fac ? fac->m : 0;
    const size_t cw = unfolded_cw(dtype, algo, B, P, n, mf);
    const FusedWs fw = fused_ws(dtype, algo, B, P, n, mf);
    const size_t NBa = (NB + 255) / 256 * 256;      // 256-byte aligned workspace regions
    char* w8 = (char*)ws;
    char *Tb = w8 + cw, *C = Tb + NBa, *ga = C + NBa, *part = ga + NBa;
    size_t part_stride = 0;              // elements between the partial rows of consecutive levels
    int nchunks = 0;                     // partial-sum rows per (problem, agent)
    const size_t part_elems = sweep_part_elems(dtype, B, P, n, K, &part_stride, &nchunks);
    if (!part_elems) DADMM_FAIL(-2, "unfolded_bwd: P=%d does not fit in shared memory", P);
    unsigned* slots = (unsigned*)(part + (part_elems * es + 255) / 256 * 256);   // max|adj(y_{k+1})| bits
    const int64_t sn = n, sPn = (int64_t)P * n;
    const size_t row = (size_t)P * 4 * es;
    // fused loss term coef[k] (Y[k] - label): coefficients on the host (loss_coef) or on the device (loss_coef_dev)
    const bool with_loss = label && (loss_coef || loss_coef_dev);
    const bool dev_coef = label && !loss_coef && loss_coef_dev;
    auto host_coef = [&](int k) { return (with_loss && !dev_coef) ? loss_coef[k] : 0.0; };
    const bool fused = fused_f16(dtype, algo, B, P, n);
    const size_t wb = fw.w1 + fw.w2;
    f16::Split xs{};
    char* wsp = (fused && ops && ops->buf) ? (char*)ops->buf : w8;
    DADMM_CUDA(cudaMemsetAsync(ghyp, 0, (size_t)K * row, s));
    if (fused) {
        xs = f16::split_view(w8 + wb, (long long)B * P, n);
        DADMM_CUDA(cudaMemsetAsync(slots, 0, amax_slots_bytes(K), s));
        if (K > 1 && !(ops && ops->buf && ops->ready))
            if (int e = fused_prepare(fw, wsp, P, n, mf, Wt, fac, s)) return e;
    }
    {   // adjoint of y_K
        const long long rows = (long long)B * P;
        const int nblk = (int)std::min<long long>(148 * 8, ceil_div64(rows, 8));
        ProfScope prof(PROF_LOSS, s);
        DADMM_CUDA(launch_chain(seed_adjoint_kernel<T>, dim3(nblk), dim3(256), 0, s, (const T*)((const char*)Y + (size_t)(K - 1) * NB),
                                gY ? (const T*)((const char*)gY + (size_t)(K - 1) * NB) : (const T*)nullptr,
                                (dev_coef || host_coef(K - 1) != 0.0) ? (const T*)label : (const T*)nullptr, (T)host_coef(K - 1),
                                dev_coef ? loss_coef_dev + (K - 1) : (const double*)nullptr, B, P, n, (T*)Tb,
                                fused ? slots + (K - 1) : (unsigned*)nullptr));
        DADMM_LAUNCHED();
    }
    for (int k = K - 1; k >= 0; --k) {
        const char* yk = k ? (const char*)Y + (size_t)(k - 1) * NB : (const char*)y0;
        const char* Uprev = (k <= 1) ? (const char*)U0 : (const char*)U_save + (size_t)(k - 2) * NB;   // U_{k-1}
        SplitOut sp{};
        if (fused && k > 0) sp = SplitOut{xs.hi, xs.lo, xs.exp, slots + k, nullptr};
        if (int e = level_bwd_impl<T>(dtype, B, P, n, graph, clamps + k, k ? clamps + k - 1 : nullptr, (const char*)hyp + k * row,
                                      k ? (const char*)hyp + (k - 1) * row : nullptr, k == K - 1, yk, Uprev, d0,
                                      (const char*)R_save + (size_t)k * NB, Tb, C, ga,
                                      (gY && k) ? (const char*)gY + (size_t)(k - 1) * NB : nullptr, label,
                                      k ? host_coef(k - 1) : 0.0, part + (size_t)k * part_stride * es, sp, fused ? 1 : 0, s,
                                      (dev_coef && k) ? loss_coef_dev + (k - 1) : nullptr))
            return e;
        if (k > 0) {
            if (fused) {
                if (int e = fused_contract(fw, wsp, w8 + wb, B, P, n, mf, (float*)Tb, 1, s, slots + (k - 1), nullptr, algo == DADMM_ALGO_TC_F16X1))
                    return e;
            } else {
                if (int e = contract_impl(dtype, algo, B, P, n, n, Wt, (int64_t)n * n, sn, 1, ga, sPn, sn, 1, Tb, sPn, sn, 1, 1, w8, cw, s,
                                          k < K - 1))
                    return e;
            }
        }
    }
    {   // all K levels' partial sums -> ghyp [K,P,4]
        ProfScope prof(PROF_REDUCE_HYP, s);
        DADMM_CUDA(launch_chain(reduce_levels_kernel<T>, dim3(P, K), dim3(256), 0, s, (const T*)part, part_stride, nchunks, B, P, (T*)ghyp));
        DADMM_LAUNCHED();
    }
    return 0;
}

}  // namespace dadmm

using namespace dadmm;

extern "C" {

int dadmm_abi_version(void) { return DADMM_ABI_VERSION; }
const char* dadmm_last_error(void) { return g_err; }
int64_t dadmm_launch_count(void) { return (int64_t)g_launches.load(); }

int dadmm_set_consensus_order(int exact) {
    const int prev = exact_order() ? 1 : 0;
    g_exact_order.store(exact ? 1 : 0);
    return prev;
}

int dadmm_set_pdl(int on) {
    const int prev = pdl_enabled() ? 1 : 0;
    g_pdl.store(on ? 1 : 0);
    return prev;
}

int dadmm_profile_enable(int on) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    for (auto& r : g_prof) {
        cudaEventDestroy(r.a);
        cudaEventDestroy(r.b);
    }
    g_prof.clear();
    g_prof_on = on != 0;
    return 0;
}

int dadmm_profile_read(double* ms_by_kind, int64_t* launches_by_kind) {
    if (!ms_by_kind || !launches_by_kind) DADMM_FAIL(-1, "profile_read: null pointer");
    std::lock_guard<std::mutex> lk(g_prof_mu);
    for (int k = 0; k < PROF_KINDS; ++k) {
        ms_by_kind[k] = 0;
        launches_by_kind[k] = 0;
    }
    for (auto& r : g_prof) {
        DADMM_CUDA(cudaEventSynchronize(r.b));
        float ms = 0;
        DADMM_CUDA(cudaEventElapsedTime(&ms, r.a, r.b));
        ms_by_kind[r.kind] += ms;
        launches_by_kind[r.kind] += 1;
    }
    return 0;
}

int dadmm_device_check(void) {
    int dev = 0, major = 0, minor = 0;
    DADMM_CUDA(cudaGetDevice(&dev));
    DADMM_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    DADMM_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
    if (major != 10) DADMM_FAIL(-10, "libdadmm_sm100 needs an sm_100 device, found sm_%d%d", major, minor);
    return 0;
}

int dadmm_contract(int dtype, int algo, int B, int P, int n_out, int n_in, const void* W, int64_t w_sp, int64_t w_si,
                   int64_t w_sk, const void* x, int64_t x_sb, int64_t x_sp, int64_t x_sk, void* out, int64_t o_sb,
                   int64_t o_sp, int64_t o_si, int accumulate, void* ws, size_t ws_bytes, dadmm_stream_t stream) {
    return contract_impl(dtype, algo, B, P, n_out, n_in, W, w_sp, w_si, w_sk, x, x_sb, x_sp, x_sk, out, o_sb, o_sp, o_si,
                         accumulate, ws, ws_bytes, (cudaStream_t)stream);
}

int dadmm_contract_prepared(int dtype, int algo, int B, int P, int n_out, int n_in, const void* W, int64_t w_sp, int64_t w_si,
                            int64_t w_sk, const void* x, int64_t x_sb, int64_t x_sp, int64_t x_sk, void* out, int64_t o_sb,
                            int64_t o_sp, int64_t o_si, int accumulate, void* ws, size_t ws_bytes, int w_prepared,
                            dadmm_stream_t stream) {
    return contract_impl(dtype, algo, B, P, n_out, n_in, W, w_sp, w_si, w_sk, x, x_sb, x_sp, x_sk, out, o_sb, o_sp, o_si,
                         accumulate, ws, ws_bytes, (cudaStream_t)stream, w_prepared != 0);
}

size_t dadmm_contract_ws_bytes(int dtype, int algo, int B, int P, int n_out, int n_in) {
    return contract_ws_bytes(dtype, algo, B, P, n_out, n_in);
}

int dadmm_contract_uses_tensor_cores(int dtype, int algo, int B, int P, int n_out, int n_in) {
    const int ra = resolve_algo(dtype, algo, B, P, n_out, n_in, true);
    return (ra >= DADMM_ALGO_TC_3XTF32) ? ra : 0;
}

int dadmm_step_fwd(int dtype, int B, int P, int n, const dadmm_graph* graph, const dadmm_clamps* clamps,
                   const dadmm_hyp* hyp, const void* y, const void* U, const void* delta, const void* AtAy,
                   const void* Atb, void* y_next, void* U_next, void* delta_next, void* grad_raw, int32_t* flags,
                   dadmm_stream_t stream) {
    if (B <= 0 || P <= 0 || n <= 0) DADMM_FAIL(-1, "step_fwd: bad dims");
    if (!clamps || !hyp || !hyp->ptr || !y || !U || !AtAy || !Atb || !y_next) DADMM_FAIL(-1, "step_fwd: null pointer");
    if (int e = check_graph(graph, P)) return e;
    if (dtype == DADMM_F32)
        return step_fwd_impl<float>(dtype, B, P, n, graph, clamps, hyp, y, U, delta, AtAy, Atb, y_next, U_next, delta_next,
                                    grad_raw, flags, (cudaStream_t)stream);
    if (dtype == DADMM_F64)
        return step_fwd_impl<double>(dtype, B, P, n, graph, clamps, hyp, y, U, delta, AtAy, Atb, y_next, U_next, delta_next,
                                     grad_raw, flags, (cudaStream_t)stream);
    DADMM_FAIL(-1, "step_fwd: unknown dtype %d", dtype);
}

int dadmm_step_bwd(int dtype, int B, int P, int n, const dadmm_graph* graph, const dadmm_clamps* clamps,
                   const dadmm_hyp* hyp, const void* y, const void* U, const void* delta, const void* grad_raw,
                   const void* y_next, const void* gy_next_a, const void* gy_next_b, const void* gU_next,
                   const void* gdelta_next, const void* label, double loss_coef, void* gy, void* gAtAy, void* gU,
                   void* gdelta, void* ghyp_partials, dadmm_stream_t stream) {
    if (B <= 0 || P <= 0 || n <= 0) DADMM_FAIL(-1, "step_bwd: bad dims");
    if (!clamps || !hyp || !hyp->ptr || !y || !U || !grad_raw || !y_next || !gy || !gAtAy || !gU || !ghyp_partials)
        DADMM_FAIL(-1, "step_bwd: null pointer");
    if (int e = check_graph(graph, P)) return e;
    if (dtype == DADMM_F32)
        return step_bwd_impl<float>(dtype, B, P, n, graph, clamps, hyp, y, U, delta, grad_raw, y_next, gy_next_a, gy_next_b,
                                    gU_next, gdelta_next, label, loss_coef, gy, gAtAy, gU, gdelta, ghyp_partials,
                                    (cudaStream_t)stream);
    if (dtype == DADMM_F64)
        return step_bwd_impl<double>(dtype, B, P, n, graph, clamps, hyp, y, U, delta, grad_raw, y_next, gy_next_a, gy_next_b,
                                     gU_next, gdelta_next, label, loss_coef, gy, gAtAy, gU, gdelta, ghyp_partials,
                                     (cudaStream_t)stream);
    DADMM_FAIL(-1, "step_bwd: unknown dtype %d", dtype);
}

size_t dadmm_partials_elems(int dtype, int B, int P, int n) {
    (void)dtype;
    return partials_elems(B, P, n);
}

int dadmm_reduce_hyp(int dtype, int B, int P, int n, const void* ghyp_partials, int per_sample, void* ghyp,
                     int64_t stride_b, int64_t stride_p, int64_t stride_c, int accumulate, dadmm_stream_t stream) {
    if (!ghyp_partials || !ghyp) DADMM_FAIL(-1, "reduce_hyp: null pointer");
    int nchunks = 0;
    if (int e = bwd_nchunks(dtype, B, P, n, &nchunks)) return e;
    if (dtype == DADMM_F32)
        return reduce_hyp_impl<float>(dtype, B, P, n, ghyp_partials, per_sample, ghyp, stride_b, stride_p, stride_c,
                                      accumulate, nchunks, (cudaStream_t)stream);
    if (dtype == DADMM_F64)
        return reduce_hyp_impl<double>(dtype, B, P, n, ghyp_partials, per_sample, ghyp, stride_b, stride_p, stride_c,
                                       accumulate, nchunks, (cudaStream_t)stream);
    DADMM_FAIL(-1, "reduce_hyp: unknown dtype %d", dtype);
}

size_t dadmm_unfolded_op_split_bytes(int dtype, int algo, int B, int P, int n, int m_factor) {
    const FusedWs f = fused_ws(dtype, algo, B, P, n, m_factor);
    return f.w1 + f.w2;
}

int dadmm_unfolded_uses_factor(int dtype, int algo, int B, int P, int n, int m) { return use_factor(dtype, algo, B, P, n, m) ? 1 : 0; }

size_t dadmm_unfolded_ws_bytes(int dtype, int algo, int B, int P, int n, int K, int backward, int m_factor) {
    const size_t es = dtype == DADMM_F64 ? 8 : 4;
    const size_t NBa = ((size_t)B * P * n * es + 255) / 256 * 256;
    const size_t cw = unfolded_cw(dtype, algo, B, P, n, m_factor);
    if (!backward)   // AtAy + two U ping-pong buffers + amax slots + sum-of-squares partials
        return cw + 3 * NBa + amax_slots_bytes(K) + sq_part_bytes(dtype, B, P, n, K);
    return cw + 3 * NBa + (sweep_part_elems(dtype, B, P, n, K) * es + 255) / 256 * 256 + amax_slots_bytes(K);   // T, C, gAtAy, partials
}

int dadmm_unfolded_fwd(int dtype, int algo, int B, int P, int n, int K, const dadmm_graph* graph,
                       const dadmm_clamps* clamps, const void* hyp, const void* W, const dadmm_factor* factor,
                       const void* Atb, const void* y0,
                       const void* U0, const void* d0, void* Y, void* U_save, void* R_save, void* ws, size_t ws_bytes,
                       int32_t* flags, const dadmm_loss_sums* sums, const dadmm_op_split* op_split, dadmm_stream_t stream) {
    if (B <= 0 || P <= 0 || n <= 0 || K <= 0) DADMM_FAIL(-1, "unfolded_fwd: bad dims");
    if (!clamps || !hyp || !W || !y0 || !U0 || !d0 || !Y || !ws) DADMM_FAIL(-1, "unfolded_fwd: null pointer");
    if (factor && (factor->m <= 0 || !factor->F1 || !factor->F2)) DADMM_FAIL(-1, "unfolded_fwd: bad factor");
    // Atb may be omitted only where it is never read: the two-stage route with the observation term riding in stage 1
    if (!Atb && !(factor && factor->rhs && use_factor(dtype, algo, B, P, n, factor->m)))
        DADMM_FAIL(-1, "unfolded_fwd: Atb is required (NULL only with factor->rhs on the two-stage route)");
    if (sums && (!sums->agent_sum || !sums->sumsq || !sums->valid)) DADMM_FAIL(-1, "unfolded_fwd: bad loss-sums descriptor");
    if (ws_bytes < dadmm_unfolded_ws_bytes(dtype, algo, B, P, n, K, 0, factor ? factor->m : 0))
        DADMM_FAIL(-1, "unfolded_fwd: workspace too small");
    if (int e = check_graph(graph, P)) return e;
    if (int e = check_op_split(op_split, dtype, algo, B, P, n, factor ? factor->m : 0)) return e;
    if (dtype == DADMM_F32)
        return unfolded_fwd_impl<float>(dtype, algo, B, P, n, K, graph, clamps, hyp, W, factor, Atb, y0, U0, d0, Y, U_save, R_save, ws, flags,
                                        sums, op_split, (cudaStream_t)stream);
    if (dtype == DADMM_F64)
        return unfolded_fwd_impl<double>(dtype, algo, B, P, n, K, graph, clamps, hyp, W, factor, Atb, y0, U0, d0, Y, U_save, R_save, ws, flags,
                                         sums, op_split, (cudaStream_t)stream);
    DADMM_FAIL(-1, "unfolded_fwd: unknown dtype %d", dtype);
}

int dadmm_unfolded_bwd(int dtype, int algo, int B, int P, int n, int K, const dadmm_graph* graph,
                       const dadmm_clamps* clamps, const void* hyp, const void* Wt, const dadmm_factor* factor_t,
                       const void* y0, const void* U0,
                       const void* d0, const void* Y, const void* U_save, const void* R_save, const void* gY,
                       const void* label, const double* loss_coef, const double* loss_coef_dev, void* ghyp, void* ws,
                       size_t ws_bytes, const dadmm_op_split* op_split, dadmm_stream_t stream) {
    if (B <= 0 || P <= 0 || n <= 0 || K <= 0) DADMM_FAIL(-1, "unfolded_bwd: bad dims");
    if (!clamps || !hyp || !Wt || !y0 || !U0 || !d0 || !Y || !R_save || !ghyp || !ws) DADMM_FAIL(-1, "unfolded_bwd: null pointer");
    if (K > 2 && !U_save) DADMM_FAIL(-1, "unfolded_bwd: U_save required");
    if (factor_t && (factor_t->m <= 0 || !factor_t->F1 || !factor_t->F2)) DADMM_FAIL(-1, "unfolded_bwd: bad factor");
    if (ws_bytes < dadmm_unfolded_ws_bytes(dtype, algo, B, P, n, K, 1, factor_t ? factor_t->m : 0))
        DADMM_FAIL(-1, "unfolded_bwd: workspace too small");
    if (int e = check_graph(graph, P)) return e;
    if (int e = check_op_split(op_split, dtype, algo, B, P, n, factor_t ? factor_t->m : 0)) return e;
    if (dtype == DADMM_F32)
        return unfolded_bwd_impl<float>(dtype, algo, B, P, n, K, graph, clamps, hyp, Wt, factor_t, y0, U0, d0, Y, U_save, R_save, gY, label,
                                        loss_coef, loss_coef_dev, ghyp, ws, op_split, (cudaStream_t)stream);
    if (dtype == DADMM_F64)
        return unfolded_bwd_impl<double>(dtype, algo, B, P, n, K, graph, clamps, hyp, Wt, factor_t, y0, U0, d0, Y, U_save, R_save, gY, label,
                                         loss_coef, loss_coef_dev, ghyp, ws, op_split, (cudaStream_t)stream);
    DADMM_FAIL(-1, "unfolded_bwd: unknown dtype %d", dtype);
}

size_t dadmm_loss_ws_bytes(int dtype, int K, int B, int P, int n) {
    (void)dtype; (void)B; (void)P; (void)n;
    return (size_t)K * 1024 * sizeof(double) + (size_t)K * sizeof(int) + 16;      // per-CTA partial sums + the re-evaluation flags of dadmm_loss_from_sums
}

int dadmm_loss_fwd(int dtype, int K, int B, int P, int n, int64_t B_norm, const void* Y, const void* label, void* losses,
                   void* ws, size_t ws_bytes, dadmm_stream_t stream) {
    if (K <= 0 || B <= 0 || P <= 0 || n <= 0 || B_norm <= 0) DADMM_FAIL(-1, "loss_fwd: bad dims");
    if (!Y || !label || !losses || !ws) DADMM_FAIL(-1, "loss_fwd: null pointer");
    if (ws_bytes < dadmm_loss_ws_bytes(dtype, K, B, P, n)) DADMM_FAIL(-1, "loss_fwd: workspace too small");
    const long long rows = (long long)B * P;
    const int nblk = (int)std::min<long long>(1024, ceil_div64(rows, 8));
    const double inv = 1.0 / ((double)P * (double)B_norm * (double)n);
    cudaStream_t s = (cudaStream_t)stream;
    ProfScope prof(PROF_LOSS, s);
    const bool v4 = (n % 4 == 0) && aligned_to(Y, 16) && aligned_to(label, 16);
    if (dtype == DADMM_F32) {
        if (v4) loss_partial_kernel<float, 4><<<dim3(nblk, K), 256, 0, s>>>((const float*)Y, (const float*)label, B, P, n, (double*)ws);
        else loss_partial_kernel<float, 1><<<dim3(nblk, K), 256, 0, s>>>((const float*)Y, (const float*)label, B, P, n, (double*)ws);
        DADMM_LAUNCHED();
        loss_final_kernel<float><<<K, 256, 0, s>>>((const double*)ws, nblk, K, inv, (float*)losses);
        DADMM_LAUNCHED();
    } else if (dtype == DADMM_F64) {
        loss_partial_kernel<double, 1><<<dim3(nblk, K), 256, 0, s>>>((const double*)Y, (const double*)label, B, P, n, (double*)ws);
        DADMM_LAUNCHED();
        loss_final_kernel<double><<<K, 256, 0, s>>>((const double*)ws, nblk, K, inv, (double*)losses);
        DADMM_LAUNCHED();
    } else {
        DADMM_FAIL(-1, "loss_fwd: unknown dtype %d", dtype);
    }
    return 0;
}

int dadmm_loss_from_sums(int dtype, int k0, int k1, int B, int P, int n, int64_t B_norm, const void* agent_sum,
                         const double* sumsq, const void* label, const void* Y, void* losses, void* ws, size_t ws_bytes,
                         dadmm_stream_t stream) {
    if (k0 < 0 || k1 <= k0 || B <= 0 || P <= 0 || n <= 0 || B_norm <= 0) DADMM_FAIL(-1, "loss_from_sums: bad dims");
    if (!agent_sum || !sumsq || !label || !losses || !ws) DADMM_FAIL(-1, "loss_from_sums: null pointer");
    if (ws_bytes < dadmm_loss_ws_bytes(dtype, k1, B, P, n)) DADMM_FAIL(-1, "loss_from_sums: workspace too small");
    const long long tot = (long long)B * n;
    const int nblk = (int)std::min<long long>(512, ceil_div64(tot, 2048));
    const int Kn = k1 - k0;
    const double inv = 1.0 / ((double)P * (double)B_norm * (double)n);
    const size_t es = dtype == DADMM_F64 ? 8 : 4;
    cudaStream_t s = (cudaStream_t)stream;
    ProfScope prof(PROF_LOSS, s);
    double* part = (double*)ws;
    int* need = Y ? (int*)((char*)ws + (size_t)k1 * 1024 * sizeof(double)) : nullptr;     // tail of the workspace (dadmm_loss_ws_bytes)
    // where the sums cancel (need[k]), the iteration is evaluated again exactly from Y: same kernels as dadmm_loss_fwd, gated
    // on the device flag -- no host round trip; an unflagged iteration costs its CTAs one load
    const long long rows = (long long)B * P;
    const int nblk_x = (int)std::min<long long>(1024, ceil_div64(rows, 8));
    const bool v4 = (n % 4 == 0) && aligned_to(Y, 16) && aligned_to(label, 16);
    if (dtype == DADMM_F32) {
        loss_sums_partial_kernel<float><<<dim3(nblk, Kn), 256, 0, s>>>((const float*)agent_sum + (size_t)k0 * tot, (const float*)label, tot, part);
        DADMM_LAUNCHED();
        loss_sums_final_kernel<float><<<Kn, 256, 0, s>>>(part, nblk, sumsq + k0, P, inv, (float*)((char*)losses + (size_t)k0 * es), need);
        DADMM_LAUNCHED();
        if (Y) {
            const float* Yk = (const float*)Y + (size_t)k0 * rows * n;
            if (v4) loss_partial_kernel<float, 4><<<dim3(nblk_x, Kn), 256, 0, s>>>(Yk, (const float*)label, B, P, n, part, need);
            else loss_partial_kernel<float, 1><<<dim3(nblk_x, Kn), 256, 0, s>>>(Yk, (const float*)label, B, P, n, part, need);
            DADMM_LAUNCHED();
            loss_final_kernel<float><<<Kn, 256, 0, s>>>(part, nblk_x, Kn, inv, (float*)((char*)losses + (size_t)k0 * es), need);
            DADMM_LAUNCHED();
        }
    } else if (dtype == DADMM_F64) {
        loss_sums_partial_kernel<double><<<dim3(nblk, Kn), 256, 0, s>>>((const double*)agent_sum + (size_t)k0 * tot, (const double*)label, tot, part);
        DADMM_LAUNCHED();
        loss_sums_final_kernel<double><<<Kn, 256, 0, s>>>(part, nblk, sumsq + k0, P, inv, (double*)((char*)losses + (size_t)k0 * es), need);
        DADMM_LAUNCHED();
        if (Y) {
            loss_partial_kernel<double, 1><<<dim3(nblk_x, Kn), 256, 0, s>>>((const double*)Y + (size_t)k0 * rows * n, (const double*)label, B, P, n, part, need);
            DADMM_LAUNCHED();
            loss_final_kernel<double><<<Kn, 256, 0, s>>>(part, nblk_x, Kn, inv, (double*)((char*)losses + (size_t)k0 * es), need);
            DADMM_LAUNCHED();
        }
    } else {
        DADMM_FAIL(-1, "loss_from_sums: unknown dtype %d", dtype);
    }
    return 0;
}

int dadmm_loss_bwd(int dtype, int K, int B, int P, int n, const void* Y, const void* label, const double* coef, void* gY,
                   dadmm_stream_t stream) {
    if (K <= 0 || B <= 0 || P <= 0 || n <= 0) DADMM_FAIL(-1, "loss_bwd: bad dims");
    if (!Y || !label || !coef || !gY) DADMM_FAIL(-1, "loss_bwd: null pointer");
    const size_t es = dtype == DADMM_F64 ? 8 : 4;
    const long long per_k = (long long)B * P * n;
    const int nblk = (int)std::min<long long>(148 * 16, ceil_div64(per_k, 256));
    cudaStream_t s = (cudaStream_t)stream;
    for (int k = 0; k < K; ++k) {
        char* g = (char*)gY + (size_t)k * per_k * es;
        const char* y = (const char*)Y + (size_t)k * per_k * es;
        if (coef[k] == 0.0) {
            DADMM_CUDA(cudaMemsetAsync(g, 0, (size_t)per_k * es, s));
        } else if (dtype == DADMM_F32) {
            loss_bwd_kernel<float><<<nblk, 256, 0, s>>>((const float*)y, (const float*)label, B, P, n, (float)coef[k], (float*)g);
            DADMM_LAUNCHED();
        } else if (dtype == DADMM_F64) {
            loss_bwd_kernel<double><<<nblk, 256, 0, s>>>((const double*)y, (const double*)label, B, P, n, coef[k], (double*)g);
            DADMM_LAUNCHED();
        } else {
            DADMM_FAIL(-1, "loss_bwd: unknown dtype %d", dtype);
        }
    }
    return 0;
}

static int gcn_check(int B, int P, int C) {
    if (B <= 0 || P <= 0 || C <= 0) DADMM_FAIL(-1, "gcn_epilogue: bad dims");
    if (P > gcn::kMaxP) DADMM_FAIL(-2, "gcn_epilogue: P=%d above the %d agents one tile holds", P, gcn::kMaxP);
    return 0;
}

int dadmm_gcn_partial_rows(int B, int C) { return B > 0 && C > 0 ? ceil_div(B, gcn::problems_per_cta(B, C)) : 0; }

int dadmm_gcn_epilogue_fwd(int B, int P, int C, const void* H, const void* adj, const void* bias, const void* bn_w,
                           const void* bn_b, const void* run_mean, const void* run_var, int training, double eps,
                           double slope, const void* mask, void* out, void* act, void* mean, void* var,
                           dadmm_stream_t stream) {
    if (int e = gcn_check(B, P, C)) return e;
    if (!H || !adj || !bias || !bn_w || !bn_b || !out || !act) DADMM_FAIL(-1, "gcn_epilogue_fwd: null pointer");
    if (training ? (!mean || !var) : (!run_mean || !run_var)) DADMM_FAIL(-1, "gcn_epilogue_fwd: statistics pointers missing");
    gcn::Params p{};
    p.B = B; p.P = P; p.C = C; p.G = gcn::problems_per_cta(B, C);
    p.H = (const float*)H; p.adj = (const float*)adj; p.bias = (const float*)bias; p.bn_w = (const float*)bn_w;
    p.bn_b = (const float*)bn_b; p.run_mean = (const float*)run_mean; p.run_var = (const float*)run_var;
    p.mask = (const float*)mask; p.eps = (float)eps; p.slope = (float)slope; p.training = training ? 1 : 0;
    p.out = (float*)out; p.act = (float*)act; p.mean = (float*)mean; p.var = (float*)var;
    const size_t smem = gcn::smem_bytes(P);
    cudaStream_t s = (cudaStream_t)stream;
    if (int e = allow_smem(gcn::epilogue_fwd_kernel, smem)) return e;
    gcn::epilogue_fwd_kernel<<<dim3(ceil_div(C, gcn::kThreads), ceil_div(B, p.G)), gcn::kThreads, smem, s>>>(p);
    DADMM_LAUNCHED();
    return 0;
}

int dadmm_gcn_epilogue_bwd(int B, int P, int C, const void* gout, const void* adj, const void* bn_w, const void* run_mean,
                           const void* run_var, int training, double eps, double slope, const void* mask,
                           const void* act, const void* mean, const void* var, void* gH, void* partials,
                           dadmm_stream_t stream) {
    if (int e = gcn_check(B, P, C)) return e;
    if (!gout || !adj || !bn_w || !act || !gH || !partials) DADMM_FAIL(-1, "gcn_epilogue_bwd: null pointer");
    if (training ? (!mean || !var) : (!run_mean || !run_var)) DADMM_FAIL(-1, "gcn_epilogue_bwd: statistics pointers missing");
    gcn::Params p{};
    p.B = B; p.P = P; p.C = C; p.G = gcn::problems_per_cta(B, C);
    p.gout = (const float*)gout; p.adj = (const float*)adj; p.bn_w = (const float*)bn_w;
    p.run_mean = (const float*)run_mean; p.run_var = (const float*)run_var; p.mask = (const float*)mask;
    p.eps = (float)eps; p.slope = (float)slope; p.training = training ? 1 : 0;
    p.act = (float*)act; p.mean = (float*)mean; p.var = (float*)var;
    p.gH = (float*)gH; p.partials = (float*)partials;
    const size_t smem = gcn::smem_bytes(P);
    cudaStream_t s = (cudaStream_t)stream;
    if (int e = allow_smem(gcn::epilogue_bwd_kernel, smem)) return e;
    gcn::epilogue_bwd_kernel<<<dim3(ceil_div(C, gcn::kThreads), ceil_div(B, p.G)), gcn::kThreads, smem, s>>>(p);
    DADMM_LAUNCHED();
    return 0;
}

}  // extern "C"

// common.cuh -- shared device helpers for libdadmm_sm100 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <utility>

namespace dadmm {

extern thread_local char g_err[512];
extern std::atomic<long long> g_launches;

#define DADMM_FAIL(code, ...)                                   \
    do {                                                        \
        snprintf(::dadmm::g_err, sizeof(::dadmm::g_err), __VA_ARGS__); \
        return (code);                                          \
    } while (0)

#define DADMM_CUDA(expr)                                                                       \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            snprintf(::dadmm::g_err, sizeof(::dadmm::g_err), "%s:%d %s -> %s", __FILE__, __LINE__, #expr, \
                     cudaGetErrorString(_e));                                                  \
            return (int)_e;                                                                    \
        }                                                                                      \
    } while (0)

#define DADMM_LAUNCHED()                         \
    do {                                         \
        ::dadmm::g_launches.fetch_add(1);        \
        DADMM_CUDA(cudaGetLastError());          \
    } while (0)

// Optional per-kernel-kind timing (dadmm_profile_enable / dadmm_profile_read): when enabled, every
// launch is bracketed by a pair of CUDA events recorded on the launching stream.
enum ProfKind { PROF_CONTRACT_SIMT = 0, PROF_CONTRACT_TC = 1, PROF_STEP_FWD = 2, PROF_STEP_BWD = 3,
                PROF_REDUCE_HYP = 4, PROF_LOSS = 5, PROF_SPLIT = 6, PROF_CONTRACT_STAGE1 = 7, PROF_KINDS = 8 };
void prof_begin(int kind, cudaStream_t s);
void prof_end(cudaStream_t s);
struct ProfScope {
    cudaStream_t s;
    ProfScope(int kind, cudaStream_t stream) : s(stream) { prof_begin(kind, s); }
    ~ProfScope() { prof_end(s); }
};

// ---------------------------------------------------------------------------------------------
// Arithmetic that mirrors the reference's eager PyTorch ops: one IEEE rounding per op, never
// contracted into an FMA (unfolded_DLASSO.py:73-99 evaluates `a*b` and `+` as separate kernels).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub_rn(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double sub_rn(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }

// torch.clamp(x, -c, c): NaN propagates (both comparisons false).
template <typename T>
__device__ __forceinline__ T clamp_sym(T x, T c) {
    return x < -c ? -c : (x > c ? c : x);
}
// torch.clamp backward mask: closed interval, false for NaN.
template <typename T>
__device__ __forceinline__ bool in_closed(T x, T c) {
    return (x >= -c) && (x <= c);
}
// torch.sign: 0 for 0 and for NaN.
template <typename T>
__device__ __forceinline__ T sign_of(T x) {
    return (T)((x > (T)0) - (x < (T)0));
}
template <typename T>
__device__ __forceinline__ bool finite_val(T x) {
    return isfinite(x);
}

// ---------------------------------------------------------------------------------------------
// Small aligned vectors for 4/8/16-byte global and shared accesses.
// ---------------------------------------------------------------------------------------------
template <typename T, int N>
struct alignas(sizeof(T) * N) Vec {
    T v[N];
};

template <typename T, int N>
__device__ __forceinline__ Vec<T, N> ld_vec(const T* p) {
    return *reinterpret_cast<const Vec<T, N>*>(p);
}
// streaming (read-once) global load: bypass L1 allocation
template <typename T, int N>
__device__ __forceinline__ Vec<T, N> ld_stream(const T* p) {
    Vec<T, N> r;
    if constexpr (sizeof(T) * N == 16) {
        int4 t = __ldcs(reinterpret_cast<const int4*>(p));
        r = *reinterpret_cast<Vec<T, N>*>(&t);
    } else if constexpr (sizeof(T) * N == 8) {
        int2 t = __ldcs(reinterpret_cast<const int2*>(p));
        r = *reinterpret_cast<Vec<T, N>*>(&t);
    } else {
        static_assert(sizeof(T) * N == 4, "unsupported vector width");
        int t = __ldcs(reinterpret_cast<const int*>(p));
        r = *reinterpret_cast<Vec<T, N>*>(&t);
    }
    return r;
}
template <typename T, int N>
__device__ __forceinline__ void st_vec(T* p, const Vec<T, N>& x) {
    *reinterpret_cast<Vec<T, N>*>(p) = x;
}
template <typename T, int N>
__device__ __forceinline__ void st_stream(T* p, const Vec<T, N>& x) {
    if constexpr (sizeof(T) * N == 16) {
        __stcs(reinterpret_cast<int4*>(p), *reinterpret_cast<const int4*>(&x));
    } else if constexpr (sizeof(T) * N == 8) {
        __stcs(reinterpret_cast<int2*>(p), *reinterpret_cast<const int2*>(&x));
    } else {
        __stcs(reinterpret_cast<int*>(p), *reinterpret_cast<const int*>(&x));
    }
}

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---------------------------------------------------------------------------------------------
// Programmatic dependent launch.  The K-loop is a chain of ~1000 short dependent kernels per training step; with the
// attribute below the next kernel of the chain is scheduled while the current one drains, so its launch latency and
// its constant-input prologue (barrier / TMEM set-up, neighbour lists, per-agent scalars) overlap the predecessor's
// tail.  Rules every chain kernel follows: (1) `pdl_wait()` before the first access to anything an earlier kernel of
// the chain writes or still reads, executed by every thread; (2) `pdl_trigger()` only AFTER its own wait -- so when a
// kernel starts, everything older than its immediate predecessor is complete and visible, and only call-constant
// inputs (graph lists, the hyper-parameter table) may be touched before the wait.  Both instructions are no-ops in a
// kernel launched the classic way.  DADMM_PDL=0 / dadmm_set_pdl(0) turn the attribute off (A/B measurements, tests).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

extern std::atomic<int> g_pdl;           // -1: not read yet; dadmm_set_pdl() flips it at run time
inline bool pdl_enabled() {
    int v = g_pdl.load(std::memory_order_relaxed);
    if (v < 0) {
        const char* e = getenv("DADMM_PDL");
        v = (e && !strcmp(e, "0")) ? 0 : 1;
        g_pdl.store(v, std::memory_order_relaxed);
    }
    return v != 0;
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_chain(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

// Function attributes (dynamic shared-memory limits) are per device: `first()` is true once per device of the process
struct PerDeviceOnce {
    std::atomic<unsigned long long> seen{0};
    bool first() {
        int dev = 0;
        cudaGetDevice(&dev);
        const unsigned long long bit = 1ull << (dev & 63);
        return !(seen.fetch_or(bit) & bit);
    }
};

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline long long ceil_div64(long long a, long long b) { return (a + b - 1) / b; }

}  // namespace dadmm

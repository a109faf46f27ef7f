// step.cuh -- fused D-ADMM iteration kernels (forward, backward), hyper-parameter gradient
// reduction and the MSE loss.  HBM-bound streaming kernels.
//
// Work decomposition (forward and backward alike): a CTA owns a tile of TB problems x all P agents
// x CH = 32*VEC consecutive unknowns.  One warp processes one (problem, agent) row segment at a
// time -- CH contiguous elements, i.e. 128/256/512-byte fully coalesced vector accesses per tensor --
// and the tile's y values of ALL agents sit in shared memory, so the neighbour-consensus operator
// delta = 2*L*y (reference compute_delta, unfolded_DLASSO.py:127-140) is evaluated from shared
// memory in the reference's own accumulation order ("event lists", see dadmm.h) with no global
// gather and no atomics.
#pragma once
#include "common.cuh"

namespace dadmm {

constexpr int kStepThreads = 256;

template <typename T>
struct StepFwdParams {
    int B, P, n, TB;
    const int32_t *ev_ptr, *ev_idx, *deg, *gid;
    const T* hyp;
    long long hsb, hsp, hsc;
    T G, V, D, Uc;
    int hasD;
    const T *y, *U, *delta, *a, *atb;
    T *y_next, *U_next, *delta_next, *graw;
    int32_t* flags;
};

template <typename T>
struct StepBwdParams {
    int B, P, n, TB;
    const int32_t *ev_ptr, *ev_idx, *deg, *gid;
    const T* hyp;
    long long hsb, hsp, hsc;
    T G, V, D, Uc;
    int hasD;
    const T *y, *U, *delta, *graw, *y_next;
    const T *gy_a, *gy_b, *gU_next, *gd_next, *label;
    T loss_coef;
    T *gy, *ga, *gU, *gd, *partials;
};

// 2*L*x for node q of one problem, rows of the tile in shared memory (Sb = first row of the
// problem, CH elements per row).  Sequential accumulation of (x_q - x_e) over the event list:
// bit-identical to the reference's `delta[b,p] += diff; delta[b,j] -= diff` loops.
template <typename T, int VEC>
__device__ __forceinline__ Vec<T, VEC> laplace2(const T* Sb, int q, const int32_t* __restrict__ ev_idx,
                                                int e0, int e1, int lane) {
    constexpr int CH = 32 * VEC;
    const Vec<T, VEC> xq = ld_vec<T, VEC>(Sb + q * CH + lane * VEC);
    Vec<T, VEC> acc;
#pragma unroll
    for (int v = 0; v < VEC; ++v) acc.v[v] = (T)0;
    for (int base = e0; base < e1; base += 32) {
        const int mine = (base + lane < e1) ? __ldg(ev_idx + base + lane) : 0;
        const int cnt = min(32, e1 - base);
        for (int t = 0; t < cnt; ++t) {
            const int j = __shfl_sync(0xffffffffu, mine, t);
            const Vec<T, VEC> xj = ld_vec<T, VEC>(Sb + j * CH + lane * VEC);
#pragma unroll
            for (int v = 0; v < VEC; ++v) acc.v[v] = add_rn(acc.v[v], sub_rn(xq.v[v], xj.v[v]));
        }
    }
    return acc;
}

template <typename T, int VEC>
__device__ __forceinline__ Vec<T, VEC> vzero() {
    Vec<T, VEC> z;
#pragma unroll
    for (int v = 0; v < VEC; ++v) z.v[v] = (T)0;
    return z;
}

// ---------------------------------------------------------------------------------------------
// forward: reference unfolded_DLASSO.py:73-99 / gnn_dlasso_models_progressive.py:205-232
// ---------------------------------------------------------------------------------------------
template <typename T, int VEC>
__global__ void __launch_bounds__(kStepThreads) step_fwd_kernel(const StepFwdParams<T> p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int CH = 32 * VEC;
    const int P = p.P;
    const int R = p.TB * P;
    const bool recompute = (p.delta == nullptr);
    const bool consensus = (p.U_next != nullptr) || (p.delta_next != nullptr);
    T* S0 = reinterpret_cast<T*>(smem_raw);           // y_k tile (recompute only)
    T* S1 = S0 + (recompute ? (size_t)R * CH : 0);    // y_{k+1} tile
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const int nchunks = (p.n + CH - 1) / CH;
    const int chunk = blockIdx.x % nchunks;
    const int b0 = (blockIdx.x / nchunks) * p.TB;
    const int i = chunk * CH + lane * VEC;
    const bool act_i = i < p.n;

    if (recompute) {
        for (int r = warp; r < R; r += nwarps) {
            const int b = b0 + r / P, pp = r % P;
            Vec<T, VEC> v = vzero<T, VEC>();
            if (act_i && b < p.B) v = ld_vec<T, VEC>(p.y + ((size_t)b * P + pp) * p.n + i);
            st_vec<T, VEC>(S0 + (size_t)r * CH + lane * VEC, v);
        }
        __syncthreads();
    }

    unsigned bad = 0;
    for (int r = warp; r < R; r += nwarps) {
        const int bl = r / P, pp = r % P, b = b0 + bl;
        if (b >= p.B) {
            if (consensus) st_vec<T, VEC>(S1 + (size_t)r * CH + lane * VEC, vzero<T, VEC>());
            continue;
        }
        const size_t off = ((size_t)b * P + pp) * p.n + i;
        const int node = (p.gid ? __ldg(p.gid + b) : 0) * P + pp;
        const T* hp = p.hyp + b * p.hsb + pp * p.hsp;
        const T alpha = __ldg(hp), tau = __ldg(hp + p.hsc), rho = __ldg(hp + 2 * p.hsc);
        const T dg = (T)__ldg(p.deg + node);
        Vec<T, VEC> yv = vzero<T, VEC>(), Uv = yv, av = yv, bv = yv, dv = yv;
        if (recompute) {
            yv = ld_vec<T, VEC>(S0 + (size_t)r * CH + lane * VEC);
            dv = laplace2<T, VEC>(S0 + (size_t)bl * P * CH, pp, p.ev_idx, __ldg(p.ev_ptr + node),
                                  __ldg(p.ev_ptr + node + 1), lane);
            if (p.hasD) {
#pragma unroll
                for (int v = 0; v < VEC; ++v) dv.v[v] = clamp_sym(dv.v[v], p.D);
            }
        }
        if (act_i) {
            if (!recompute) {
                yv = ld_vec<T, VEC>(p.y + off);
                dv = ld_stream<T, VEC>(p.delta + off);
            }
            Uv = ld_vec<T, VEC>(p.U + off);
            av = ld_stream<T, VEC>(p.a + off);
            bv = ld_stream<T, VEC>(p.atb + off);
        }
        Vec<T, VEC> yn, rv;
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            const T y = yv.v[v], U = Uv.v[v];
            T rr = sub_rn(av.v[v], bv.v[v]);
            rr = add_rn(rr, mul_rn(sign_of(y), tau));
            rr = add_rn(rr, mul_rn(U, dg));
            rr = add_rn(rr, mul_rn(dv.v[v], rho));
            const T g = clamp_sym(rr, p.G);
            const T z = clamp_sym(sub_rn(y, mul_rn(alpha, g)), p.V);
            rv.v[v] = rr;
            yn.v[v] = z;
            if (!finite_val(y)) bad |= 1u;
            if (!finite_val(U)) bad |= 2u;
            if (g != g) bad |= 4u;
            if (!finite_val(z)) bad |= 8u;
        }
        if (act_i) {
            st_vec<T, VEC>(p.y_next + off, yn);
            if (p.graw) st_stream<T, VEC>(p.graw + off, rv);
        }
        if (consensus) st_vec<T, VEC>(S1 + (size_t)r * CH + lane * VEC, yn);
    }

    if (consensus) {
        __syncthreads();
        for (int r = warp; r < R; r += nwarps) {
            const int bl = r / P, pp = r % P, b = b0 + bl;
            if (b >= p.B) continue;
            const size_t off = ((size_t)b * P + pp) * p.n + i;
            const int node = (p.gid ? __ldg(p.gid + b) : 0) * P + pp;
            Vec<T, VEC> dn = laplace2<T, VEC>(S1 + (size_t)bl * P * CH, pp, p.ev_idx, __ldg(p.ev_ptr + node),
                                              __ldg(p.ev_ptr + node + 1), lane);
            if (p.hasD) {
#pragma unroll
                for (int v = 0; v < VEC; ++v) dn.v[v] = clamp_sym(dn.v[v], p.D);
            }
            if (act_i) {
                if (p.U_next) {
                    const T eta = __ldg(p.hyp + b * p.hsb + pp * p.hsp + 3 * p.hsc);
                    const Vec<T, VEC> Uv = ld_vec<T, VEC>(p.U + off);
                    Vec<T, VEC> Un;
#pragma unroll
                    for (int v = 0; v < VEC; ++v)
                        Un.v[v] = clamp_sym(add_rn(Uv.v[v], mul_rn(dn.v[v], eta)), p.Uc);
                    st_vec<T, VEC>(p.U_next + off, Un);
                }
                if (p.delta_next) st_vec<T, VEC>(p.delta_next + off, dn);
            }
        }
    }
    if (p.flags) {
        bad = __reduce_or_sync(0xffffffffu, bad);
        if (bad && lane == 0) atomicOr(p.flags, (int)bad);
    }
}

// ---------------------------------------------------------------------------------------------
// backward of one iteration (what autograd derives from the same reference lines)
// ---------------------------------------------------------------------------------------------
template <typename T, int VEC>
__global__ void __launch_bounds__(kStepThreads) step_bwd_kernel(const StepBwdParams<T> p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int CH = 32 * VEC;
    const int P = p.P;
    const int R = p.TB * P;
    const bool recompute = (p.delta == nullptr);
    T* S0 = reinterpret_cast<T*>(smem_raw);   // y_{k+1} tile, later y_k tile
    T* S1 = S0 + (size_t)R * CH;              // adjoint of the unclamped d_{k+1}
    T* Sp = S1 + (size_t)R * CH;              // per-row partial of d/d eta
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const int nchunks = (p.n + CH - 1) / CH;
    const int chunk = blockIdx.x % nchunks;
    const int b0 = (blockIdx.x / nchunks) * p.TB;
    const int i = chunk * CH + lane * VEC;
    const bool act_i = i < p.n;

    // phase A: y_{k+1} tile
    for (int r = warp; r < R; r += nwarps) {
        const int b = b0 + r / P, pp = r % P;
        Vec<T, VEC> v = vzero<T, VEC>();
        if (act_i && b < p.B) v = ld_vec<T, VEC>(p.y_next + ((size_t)b * P + pp) * p.n + i);
        st_vec<T, VEC>(S0 + (size_t)r * CH + lane * VEC, v);
    }
    __syncthreads();

    // phase B: dual update backward; adjoint of d_{k+1}; direct adjoint terms of y_{k+1}
    for (int r = warp; r < R; r += nwarps) {
        const int bl = r / P, pp = r % P, b = b0 + bl;
        if (b >= p.B) {
            st_vec<T, VEC>(S1 + (size_t)r * CH + lane * VEC, vzero<T, VEC>());
            continue;
        }
        const size_t off = ((size_t)b * P + pp) * p.n + i;
        const int node = (p.gid ? __ldg(p.gid + b) : 0) * P + pp;
        const T eta = __ldg(p.hyp + b * p.hsb + pp * p.hsp + 3 * p.hsc);
        const Vec<T, VEC> dn = laplace2<T, VEC>(S0 + (size_t)bl * P * CH, pp, p.ev_idx, __ldg(p.ev_ptr + node),
                                                __ldg(p.ev_ptr + node + 1), lane);
        Vec<T, VEC> Uv = vzero<T, VEC>(), gUn = Uv, gdn = Uv, dir = Uv, lab = Uv;
        if (act_i) {
            Uv = ld_vec<T, VEC>(p.U + off);
            if (p.gU_next) gUn = ld_vec<T, VEC>(p.gU_next + off);
            if (p.gd_next) gdn = ld_vec<T, VEC>(p.gd_next + off);
            if (p.gy_a) dir = ld_vec<T, VEC>(p.gy_a + off);
            if (p.gy_b) {
                const Vec<T, VEC> t = ld_stream<T, VEC>(p.gy_b + off);
#pragma unroll
                for (int v = 0; v < VEC; ++v) dir.v[v] += t.v[v];
            }
            if (p.label) lab = ld_vec<T, VEC>(p.label + (size_t)b * p.n + i);
        }
        const Vec<T, VEC> yn = ld_vec<T, VEC>(S0 + (size_t)r * CH + lane * VEC);
        Vec<T, VEC> um, tt;
        T peta = (T)0;
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            const T draw = dn.v[v];
            const bool mD = p.hasD ? in_closed(draw, p.D) : true;
            const T dc = p.hasD ? clamp_sym(draw, p.D) : draw;
            const T w = add_rn(Uv.v[v], mul_rn(dc, eta));
            const T u = in_closed(w, p.Uc) ? gUn.v[v] : (T)0;
            peta += u * dc;
            const T t = gdn.v[v] + eta * u;
            um.v[v] = u;
            tt.v[v] = (mD && act_i) ? t : (T)0;
            if (p.label) dir.v[v] += p.loss_coef * (yn.v[v] - lab.v[v]);
        }
        st_vec<T, VEC>(S1 + (size_t)r * CH + lane * VEC, tt);
        if (act_i) {
            st_vec<T, VEC>(p.gU + off, um);    // temporaries, re-read by the same thread in phase D
            st_vec<T, VEC>(p.gy + off, dir);
        }
        peta = warp_sum(peta);
        if (lane == 0) Sp[r] = peta;
    }
    __syncthreads();

    // phase C: y_k tile (needed only to recompute delta_k = clampD(2L y_k))
    if (recompute) {
        for (int r = warp; r < R; r += nwarps) {
            const int b = b0 + r / P, pp = r % P;
            Vec<T, VEC> v = vzero<T, VEC>();
            if (act_i && b < p.B) v = ld_vec<T, VEC>(p.y + ((size_t)b * P + pp) * p.n + i);
            st_vec<T, VEC>(S0 + (size_t)r * CH + lane * VEC, v);
        }
        __syncthreads();
    }

    // phase D: primal update backward
    for (int r = warp; r < R; r += nwarps) {
        const int bl = r / P, pp = r % P, b = b0 + bl;
        if (b >= p.B) continue;
        const size_t off = ((size_t)b * P + pp) * p.n + i;
        const int node = (p.gid ? __ldg(p.gid + b) : 0) * P + pp;
        const int e0 = __ldg(p.ev_ptr + node), e1 = __ldg(p.ev_ptr + node + 1);
        const T* hp = p.hyp + b * p.hsb + pp * p.hsp;
        const T alpha = __ldg(hp), rho = __ldg(hp + 2 * p.hsc);
        const T dg = (T)__ldg(p.deg + node);
        const Vec<T, VEC> lt = laplace2<T, VEC>(S1 + (size_t)bl * P * CH, pp, p.ev_idx, e0, e1, lane);
        Vec<T, VEC> yv = vzero<T, VEC>(), dv = yv, rv = yv, um = yv, ytot = yv;
        if (recompute) {
            yv = ld_vec<T, VEC>(S0 + (size_t)r * CH + lane * VEC);
            dv = laplace2<T, VEC>(S0 + (size_t)bl * P * CH, pp, p.ev_idx, e0, e1, lane);
            if (p.hasD) {
#pragma unroll
                for (int v = 0; v < VEC; ++v) dv.v[v] = clamp_sym(dv.v[v], p.D);
            }
        }
        if (act_i) {
            if (!recompute) {
                yv = ld_vec<T, VEC>(p.y + off);
                dv = ld_stream<T, VEC>(p.delta + off);
            }
            rv = ld_stream<T, VEC>(p.graw + off);
            um = ld_vec<T, VEC>(p.gU + off);
            ytot = ld_vec<T, VEC>(p.gy + off);
        }
        Vec<T, VEC> o_gy, o_ga, o_gU, o_gd;
        T pa = (T)0, pt = (T)0, pr = (T)0;
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            const T y = yv.v[v], rr = rv.v[v];
            const T g = clamp_sym(rr, p.G);
            const T z = sub_rn(y, mul_rn(alpha, g));
            const T zb = in_closed(z, p.V) ? (ytot.v[v] + lt.v[v]) : (T)0;
            pa -= zb * g;
            const T rb = in_closed(rr, p.G) ? (-alpha * zb) : (T)0;
            pt += rb * sign_of(y);
            pr += rb * dv.v[v];
            o_gy.v[v] = zb;
            o_ga.v[v] = rb;
            o_gU.v[v] = um.v[v] + dg * rb;
            o_gd.v[v] = rho * rb;
        }
        if (act_i) {
            st_vec<T, VEC>(p.gy + off, o_gy);
            st_vec<T, VEC>(p.ga + off, o_ga);
            st_vec<T, VEC>(p.gU + off, o_gU);
            if (p.gd) st_vec<T, VEC>(p.gd + off, o_gd);
        } else {
            pa = pt = pr = (T)0;
        }
        pa = warp_sum(pa);
        pt = warp_sum(pt);
        pr = warp_sum(pr);
        if (lane == 0) {
            T* q = p.partials + (((size_t)chunk * p.B + b) * P + pp) * 4;
            q[0] = pa;
            q[1] = pt;
            q[2] = pr;
            q[3] = Sp[r];
        }
    }
}

// ---------------------------------------------------------------------------------------------
// finish d/d(alpha,tau,rho,eta): sum the per-tile partials [nchunks][B][P][4] in fp64.
//   per_sample == 0: one CTA per agent, sums over tiles and problems   (model #1: hyp [P,4])
//   per_sample == 1: one thread per (problem, agent), sums over tiles  (model #3: hyp [B,4,P])
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) reduce_hyp_table_kernel(const T* __restrict__ part, int nchunks, int B, int P,
                                                               T* out, long long sp, long long sc, int accumulate) {
    const int pp = blockIdx.x;
    double acc[4] = {0, 0, 0, 0};
    const long long rows = (long long)nchunks * B;
    for (long long rI = threadIdx.x; rI < rows; rI += blockDim.x) {
        const T* q = part + (rI * P + pp) * 4;
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[c] += (double)q[c];
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
        T* o = out + pp * sp + threadIdx.x * sc;
        *o = accumulate ? (T)((double)*o + s) : (T)s;
    }
}

template <typename T>
__global__ void __launch_bounds__(256) reduce_hyp_sample_kernel(const T* __restrict__ part, int nchunks, int B, int P,
                                                                T* out, long long sb, long long sp, long long sc,
                                                                int accumulate) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)B * P) return;
    const int b = (int)(idx / P), pp = (int)(idx % P);
    double acc[4] = {0, 0, 0, 0};
    for (int ch = 0; ch < nchunks; ++ch) {
        const T* q = part + (((long long)ch * B + b) * P + pp) * 4;
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[c] += (double)q[c];
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        T* o = out + b * sb + pp * sp + c * sc;
        *o = accumulate ? (T)((double)*o + acc[c]) : (T)acc[c];
    }
}

// ---------------------------------------------------------------------------------------------
// MSE loss per iteration (gnn_dlasso_utils.py:54-66) and its gradient.
// ---------------------------------------------------------------------------------------------
// grid = (nblk, K); one warp per (problem, agent) row, VEC-wide coalesced loads; fp32 within a row segment,
// fp64 across rows.
// `need` (optional, [gridDim.y]): launch-time-unknown selection of the iterations to evaluate -- a CTA whose iteration is
// not flagged returns at once (dadmm_loss_from_sums re-evaluates from Y only where the sums cancelled)
template <typename T, int VEC>
__global__ void __launch_bounds__(256) loss_partial_kernel(const T* __restrict__ Y, const T* __restrict__ label,
                                                           int B, int P, int n, double* __restrict__ partial,
                                                           const int* __restrict__ need = nullptr) {
    const int k = blockIdx.y;
    if (need && !need[k]) return;
    const long long rows = (long long)B * P;
    const T* Yk = Y + (long long)k * rows * n;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    double acc = 0;
    for (long long r = (long long)blockIdx.x * nw + warp; r < rows; r += (long long)gridDim.x * nw) {
        const T* yr = Yk + r * n;
        const T* lr = label + (r / P) * n;
        T s = (T)0;
        for (int i = lane * VEC; i < n; i += 32 * VEC) {
            const Vec<T, VEC> a = ld_stream<T, VEC>(yr + i);
            const Vec<T, VEC> l = ld_vec<T, VEC>(lr + i);
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                const T d = a.v[v] - l.v[v];
                s += d * d;
            }
        }
        acc += (double)s;
    }
    __shared__ double sh[8];
    acc = warp_sum(acc);
    if (lane == 0) sh[warp] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0;
        for (int w = 0; w < nw; ++w) t += sh[w];
        partial[(long long)k * gridDim.x + blockIdx.x] = t;
    }
}

template <typename T>
__global__ void loss_final_kernel(const double* __restrict__ partial, int nblk, int K, double inv_norm, T* losses,
                                  const int* __restrict__ need = nullptr) {
    const int k = blockIdx.x;
    if (need && !need[k]) return;
    double acc = 0;
    for (int j = threadIdx.x; j < nblk; j += blockDim.x) acc += partial[(long long)k * nblk + j];
    __shared__ double sh[8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    acc = warp_sum(acc);
    if (lane == 0) sh[warp] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += sh[w];
        losses[k] = (T)(s * inv_norm);
    }
}

template <typename T>
__global__ void __launch_bounds__(256) loss_bwd_kernel(const T* __restrict__ Yk, const T* __restrict__ label, int B,
                                                       int P, int n, T coef, T* __restrict__ gYk) {
    const long long per_k = (long long)B * P * n;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < per_k; e += (long long)gridDim.x * blockDim.x) {
        const long long row = e / n;
        const int ii = (int)(e - row * n);
        const long long b = row / P;
        gYk[e] = coef * (Yk[e] - label[b * n + ii]);
    }
}

}  // namespace dadmm

// gcn.cuh -- per-problem epilogue of one graph-convolution layer of the model-#3 hypernetwork, forward and backward.
//
// Reference (gnn_dlasso_models_progressive.py:37-72): for every sample b of the batch, in a Python loop,
//     x = leaky_relu(GCNConv_i(x, edge_index));  x = BatchNorm1d_i(x);  x = dropout(x)        (x: [P, C], ONE graph)
// with GCNConv = A_hat (x W^T) + bias, A_hat = D^-1/2 (Adj + I) D^-1/2 of that sample's graph, and BatchNorm statistics
// taken over the P nodes of that one graph.  Everything after the dense product H = x W^T is local to one problem and
// one channel: a P x P mix, an activation, a P-sample normalisation, a mask.  One thread owns one (problem, channel)
// column of P values; a CTA owns 128 channels of a run of problems; H is read once, the layer's output written once
// (the PyTorch composition this replaces was ~12 launches and ~10 round trips of a [B*P, C] tensor per layer).
//
//   forward :  Z = A_hat_b H_b + bias;  A = leaky(Z);  train: (mean_b, var_b) over the P rows, xhat = (A - mean) rstd
//              eval : xhat = (A - running_mean) rsqrt(running_var + eps);   out = (xhat gamma + beta) * mask
//   backward:  g = gout * mask;  d gamma += <g, xhat>, d beta += sum g;  train: gA = (g gamma - m1 - xhat m2) rstd with
//              m1 = mean_rows(g gamma), m2 = mean_rows(g gamma xhat);  eval: gA = g gamma rstd_run
//              gZ = gA * (A > 0 ? 1 : slope);  d bias += sum_rows gZ;  gH = A_hat_b^T gZ
// Parameter-gradient sums leave as one partial row per CTA ([nblk][3][C]; summed by the caller): deterministic, no atomics.
#pragma once
#include "common.cuh"

namespace dadmm {
namespace gcn {

constexpr int kThreads = 128;     // channels per CTA
constexpr int kMaxP = 64;

struct Params {
    int B, P, C, G;                       // G: problems per CTA (consecutive)
    const float *H, *adj, *bias, *bn_w, *bn_b, *run_mean, *run_var, *mask;
    float eps, slope;
    int training;
    float *out, *act, *mean, *var;        // forward outputs (act: post-activation, saved for backward; mean / var [B,C]: training)
    const float* gout;                    // backward
    float *gH, *partials;
};

// dynamic shared memory: adjacency [P][P] | two column tiles [P][kThreads]
inline size_t smem_bytes(int P) { return ((size_t)P * P + 2 * (size_t)P * kThreads) * sizeof(float); }

__global__ void __launch_bounds__(kThreads) epilogue_fwd_kernel(const Params p) {
    extern __shared__ __align__(16) float sm[];
    const int P = p.P, C = p.C;
    float* sAdj = sm;                       // [P][P]
    float* sH = sm + P * P;                 // [P][kThreads]: this thread's column of H ...
    float* sA = sH + P * kThreads;          // ... and of A = leaky(A_hat H + bias)   (columns are thread-private)
    const int t = threadIdx.x, c = blockIdx.x * kThreads + t;
    const bool act_c = c < C;
    const float bias = act_c ? __ldg(p.bias + c) : 0.f, gamma = act_c ? __ldg(p.bn_w + c) : 0.f, beta = act_c ? __ldg(p.bn_b + c) : 0.f;
    float e_mean = 0.f, e_rstd = 0.f;
    if (!p.training && act_c) {
        e_mean = __ldg(p.run_mean + c);
        e_rstd = rsqrtf(__ldg(p.run_var + c) + p.eps);
    }
    const float invP = 1.f / (float)P;
    const int b_end = min(p.B, (blockIdx.y + 1) * p.G);
    for (int b = blockIdx.y * p.G; b < b_end; ++b) {
        __syncthreads();                                            // the previous problem's adjacency is done with
        for (int q = t; q < P * P; q += kThreads) sAdj[q] = __ldg(p.adj + (size_t)b * P * P + q);
        const size_t row0 = (size_t)b * P;
        if (act_c)
            for (int j = 0; j < P; ++j) sH[j * kThreads + t] = __ldg(p.H + (row0 + j) * C + c);
        __syncthreads();
        if (!act_c) continue;
        float s1 = 0.f;
        for (int i = 0; i < P; ++i) {
            float z = bias;
            for (int j = 0; j < P; ++j) z = fmaf(sAdj[i * P + j], sH[j * kThreads + t], z);
            const float a = z > 0.f ? z : z * p.slope;
            sA[i * kThreads + t] = a;
            p.act[(row0 + i) * C + c] = a;
            s1 += a;
        }
        float mean = e_mean, rstd = e_rstd;
        if (p.training) {
            mean = s1 * invP;
            float s2 = 0.f;
            for (int i = 0; i < P; ++i) {
                const float d = sA[i * kThreads + t] - mean;
                s2 = fmaf(d, d, s2);
            }
            const float var = s2 * invP;                           // biased, as BatchNorm normalises with
            rstd = rsqrtf(var + p.eps);
            p.mean[(size_t)b * C + c] = mean;
            p.var[(size_t)b * C + c] = var;
        }
        for (int i = 0; i < P; ++i) {
            const size_t o = (row0 + i) * C + c;
            float v = fmaf((sA[i * kThreads + t] - mean) * rstd, gamma, beta);
            if (p.mask) v *= __ldg(p.mask + o);
            p.out[o] = v;
        }
    }
}

__global__ void __launch_bounds__(kThreads) epilogue_bwd_kernel(const Params p) {
    extern __shared__ __align__(16) float sm[];
    const int P = p.P, C = p.C;
    float* sAdj = sm;                       // [P][P]
    float* sCol = sm + P * P;               // [P][kThreads]: gZ of this thread's column
    const int t = threadIdx.x, c = blockIdx.x * kThreads + t;
    const bool act_c = c < C;
    const float gamma = act_c ? __ldg(p.bn_w + c) : 0.f;
    float e_mean = 0.f, e_rstd = 0.f;
    if (!p.training && act_c) {
        e_mean = __ldg(p.run_mean + c);
        e_rstd = rsqrtf(__ldg(p.run_var + c) + p.eps);
    }
    const float invP = 1.f / (float)P;
    float d_gamma = 0.f, d_beta = 0.f, d_bias = 0.f;
    const int b_end = min(p.B, (blockIdx.y + 1) * p.G);
    for (int b = blockIdx.y * p.G; b < b_end; ++b) {
        __syncthreads();
        for (int q = t; q < P * P; q += kThreads) sAdj[q] = __ldg(p.adj + (size_t)b * P * P + q);
        const size_t row0 = (size_t)b * P;
        if (act_c) {
            float mean = e_mean, rstd = e_rstd;
            if (p.training) {
                mean = __ldg(p.mean + (size_t)b * C + c);
                rstd = rsqrtf(__ldg(p.var + (size_t)b * C + c) + p.eps);
            }
            // pass 1: sums over the problem's rows
            float m1 = 0.f, m2 = 0.f;
            for (int i = 0; i < P; ++i) {
                const size_t o = (row0 + i) * C + c;
                float g = __ldg(p.gout + o);
                if (p.mask) g *= __ldg(p.mask + o);
                const float xhat = (__ldg(p.act + o) - mean) * rstd;
                d_beta += g;
                d_gamma = fmaf(g, xhat, d_gamma);
                m1 = fmaf(g, gamma, m1);
                m2 = fmaf(g * gamma, xhat, m2);
            }
            m1 *= invP;
            m2 *= invP;
            // pass 2: gZ per row (the loads hit L1: this thread read the same addresses a moment ago)
            for (int i = 0; i < P; ++i) {
                const size_t o = (row0 + i) * C + c;
                float g = __ldg(p.gout + o);
                if (p.mask) g *= __ldg(p.mask + o);
                const float a = __ldg(p.act + o);
                const float gx = g * gamma;
                float ga = p.training ? (gx - m1 - (a - mean) * rstd * m2) * rstd : gx * rstd;
                const float gz = a > 0.f ? ga : ga * p.slope;
                d_bias += gz;
                sCol[i * kThreads + t] = gz;
            }
        }
        __syncthreads();                                            // adjacency staged (columns are thread-private)
        if (act_c)
            for (int j = 0; j < P; ++j) {
                float acc = 0.f;
                for (int i = 0; i < P; ++i) acc = fmaf(sAdj[i * P + j], sCol[i * kThreads + t], acc);
                p.gH[(row0 + j) * C + c] = acc;
            }
    }
    if (act_c) {
        float* row = p.partials + (size_t)blockIdx.y * 3 * C;
        row[c] = d_gamma;
        row[C + c] = d_beta;
        row[2 * C + c] = d_bias;
    }
}

// problems per CTA: enough CTAs for ~8 per SM, at most 64 problems each
inline int problems_per_cta(int B, int C) {
    const int chunks = ceil_div(C, kThreads);
    const int want = 148 * 8;
    int G = std::max(1, (B * chunks) / want);
    return std::min(G, 64);
}

}  // namespace gcn
}  // namespace dadmm

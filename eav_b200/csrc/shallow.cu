// ShallowConvNet of the reference's Transformer_torch/Transformer_EEG.py:107-148 on sm_100a:
// Conv2d(1,40,(1,13)) -> 40 per-filter spatial Linear(30,1) -> 12 single-head transformer layers (T = 488, d = 40)
// -> BatchNorm2d(40) -> square -> AvgPool((1,35), stride 7) -> log(clamp) -> dropout -> Linear(2600, nb_classes) ->
// softmax, forward and backward.  fp32 throughout (the parity gate is 1e-4 against the unmodified reference).
//
// The model is small (0.23 M parameters, 0.2 GFLOP per sample) and every contraction is a skinny matrix product, so one
// register-tiled fp32 GEMM kernel with strides / transposes / batching carries all of them (linear layers, Q K^T, P V and
// every weight gradient), next to fused row kernels for softmax, LayerNorm + dropout + residual and the BN/pool/log head.
// All reductions have a fixed order (deterministic).  Activations saved for backward live in the caller's workspace.
#include <math.h>

#include "eav_common.cuh"

#define TRY_RC(call) do { int rc__ = (call); if (rc__) return rc__; } while (0)

namespace eav {
namespace {

// ------------------------------------------------------------------------------------------------------------------
// C[b] (M x N) = alpha * op(A[b]) (M x K) * op(B[b]) (K x N) [+ bias[n]] [relu] [+ beta * C[b]]
// op(A)[m][k] = ta ? A[k*lda + m] : A[m*lda + k];  op(B)[k][n] = tb ? B[n*ldb + k] : B[k*ldb + n].
// 64x64 tile per CTA, 16 k per step, 4x4 outputs per thread (256 threads).
// ------------------------------------------------------------------------------------------------------------------
constexpr int GT = 64, GK = 16;
__global__ void __launch_bounds__(256)
gemm_kernel(int M, int N, int K, float alpha, const float *__restrict__ A, int lda, int64_t sa, int ta,
            const float *__restrict__ B, int ldb, int64_t sb, int tb, float beta, float *__restrict__ C, int ldc,
            int64_t sc, const float *__restrict__ bias, int relu) {
    __shared__ float As[GK][GT + 1], Bs[GK][GT + 1];
    const int bz = blockIdx.z;
    A += (int64_t)bz * sa; B += (int64_t)bz * sb; C += (int64_t)bz * sc;
    const int m0 = blockIdx.y * GT, n0 = blockIdx.x * GT, tid = threadIdx.x;
    const int tm = (tid >> 4) * 4, tn = (tid & 15) * 4;
    float acc[4][4] = {};
    for (int k0 = 0; k0 < K; k0 += GK) {
        for (int i = tid; i < GK * GT; i += 256) {
            int kk, mm;
            if (ta) { mm = i % GT; kk = i / GT; } else { kk = i % GK; mm = i / GK; }
            const int m = m0 + mm, k = k0 + kk;
            As[kk][mm] = (m < M && k < K) ? (ta ? A[(int64_t)k * lda + m] : A[(int64_t)m * lda + k]) : 0.f;
        }
        for (int i = tid; i < GK * GT; i += 256) {
            int kk, nn;
            if (tb) { kk = i % GK; nn = i / GK; } else { nn = i % GT; kk = i / GT; }
            const int n = n0 + nn, k = k0 + kk;
            Bs[kk][nn] = (n < N && k < K) ? (tb ? B[(int64_t)n * ldb + k] : B[(int64_t)k * ldb + n]) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < GK; ++kk) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { a[i] = As[kk][tm + i]; b[i] = Bs[kk][tn + i]; }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int m = m0 + tm + i, n = n0 + tn + j;
            if (m < M && n < N) {
                float v = alpha * acc[i][j];
                if (bias) v += bias[n];
                if (relu) v = v > 0.f ? v : 0.f;
                float *c = C + (int64_t)m * ldc + n;
                *c = beta != 0.f ? v + beta * *c : v;
            }
        }
}

struct Gemm {
    cudaStream_t st;
    int run(int M, int N, int K, float alpha, const float *A, int lda, int ta, const float *B, int ldb, int tb, float beta,
            float *C, int ldc, const float *bias = nullptr, int relu = 0, int batch = 1, int64_t sa = 0, int64_t sb = 0,
            int64_t sc = 0) const {
        dim3 grid(cdiv(N, GT), cdiv(M, GT), batch);
        gemm_kernel<<<grid, 256, 0, st>>>(M, N, K, alpha, A, lda, sa, ta, B, ldb, sb, tb, beta, C, ldc, sc, bias, relu);
        EAV_CUDA_LAUNCH_CHECK("shallow_gemm");
        return 0;
    }
};

// ------------------------------------------------------------------------------------------------------------------
// conv (1,13) + per-filter spatial projection:  h[b,f,c,t] = sum_k Wc[f,k] x[b,c,t+k];  v[b,t,f] = sum_c We[f,c] h[b,f,c,t]
// ------------------------------------------------------------------------------------------------------------------
__global__ void conv_embed_fwd_kernel(const float *__restrict__ x, const float *__restrict__ Wc, const float *__restrict__ We,
                                      int C, int T, int F, int K, int Tp, float *__restrict__ h, float *__restrict__ v) {
    const int b = blockIdx.z, f = blockIdx.y, t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= Tp) return;
    float w[16];
    for (int k = 0; k < K; ++k) w[k] = Wc[f * K + k];
    float acc = 0.f;
    for (int c = 0; c < C; ++c) {
        const float *xr = x + ((int64_t)b * C + c) * T + t;
        float s = 0.f;
        for (int k = 0; k < K; ++k) s = fmaf(w[k], xr[k], s);
        h[(((int64_t)b * F + f) * C + c) * Tp + t] = s;
        acc = fmaf(We[f * C + c], s, acc);
    }
    v[((int64_t)b * Tp + t) * F + f] = acc;
}

// dWe[f][c] = sum_{b,t} dv[b,t,f] h[b,f,c,t];  dWc[f][k] = sum_{b,t} dv[b,t,f] * sum_c We[f,c] x[b,c,t+k].  One CTA per filter.
__global__ void __launch_bounds__(256)
conv_embed_bwd_kernel(const float *__restrict__ x, const float *__restrict__ We, const float *__restrict__ h,
                      const float *__restrict__ dv, int B, int C, int T, int F, int K, int Tp, float *__restrict__ dWc,
                      float *__restrict__ dWe) {
    __shared__ float red[256];
    const int f = blockIdx.x, tid = threadIdx.x;
    const int n_out = C + K;                       // outputs of this filter: C spatial + K temporal taps
    for (int o = 0; o < n_out; ++o) {
        float s = 0.f;
        for (int i = tid; i < B * Tp; i += 256) {
            const int b = i / Tp, t = i - b * Tp;
            const float g = dv[((int64_t)b * Tp + t) * F + f];
            if (o < C) {
                s = fmaf(g, h[(((int64_t)b * F + f) * C + o) * Tp + t], s);
            } else {
                const int k = o - C;
                float xe = 0.f;
                for (int c = 0; c < C; ++c) xe = fmaf(We[f * C + c], x[((int64_t)b * C + c) * T + t + k], xe);
                s = fmaf(g, xe, s);
            }
        }
        red[tid] = s;
        __syncthreads();
        for (int w = 128; w > 0; w >>= 1) { if (tid < w) red[tid] += red[tid + w]; __syncthreads(); }
        if (tid == 0) { if (o < C) dWe[f * C + o] = red[0]; else dWc[f * K + (o - C)] = red[0]; }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------------------------
// row softmax (in place) and its backward  dS = scale * P o (dP - sum_j dP o P)   (one warp per row)
// ------------------------------------------------------------------------------------------------------------------
__global__ void softmax_rows_kernel(float *__restrict__ S, int64_t rows, int n) {
    const int lane = threadIdx.x & 31;
    const int64_t r = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= rows) return;
    float *p = S + r * n;
    float mx = -INFINITY;
    for (int j = lane; j < n; j += 32) mx = fmaxf(mx, p[j]);
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float s = 0.f;
    for (int j = lane; j < n; j += 32) { const float e = expf(p[j] - mx); p[j] = e; s += e; }
    s = warp_sum(s);
    const float inv = 1.f / s;
    for (int j = lane; j < n; j += 32) p[j] *= inv;
}
__global__ void softmax_bwd_rows_kernel(const float *__restrict__ P, float *__restrict__ dP, int64_t rows, int n, float scale) {
    const int lane = threadIdx.x & 31;
    const int64_t r = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= rows) return;
    const float *p = P + r * n;
    float *d = dP + r * n;
    float s = 0.f;
    for (int j = lane; j < n; j += 32) s = fmaf(d[j], p[j], s);
    s = warp_sum(s);
    for (int j = lane; j < n; j += 32) d[j] = scale * p[j] * (d[j] - s);
}

// ------------------------------------------------------------------------------------------------------------------
// out = res + keep * LayerNorm(a [+ a2]) * gamma + beta  (keep = mask / (1-p), or 1);  saves (mean, rstd) per row.
// One warp per row, D <= 64.
// ------------------------------------------------------------------------------------------------------------------
__global__ void ln_drop_res_fwd_kernel(const float *__restrict__ a, const float *__restrict__ a2, int lda2,
                                       const float *__restrict__ res, const float *__restrict__ gamma,
                                       const float *__restrict__ beta, const uint8_t *__restrict__ mask, float keep_scale,
                                       int64_t rows, int D, int lda, float eps, float *__restrict__ out,
                                       float *__restrict__ stats) {
    const int lane = threadIdx.x & 31;
    const int64_t r = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= rows) return;
    float v[2];
    float s = 0.f;
#pragma unroll
    for (int e = 0; e < 2; ++e) {
        const int d = lane + 32 * e;
        v[e] = d < D ? a[r * lda + d] + (a2 ? a2[r * lda2 + d] : 0.f) : 0.f;
        s += v[e];
    }
    const float mean = warp_sum(s) / D;
    float q = 0.f;
#pragma unroll
    for (int e = 0; e < 2; ++e) { const int d = lane + 32 * e; if (d < D) q = fmaf(v[e] - mean, v[e] - mean, q); }
    const float rstd = rsqrtf(warp_sum(q) / D + eps);
#pragma unroll
    for (int e = 0; e < 2; ++e) {
        const int d = lane + 32 * e;
        if (d < D) {
            float y = (v[e] - mean) * rstd * gamma[d] + beta[d];
            if (mask) y = mask[r * D + d] ? y * keep_scale : 0.f;
            out[r * D + d] = res[r * D + d] + y;
        }
    }
    if (lane == 0) { stats[2 * r] = mean; stats[2 * r + 1] = rstd; }
}

// Backward of the above w.r.t. the LayerNorm input (da, written; the residual path passes dout through unchanged) and
// per-row-block partial sums of d(gamma), d(beta):  part[blk][2][D].
constexpr int LNB_ROWS = 64;
__global__ void __launch_bounds__(256)
ln_drop_bwd_kernel(const float *__restrict__ dout, const float *__restrict__ a, const float *__restrict__ a2, int lda2,
                   const float *__restrict__ gamma, const uint8_t *__restrict__ mask, float keep_scale,
                   const float *__restrict__ stats, int64_t rows, int D, int lda, float *__restrict__ da,
                   float *__restrict__ part) {
    __shared__ float sg[8][64], sb[8][64];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float gsum[2] = {0.f, 0.f}, bsum[2] = {0.f, 0.f};
    for (int rr = warp; rr < LNB_ROWS; rr += 8) {
        const int64_t r = (int64_t)blockIdx.x * LNB_ROWS + rr;
        if (r >= rows) break;
        const float mean = stats[2 * r], rstd = stats[2 * r + 1];
        float xh[2], dy[2], s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int d = lane + 32 * e;
            xh[e] = dy[e] = 0.f;
            if (d < D) {
                const float v = a[r * lda + d] + (a2 ? a2[r * lda2 + d] : 0.f);
                xh[e] = (v - mean) * rstd;
                float g = dout[r * D + d];
                if (mask) g = mask[r * D + d] ? g * keep_scale : 0.f;
                gsum[e] = fmaf(g, xh[e], gsum[e]);
                bsum[e] += g;
                dy[e] = g * gamma[d];
                s1 += dy[e];
                s2 = fmaf(dy[e], xh[e], s2);
            }
        }
        s1 = warp_sum(s1) / D;
        s2 = warp_sum(s2) / D;
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int d = lane + 32 * e;
            if (d < D) da[r * D + d] = rstd * (dy[e] - s1 - xh[e] * s2);
        }
    }
#pragma unroll
    for (int e = 0; e < 2; ++e) { sg[warp][lane + 32 * e] = gsum[e]; sb[warp][lane + 32 * e] = bsum[e]; }
    __syncthreads();
    if (threadIdx.x < D) {
        float g = 0.f, b = 0.f;
        for (int w = 0; w < 8; ++w) { g += sg[w][threadIdx.x]; b += sb[w][threadIdx.x]; }
        part[((int64_t)blockIdx.x * 2) * D + threadIdx.x] = g;
        part[((int64_t)blockIdx.x * 2 + 1) * D + threadIdx.x] = b;
    }
}

// out[j] = sum_i part[i][j] (fixed order); used for LayerNorm affine gradients and bias gradients
__global__ void colsum_kernel(const float *__restrict__ part, int n_part, int n, int64_t stride, float *__restrict__ out) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    float s = 0.f;
    for (int i = 0; i < n_part; ++i) s += part[(int64_t)i * stride + j];
    out[j] = s;
}
// column sums of a [rows][n] matrix in two deterministic stages: part[blk][n] over 256-row blocks
__global__ void colsum_rows_kernel(const float *__restrict__ X, int64_t rows, int n, float *__restrict__ part) {
    const int j = blockIdx.y * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const int64_t r0 = (int64_t)blockIdx.x * 256, r1 = r0 + 256 < rows ? r0 + 256 : rows;
    float s = 0.f;
    for (int64_t r = r0; r < r1; ++r) s += X[r * n + j];
    part[(int64_t)blockIdx.x * n + j] = s;
}

// f = relu'(f) * keep: in place on the gradient; `act` is the post-ReLU (pre-dropout) activation
__global__ void relu_drop_bwd_kernel(float *__restrict__ g, const float *__restrict__ act, const uint8_t *__restrict__ mask,
                                     float keep_scale, int64_t n) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    float v = act[i] > 0.f ? g[i] : 0.f;
    if (mask) v = mask[i] ? v * keep_scale : 0.f;
    g[i] = v;
}
__global__ void drop_fwd_kernel(float *__restrict__ a, const uint8_t *__restrict__ mask, float keep_scale, int64_t n) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) a[i] = mask[i] ? a[i] * keep_scale : 0.f;
}
__global__ void add_kernel(float *__restrict__ a, const float *__restrict__ b, int64_t n) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) a[i] += b[i];
}

// ------------------------------------------------------------------------------------------------------------------
// Head.  BatchNorm2d(F) over (B, Tp) per channel f of v[b,t,f]; one CTA per channel.
// stats[f] = {mean, invstd, scale, shift}
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
head_bn_stats_kernel(const float *__restrict__ v, int B, int Tp, int F, const float *__restrict__ gamma,
                     const float *__restrict__ beta, float *__restrict__ rm, float *__restrict__ rv, int train, float eps,
                     float momentum, float4 *__restrict__ stats) {
    __shared__ double r1[256], r2[256];
    const int f = blockIdx.x, tid = threadIdx.x;
    float mean, var;
    if (train) {
        double s = 0.0, q = 0.0;
        for (int i = tid; i < B * Tp; i += 256) { const double x = v[(int64_t)i * F + f]; s += x; q += x * x; }
        r1[tid] = s; r2[tid] = q;
        __syncthreads();
        for (int w = 128; w > 0; w >>= 1) { if (tid < w) { r1[tid] += r1[tid + w]; r2[tid] += r2[tid + w]; } __syncthreads(); }
        const double cnt = (double)B * Tp, mu = r1[0] / cnt;
        double vb = r2[0] / cnt - mu * mu;
        if (vb < 0.0) vb = 0.0;
        mean = (float)mu; var = (float)vb;
        if (tid == 0) {
            rm[f] = (1.f - momentum) * rm[f] + momentum * mean;
            rv[f] = (1.f - momentum) * rv[f] + momentum * (float)(cnt > 1.0 ? vb * cnt / (cnt - 1.0) : vb);
        }
    } else { mean = rm[f]; var = rv[f]; }
    if (tid == 0) {
        const float invstd = 1.f / sqrtf(var + eps), scale = gamma[f] * invstd;
        stats[f] = make_float4(mean, invstd, scale, beta[f] - mean * scale);
    }
}

// feat[b][f*U + u] = keep * log(clamp(mean_{w<P} bn(v[b, u*S + w, f])^2, 1e-7, 1e4));  pooled (pre-log) value saved
__global__ void head_pool_fwd_kernel(const float *__restrict__ v, const float4 *__restrict__ stats, int B, int Tp, int F,
                                     int P, int S, int U, const uint8_t *__restrict__ mask, float keep_scale,
                                     float *__restrict__ pooled, float *__restrict__ feat) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= (int64_t)B * F * U) return;
    const int u = (int)(i % U), f = (int)((i / U) % F), b = (int)(i / ((int64_t)U * F));
    const float4 st = stats[f];
    float s = 0.f;
    for (int w = 0; w < P; ++w) { const float z = fmaf(v[((int64_t)b * Tp + u * S + w) * F + f], st.z, st.w); s = fmaf(z, z, s); }
    s /= (float)P;
    pooled[i] = s;
    float y = logf(fminf(fmaxf(s, 1e-7f), 1e4f));
    if (mask) y = mask[i] ? y * keep_scale : 0.f;
    feat[i] = y;
}

__global__ void softmax_small_kernel(float *__restrict__ z, int B, int NC) {      // in place, one thread per sample
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    float *p = z + (int64_t)b * NC;
    float mx = p[0];
    for (int j = 1; j < NC; ++j) mx = fmaxf(mx, p[j]);
    float s = 0.f;
    for (int j = 0; j < NC; ++j) { p[j] = expf(p[j] - mx); s += p[j]; }
    for (int j = 0; j < NC; ++j) p[j] /= s;
}
__global__ void softmax_small_bwd_kernel(const float *__restrict__ p, const float *__restrict__ dp, int B, int NC,
                                         float *__restrict__ dz) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    float s = 0.f;
    for (int j = 0; j < NC; ++j) s = fmaf(dp[(int64_t)b * NC + j], p[(int64_t)b * NC + j], s);
    for (int j = 0; j < NC; ++j) dz[(int64_t)b * NC + j] = p[(int64_t)b * NC + j] * (dp[(int64_t)b * NC + j] - s);
}

// dpooled[i] = dfeat[i] * keep / pooled (inside the clamp), in place on dfeat
__global__ void head_log_bwd_kernel(float *__restrict__ dfeat, const float *__restrict__ pooled,
                                    const uint8_t *__restrict__ mask, float keep_scale, int64_t n) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    float g = dfeat[i];
    if (mask) g = mask[i] ? g * keep_scale : 0.f;
    const float s = pooled[i];
    dfeat[i] = (s > 1e-7f && s < 1e4f) ? g / s : 0.f;
}

// dz[b,t,f] (gradient w.r.t. the BN output) = 2 z / P * sum_{u: window covers t} dpooled[b,f,u];
// per-channel partial sums (sum dz, sum dz*xhat) per sample: part[b][f][2]
__global__ void __launch_bounds__(256)
head_pool_bwd_kernel(const float *__restrict__ v, const float4 *__restrict__ stats, const float *__restrict__ dpooled, int Tp,
                     int F, int P, int S, int U, float *__restrict__ dz, float *__restrict__ part) {
    __shared__ float r1[256], r2[256];
    const int b = blockIdx.y, f = blockIdx.x, tid = threadIdx.x;
    const float4 st = stats[f];
    float s1 = 0.f, s2 = 0.f;
    for (int t = tid; t < Tp; t += 256) {
        const float x = v[((int64_t)b * Tp + t) * F + f];
        const float z = fmaf(x, st.z, st.w);
        int u_hi = t / S; if (u_hi > U - 1) u_hi = U - 1;
        int u_lo = (t - P + S) / S; if (t - P + 1 <= 0) u_lo = 0; if (u_lo < 0) u_lo = 0;
        float g = 0.f;
        for (int u = u_lo; u <= u_hi; ++u)
            if (u * S <= t && t < u * S + P) g += dpooled[((int64_t)b * F + f) * U + u];
        const float d = 2.f * z * g / (float)P;
        dz[((int64_t)b * Tp + t) * F + f] = d;
        s1 += d;
        s2 = fmaf(d, (x - st.x) * st.y, s2);
    }
    r1[tid] = s1; r2[tid] = s2;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) { if (tid < w) { r1[tid] += r1[tid + w]; r2[tid] += r2[tid + w]; } __syncthreads(); }
    if (tid == 0) { part[((int64_t)b * F + f) * 2] = r1[0]; part[((int64_t)b * F + f) * 2 + 1] = r2[0]; }
}

// BN backward: dv = k (dz - c1 - xhat c2) (train) or k dz (eval), in place; d(gamma), d(beta)
__global__ void head_bn_bwd_kernel(float *__restrict__ dz, const float *__restrict__ v, const float4 *__restrict__ stats,
                                   const float *__restrict__ part, const float *__restrict__ gamma, int B, int Tp, int F,
                                   int train, float *__restrict__ dgamma, float *__restrict__ dbeta) {
    const int f = blockIdx.x, tid = threadIdx.x;
    double s1 = 0.0, s2 = 0.0;
    for (int b = 0; b < B; ++b) { s1 += part[((int64_t)b * F + f) * 2]; s2 += part[((int64_t)b * F + f) * 2 + 1]; }
    if (tid == 0) { dgamma[f] = (float)s2; dbeta[f] = (float)s1; }
    const float4 st = stats[f];
    const float k = gamma[f] * st.y, cnt = (float)B * Tp;
    const float c1 = train ? (float)(s1 / cnt) : 0.f, c2 = train ? (float)(s2 / cnt) : 0.f;
    for (int i = tid; i < B * Tp; i += blockDim.x) {
        const float xh = (v[(int64_t)i * F + f] - st.x) * st.y;
        dz[(int64_t)i * F + f] = k * (dz[(int64_t)i * F + f] - c1 - xh * c2);
    }
}

struct Dims {
    int B, C, T, F, K, Tp, L, H, P, S, U, NC, FEAT;
    int train, drop;
    float p, keep_scale, bn_eps, bn_mom, ln_eps;
    int64_t R;     // rows of the token matrix: B * Tp
    // parameter offsets
    int64_t oWc, og, ob, oWe, oL, layer_sz, oFc, n_params;
    // inside a layer
    int64_t lWqkv, lW1, lb1, lW2, lb2, lg1, lbe1, lg2, lbe2;
};

int make(const eav_shallow_cfg *c, Dims *d) {
    EAV_REQUIRE(c != nullptr, EAV_ERR_BAD_ARG, "shallow cfg is NULL");
    EAV_REQUIRE(c->batch > 0 && c->chans > 0 && c->samples > 0 && c->n_filters > 0 && c->kern > 0 && c->n_layers >= 0 &&
                    c->ffn > 0 && c->pool > 0 && c->stride > 0 && c->n_classes > 0, EAV_ERR_BAD_ARG,
                "shallow: all dimensions must be positive");
    EAV_REQUIRE(c->n_filters <= 64 && c->kern <= 16, EAV_ERR_UNSUPPORTED, "shallow: n_filters <= 64 and kern <= 16");
    EAV_REQUIRE(c->dropout_mode == 0 || c->dropout_mode == 1, EAV_ERR_BAD_ARG, "shallow: dropout_mode 0 (none) or 1 (masks)");
    d->B = c->batch; d->C = c->chans; d->T = c->samples; d->F = c->n_filters; d->K = c->kern; d->Tp = d->T - d->K + 1;
    d->L = c->n_layers; d->H = c->ffn; d->P = c->pool; d->S = c->stride; d->NC = c->n_classes;
    EAV_REQUIRE(d->Tp >= d->P, EAV_ERR_BAD_ARG, "shallow: samples too short for the pooling window");
    d->U = (d->Tp - d->P) / d->S + 1; d->FEAT = d->F * d->U;
    d->train = c->bn_train != 0; d->drop = c->dropout_mode == 1 && c->dropout_p > 0.f;
    d->p = c->dropout_p; d->keep_scale = c->dropout_p < 1.f ? 1.f / (1.f - c->dropout_p) : 0.f;
    d->bn_eps = c->bn_eps; d->bn_mom = c->bn_momentum; d->ln_eps = c->ln_eps;
    d->R = (int64_t)d->B * d->Tp;
    int64_t o = 0;
    d->oWc = o; o += (int64_t)d->F * d->K;
    d->og = o; o += d->F;
    d->ob = o; o += d->F;
    d->oWe = o; o += (int64_t)d->F * d->C;
    d->oL = o;
    int64_t l = 0;
    d->lWqkv = l; l += 3ll * d->F * d->F;
    d->lW1 = l; l += (int64_t)d->H * d->F;
    d->lb1 = l; l += d->H;
    d->lW2 = l; l += (int64_t)d->F * d->H;
    d->lb2 = l; l += d->F;
    d->lg1 = l; l += d->F; d->lbe1 = l; l += d->F; d->lg2 = l; l += d->F; d->lbe2 = l; l += d->F;
    d->layer_sz = l;
    o += l * d->L;
    d->oFc = o; o += (int64_t)d->NC * d->FEAT;
    d->n_params = o;
    return 0;
}

// workspace (floats).  Per layer: vin qkv P lnin1(=attn out) st1 vmid f1 f2 st2; globals: h, v0.., head buffers, scratch.
struct Ws {
    size_t h, v, layer, layer_sz, l_qkv, l_P, l_a, l_st1, l_vmid, l_f1, l_f2, l_st2;
    size_t bnstats, pooled, feat, probs, g0, g1, gH, gqkv, gP, part, total;
};
Ws layout(const Dims &d) {
    Ws w;
    size_t o = 0;
    auto take = [&](size_t n) { size_t at = o; o += (n + 63) / 64 * 64; return at; };
    const size_t R = (size_t)d.R, F = d.F;
    w.h = take((size_t)d.B * d.F * d.C * d.Tp);
    w.v = take(R * F * (d.L + 1));                         // v[0] = embedding, v[l+1] = output of layer l
    size_t l = 0;
    auto tl = [&](size_t n) { size_t at = l; l += (n + 63) / 64 * 64; return at; };
    w.l_qkv = tl(R * 3 * F); w.l_P = tl((size_t)d.B * d.Tp * d.Tp); w.l_a = tl(R * F); w.l_st1 = tl(R * 2);
    w.l_vmid = tl(R * F); w.l_f1 = tl(R * d.H); w.l_f2 = tl(R * F); w.l_st2 = tl(R * 2);
    w.layer_sz = l;
    w.layer = take(l * d.L);
    w.bnstats = take(4 * F); w.pooled = take((size_t)d.B * d.FEAT); w.feat = take((size_t)d.B * d.FEAT);
    w.probs = take((size_t)d.B * d.NC);
    w.g0 = take(R * F); w.g1 = take(R * F); w.gH = take(R * d.H); w.gqkv = take(R * 3 * F);
    w.gP = take((size_t)d.B * d.Tp * d.Tp);
    size_t np = (size_t)cdiv64(d.R, LNB_ROWS) * 2 * F;
    const size_t np2 = (size_t)cdiv64(d.R, 256) * (d.H > 3 * d.F ? d.H : 3 * d.F);
    if (np2 > np) np = np2;
    if ((size_t)d.B * F * 2 > np) np = (size_t)d.B * F * 2;
    w.part = take(np);
    w.total = o * sizeof(float);
    return w;
}

inline unsigned nblk(int64_t n, int per) { return (unsigned)((n + per - 1) / per); }

}  // namespace
}  // namespace eav

using namespace eav;

extern "C" int64_t eav_shallow_param_layout(const eav_shallow_cfg *cfg, int64_t *offsets, int64_t *layer_stride) {
    Dims d;
    int rc = make(cfg, &d);
    if (rc) return rc;
    if (offsets) {
        // conv bn.w bn.b emb | (first layer) Wqkv W1 b1 W2 b2 g1 be1 g2 be2 | fc
        const int64_t o[14] = {d.oWc, d.og, d.ob, d.oWe, d.oL + d.lWqkv, d.oL + d.lW1, d.oL + d.lb1, d.oL + d.lW2, d.oL + d.lb2,
                               d.oL + d.lg1, d.oL + d.lbe1, d.oL + d.lg2, d.oL + d.lbe2, d.oFc};
        for (int i = 0; i < 14; ++i) offsets[i] = o[i];
    }
    if (layer_stride) *layer_stride = d.layer_sz;
    return d.n_params;
}

extern "C" size_t eav_shallow_workspace_bytes(const eav_shallow_cfg *cfg) {
    Dims d;
    if (make(cfg, &d)) return 0;
    return layout(d).total;
}

// masks: [L][ m1 (R*F) | mffn (R*H) | m2 (R*F) ] then the head mask (B*FEAT), uint8 keep flags, or NULL
static inline const uint8_t *mask_at(const Dims &d, const uint8_t *masks, int layer, int which) {
    if (!d.drop || masks == nullptr) return nullptr;
    const size_t per = (size_t)d.R * (2 * d.F + d.H);
    if (layer >= d.L) return masks + per * d.L;
    const size_t off = which == 0 ? 0 : which == 1 ? (size_t)d.R * d.F : (size_t)d.R * (d.F + d.H);
    return masks + per * layer + off;
}

extern "C" int eav_shallow_forward(const eav_shallow_cfg *cfg, const float *x, const float *params, float *bn_state,
                                   const uint8_t *masks, float *out, void *workspace, size_t workspace_bytes,
                                   void *stream) {
    Dims d;
    int rc = make(cfg, &d);
    if (rc) return rc;
    const Ws w = layout(d);
    EAV_REQUIRE(x && params && bn_state && out && workspace, EAV_ERR_BAD_ARG, "shallow_forward: null pointer");
    EAV_REQUIRE(workspace_bytes >= w.total, EAV_ERR_WORKSPACE, "shallow_forward: workspace %zu < %zu", workspace_bytes, w.total);
    EAV_REQUIRE(!(cfg->dropout_mode == 1 && masks == nullptr), EAV_ERR_BAD_ARG, "shallow_forward: dropout masks required");
    cudaStream_t st = (cudaStream_t)stream;
    float *ws = reinterpret_cast<float *>(workspace);
    const Gemm G{st};
    const int F = d.F, Tp = d.Tp;
    const int64_t R = d.R;
    conv_embed_fwd_kernel<<<dim3(cdiv(Tp, 128), F, d.B), 128, 0, st>>>(x, params + d.oWc, params + d.oWe, d.C, d.T, F, d.K, Tp,
                                                                       ws + w.h, ws + w.v);
    EAV_CUDA_LAUNCH_CHECK("shallow_conv_embed");
    const float scale = 1.f / sqrtf((float)F);
    for (int l = 0; l < d.L; ++l) {
        const float *pl = params + d.oL + (int64_t)l * d.layer_sz;
        float *L = ws + w.layer + (size_t)l * w.layer_sz;
        const float *vin = ws + w.v + (size_t)l * R * F;
        float *vout = ws + w.v + (size_t)(l + 1) * R * F;
        float *qkv = L + w.l_qkv, *P = L + w.l_P, *a = L + w.l_a, *vmid = L + w.l_vmid, *f1 = L + w.l_f1, *f2 = L + w.l_f2;
        TRY_RC(G.run((int)R, 3 * F, F, 1.f, vin, F, 0, pl + d.lWqkv, F, 1, 0.f, qkv, 3 * F));                       // q | k | v
        TRY_RC(G.run(Tp, Tp, F, scale, qkv, 3 * F, 0, qkv + F, 3 * F, 1, 0.f, P, Tp, nullptr, 0, d.B,
                     (int64_t)Tp * 3 * F, (int64_t)Tp * 3 * F, (int64_t)Tp * Tp));                                   // Q K^T / sqrt(d)
        softmax_rows_kernel<<<nblk(R, 8), 256, 0, st>>>(P, R, Tp);
        EAV_CUDA_LAUNCH_CHECK("shallow_softmax");
        TRY_RC(G.run(Tp, F, Tp, 1.f, P, Tp, 0, qkv + 2 * F, 3 * F, 0, 0.f, a, F, nullptr, 0, d.B, (int64_t)Tp * Tp,
                     (int64_t)Tp * 3 * F, (int64_t)Tp * F));                                                         // P V
        ln_drop_res_fwd_kernel<<<nblk(R, 8), 256, 0, st>>>(a, qkv + 2 * F, 3 * F, vin, pl + d.lg1, pl + d.lbe1,
                                                          mask_at(d, masks, l, 0), d.keep_scale, R, F, F, d.ln_eps, vmid,
                                                          L + w.l_st1);                                              // + V, LN1, drop, residual
        EAV_CUDA_LAUNCH_CHECK("shallow_ln1");
        TRY_RC(G.run((int)R, d.H, F, 1.f, vmid, F, 0, pl + d.lW1, F, 1, 0.f, f1, d.H, pl + d.lb1, 1));               // relu(W1 v + b1)
        const uint8_t *mf = mask_at(d, masks, l, 1);
        const float *f1d = f1;
        if (mf) {      // the dropped copy feeds W2; the undropped post-ReLU values stay for relu'
            float *tmp = ws + w.gH;
            cudaMemcpyAsync(tmp, f1, (size_t)R * d.H * sizeof(float), cudaMemcpyDeviceToDevice, st);
            drop_fwd_kernel<<<nblk(R * d.H, 256), 256, 0, st>>>(tmp, mf, d.keep_scale, R * d.H);
            EAV_CUDA_LAUNCH_CHECK("shallow_ffn_drop");
            f1d = tmp;
        }
        TRY_RC(G.run((int)R, F, d.H, 1.f, f1d, d.H, 0, pl + d.lW2, d.H, 1, 0.f, f2, F, pl + d.lb2, 0));
        ln_drop_res_fwd_kernel<<<nblk(R, 8), 256, 0, st>>>(f2, nullptr, 0, vmid, pl + d.lg2, pl + d.lbe2,
                                                          mask_at(d, masks, l, 2), d.keep_scale, R, F, F, d.ln_eps, vout,
                                                          L + w.l_st2);
        EAV_CUDA_LAUNCH_CHECK("shallow_ln2");
    }
    const float *vL = ws + w.v + (size_t)d.L * R * F;
    float4 *bst = reinterpret_cast<float4 *>(ws + w.bnstats);
    head_bn_stats_kernel<<<F, 256, 0, st>>>(vL, d.B, Tp, F, params + d.og, params + d.ob, bn_state, bn_state + F, d.train,
                                            d.bn_eps, d.bn_mom, bst);
    EAV_CUDA_LAUNCH_CHECK("shallow_bn_stats");
    head_pool_fwd_kernel<<<nblk((int64_t)d.B * d.FEAT, 256), 256, 0, st>>>(vL, bst, d.B, Tp, F, d.P, d.S, d.U,
                                                                           mask_at(d, masks, d.L, 0), d.keep_scale,
                                                                           ws + w.pooled, ws + w.feat);
    EAV_CUDA_LAUNCH_CHECK("shallow_pool");
    TRY_RC(G.run(d.B, d.NC, d.FEAT, 1.f, ws + w.feat, d.FEAT, 0, params + d.oFc, d.FEAT, 1, 0.f, out, d.NC));
    softmax_small_kernel<<<cdiv(d.B, 64), 64, 0, st>>>(out, d.B, d.NC);
    EAV_CUDA_LAUNCH_CHECK("shallow_softmax_out");
    cudaMemcpyAsync(ws + w.probs, out, (size_t)d.B * d.NC * sizeof(float), cudaMemcpyDeviceToDevice, st);
    return 0;
}

extern "C" int eav_shallow_backward(const eav_shallow_cfg *cfg, const float *x, const float *params, const float *dout,
                                    const uint8_t *masks, float *grads, void *workspace, size_t workspace_bytes,
                                    void *stream) {
    Dims d;
    int rc = make(cfg, &d);
    if (rc) return rc;
    const Ws w = layout(d);
    EAV_REQUIRE(x && params && dout && grads && workspace, EAV_ERR_BAD_ARG, "shallow_backward: null pointer");
    EAV_REQUIRE(workspace_bytes >= w.total, EAV_ERR_WORKSPACE, "shallow_backward: workspace %zu < %zu", workspace_bytes, w.total);
    cudaStream_t st = (cudaStream_t)stream;
    float *ws = reinterpret_cast<float *>(workspace);
    const Gemm G{st};
    const int F = d.F, Tp = d.Tp, H = d.H;
    const int64_t R = d.R;
    float *g0 = ws + w.g0, *g1 = ws + w.g1, *gH = ws + w.gH, *gqkv = ws + w.gqkv, *gP = ws + w.gP, *part = ws + w.part;
    const float *vL = ws + w.v + (size_t)d.L * R * F;
    const float4 *bst = reinterpret_cast<const float4 *>(ws + w.bnstats);
    // ---- head: softmax -> fc -> dropout/log/pool/square -> BatchNorm
    float *dzc = g1;                                            // [B][NC] scratch
    softmax_small_bwd_kernel<<<cdiv(d.B, 64), 64, 0, st>>>(ws + w.probs, dout, d.B, d.NC, dzc);
    EAV_CUDA_LAUNCH_CHECK("shallow_softmax_out_bwd");
    TRY_RC(G.run(d.NC, d.FEAT, d.B, 1.f, dzc, d.NC, 1, ws + w.feat, d.FEAT, 0, 0.f, grads + d.oFc, d.FEAT));        // dWfc = dz^T feat
    float *dfeat = gH;                                          // [B][FEAT]
    TRY_RC(G.run(d.B, d.FEAT, d.NC, 1.f, dzc, d.NC, 0, params + d.oFc, d.FEAT, 0, 0.f, dfeat, d.FEAT));            // dfeat = dz Wfc
    head_log_bwd_kernel<<<nblk((int64_t)d.B * d.FEAT, 256), 256, 0, st>>>(dfeat, ws + w.pooled, mask_at(d, masks, d.L, 0),
                                                                          d.keep_scale, (int64_t)d.B * d.FEAT);
    EAV_CUDA_LAUNCH_CHECK("shallow_log_bwd");
    head_pool_bwd_kernel<<<dim3(F, d.B), 256, 0, st>>>(vL, bst, dfeat, Tp, F, d.P, d.S, d.U, g0, part);
    EAV_CUDA_LAUNCH_CHECK("shallow_pool_bwd");
    head_bn_bwd_kernel<<<F, 256, 0, st>>>(g0, vL, bst, part, params + d.og, d.B, Tp, F, d.train, grads + d.og, grads + d.ob);
    EAV_CUDA_LAUNCH_CHECK("shallow_bn_bwd");
    // g0 = d(loss)/d(v_L)
    const float scale = 1.f / sqrtf((float)F);
    const int nlb = (int)cdiv64(R, LNB_ROWS), nrb = (int)cdiv64(R, 256);
    for (int l = d.L - 1; l >= 0; --l) {
        const float *pl = params + d.oL + (int64_t)l * d.layer_sz;
        float *gl = grads + d.oL + (int64_t)l * d.layer_sz;
        float *L = ws + w.layer + (size_t)l * w.layer_sz;
        const float *vin = ws + w.v + (size_t)l * R * F;
        const float *qkv = L + w.l_qkv, *P = L + w.l_P, *a = L + w.l_a, *vmid = L + w.l_vmid, *f1 = L + w.l_f1, *f2 = L + w.l_f2;
        // out = vmid + drop(LN2(f2)):  g0 = dout;  df2 -> g1
        ln_drop_bwd_kernel<<<nlb, 256, 0, st>>>(g0, f2, nullptr, 0, pl + d.lg2, mask_at(d, masks, l, 2), d.keep_scale,
                                                L + w.l_st2, R, F, F, g1, part);
        EAV_CUDA_LAUNCH_CHECK("shallow_ln2_bwd");
        colsum_kernel<<<cdiv(2 * F, 128), 128, 0, st>>>(part, nlb, 2 * F, 2 * F, gl + d.lg2);   // g2 then be2 (contiguous)
        EAV_CUDA_LAUNCH_CHECK("shallow_ln2_affine");
        // f2 = f1d W2^T + b2
        const uint8_t *mf = mask_at(d, masks, l, 1);
        const float *f1d = f1;
        if (mf) {      // recompute the dropped activations (the forward's copy was scratch)
            float *tmp = gqkv;       // R*3F >= R*H? not in general: use gP (B*Tp*Tp floats >= R*H for Tp >= H)
            tmp = gP;
            cudaMemcpyAsync(tmp, f1, (size_t)R * H * sizeof(float), cudaMemcpyDeviceToDevice, st);
            drop_fwd_kernel<<<nblk(R * H, 256), 256, 0, st>>>(tmp, mf, d.keep_scale, R * H);
            EAV_CUDA_LAUNCH_CHECK("shallow_ffn_drop");
            f1d = tmp;
        }
        TRY_RC(G.run(F, H, (int)R, 1.f, g1, F, 1, f1d, H, 0, 0.f, gl + d.lW2, H));                                  // dW2 = df2^T f1d
        colsum_rows_kernel<<<dim3(nrb, cdiv(F, 64)), 64, 0, st>>>(g1, R, F, part);
        colsum_kernel<<<cdiv(F, 128), 128, 0, st>>>(part, nrb, F, F, gl + d.lb2);
        EAV_CUDA_LAUNCH_CHECK("shallow_b2");
        TRY_RC(G.run((int)R, H, F, 1.f, g1, F, 0, pl + d.lW2, H, 0, 0.f, gH, H));                                   // df1d = df2 W2
        relu_drop_bwd_kernel<<<nblk(R * H, 256), 256, 0, st>>>(gH, f1, mf, d.keep_scale, R * H);
        EAV_CUDA_LAUNCH_CHECK("shallow_relu_bwd");
        TRY_RC(G.run(H, F, (int)R, 1.f, gH, H, 1, vmid, F, 0, 0.f, gl + d.lW1, F));                                 // dW1 = df1^T vmid
        colsum_rows_kernel<<<dim3(nrb, cdiv(H, 64)), 64, 0, st>>>(gH, R, H, part);
        colsum_kernel<<<cdiv(H, 128), 128, 0, st>>>(part, nrb, H, H, gl + d.lb1);
        EAV_CUDA_LAUNCH_CHECK("shallow_b1");
        TRY_RC(G.run((int)R, F, H, 1.f, gH, H, 0, pl + d.lW1, F, 0, 1.f, g0, F));                                   // dvmid = dout + df1 W1
        // vmid = vin + drop(LN1(a + V)):  da -> g1
        ln_drop_bwd_kernel<<<nlb, 256, 0, st>>>(g0, a, qkv + 2 * F, 3 * F, pl + d.lg1, mask_at(d, masks, l, 0), d.keep_scale,
                                                L + w.l_st1, R, F, F, g1, part);
        EAV_CUDA_LAUNCH_CHECK("shallow_ln1_bwd");
        colsum_kernel<<<cdiv(2 * F, 128), 128, 0, st>>>(part, nlb, 2 * F, 2 * F, gl + d.lg1);
        EAV_CUDA_LAUNCH_CHECK("shallow_ln1_affine");
        // a = P V + V:  dV = P^T da + da;  dP = da V^T;  dS = scale * P o (dP - rowsum);  dQ = dS K;  dK = dS^T Q
        const int64_t sq = (int64_t)Tp * 3 * F, sp = (int64_t)Tp * Tp, sa = (int64_t)Tp * F;
        cudaMemcpy2DAsync(gqkv + 2 * F, (size_t)3 * F * sizeof(float), g1, (size_t)F * sizeof(float), (size_t)F * sizeof(float),
                          (size_t)R, cudaMemcpyDeviceToDevice, st);                                                  // dV = da
        TRY_RC(G.run(Tp, F, Tp, 1.f, P, Tp, 1, g1, F, 0, 1.f, gqkv + 2 * F, 3 * F, nullptr, 0, d.B, sp, sa, sq));  // += P^T da
        TRY_RC(G.run(Tp, Tp, F, 1.f, g1, F, 0, qkv + 2 * F, 3 * F, 1, 0.f, gP, Tp, nullptr, 0, d.B, sa, sq, sp));  // dP = da V^T
        softmax_bwd_rows_kernel<<<nblk(R, 8), 256, 0, st>>>(P, gP, R, Tp, scale);
        EAV_CUDA_LAUNCH_CHECK("shallow_softmax_bwd");
        TRY_RC(G.run(Tp, F, Tp, 1.f, gP, Tp, 0, qkv + F, 3 * F, 0, 0.f, gqkv, 3 * F, nullptr, 0, d.B, sp, sq, sq));     // dQ = dS K
        TRY_RC(G.run(Tp, F, Tp, 1.f, gP, Tp, 1, qkv, 3 * F, 0, 0.f, gqkv + F, 3 * F, nullptr, 0, d.B, sp, sq, sq));     // dK = dS^T Q
        TRY_RC(G.run(3 * F, F, (int)R, 1.f, gqkv, 3 * F, 1, vin, F, 0, 0.f, gl + d.lWqkv, F));                      // dWqkv = dqkv^T vin
        TRY_RC(G.run((int)R, F, 3 * F, 1.f, gqkv, 3 * F, 0, pl + d.lWqkv, F, 0, 1.f, g0, F));                       // dvin = dvmid + dqkv Wqkv
    }
    conv_embed_bwd_kernel<<<F, 256, 0, st>>>(x, params + d.oWe, ws + w.h, g0, d.B, d.C, d.T, F, d.K, Tp, grads + d.oWc,
                                             grads + d.oWe);
    EAV_CUDA_LAUNCH_CHECK("shallow_conv_embed_bwd");
    return 0;
}

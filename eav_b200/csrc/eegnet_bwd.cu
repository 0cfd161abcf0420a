// Backward kernels of the EEGNet path for sm_100a: the gradient of
// CNN_torch/EEGNet_tor.py:50-67 (and CNN_torch/CNN_EEG.py:57-67) w.r.t. all 11 (12)
// parameter tensors.  BatchNorm backward is split the usual way: producers emit
// dz = dL/d(BN output) together with per-row partial sums (sum dz, sum dz*xhat);
// bn_bwd_finalize turns them into {k, c1, c2}; consumers rebuild
//     dL/d(BN input) = k * (dz - c1 - xhat * c2)        (c1 = c2 = 0 in eval mode)
// on the fly while staging their tiles.  All reductions are two-stage and ordered, so
// results are deterministic run to run.
#include "eegnet_kernels.cuh"

namespace eav {

// =================================================================================
// Tail backward: softmax Jacobian (variant 0), dense input gradient (with the
// re-normed weight), dropout, AvgPool(1,P2), ELU' -> dz3 + BN3 partials.
// =================================================================================
__global__ void __launch_bounds__(128)
tail_bwd_kernel(const float *__restrict__ dout, const float *__restrict__ probs,
                const float *__restrict__ params, int64_t pstride, int64_t oWd, const float *__restrict__ y3,
                const float4 *__restrict__ bn3, const uint8_t *__restrict__ mask2, int B, int F2, int T4,
                int T32, int P2, int NC, int softmax_out, int dropout_mode, float p_drop, uint64_t seed,
                uint64_t step, const unsigned long long *__restrict__ step_ptr, float *__restrict__ dz, float *__restrict__ dz3, float *__restrict__ part,
                int fold_bn) {
    // fold_bn (eval-mode BatchNorm-3, variant 0): BN backward is the per-channel scale k = gamma * invstd = bn3[].z, known
    // before this kernel starts -- dz3 leaves as dy3 = k * g and the in-place bn_bwd_apply pass (43 MB read + write) is skipped;
    // the d(gamma) / d(beta) sums are taken over the unscaled g as before.
    extern __shared__ float sm[];  // dfeat_s[FEAT] + dz_s[NC]
    const int FEAT = F2 * T32;
    float *dfeat_s = sm, *dz_s = sm + FEAT;
    const int n = blockIdx.x, m = n / B, tid = threadIdx.x;
    if (step_ptr) step = *step_ptr;
    if (tid == 0) {
        if (softmax_out) {
            float dot = 0.f;
            for (int j = 0; j < NC; ++j) dot = fmaf(dout[(int64_t)n * NC + j], probs[(int64_t)n * NC + j], dot);
            for (int j = 0; j < NC; ++j) {
                float v = probs[(int64_t)n * NC + j] * (dout[(int64_t)n * NC + j] - dot);
                dz_s[j] = v;
                dz[(int64_t)n * NC + j] = v;
            }
        } else {
            for (int j = 0; j < NC; ++j) {
                float v = dout[(int64_t)n * NC + j];
                dz_s[j] = v;
                dz[(int64_t)n * NC + j] = v;
            }
        }
    }
    __syncthreads();
    const float inv_keep = (dropout_mode != EAV_DROPOUT_NONE && p_drop < 1.f) ? 1.f / (1.f - p_drop) : 1.f;
    const float *Wd = params + (int64_t)m * pstride + oWd;
    for (int i = tid; i < FEAT; i += blockDim.x) {
        float s = 0.f;
        for (int j = 0; j < NC; ++j) s = fmaf(Wd[(int64_t)j * FEAT + i], dz_s[j], s);
        int64_t e = (int64_t)n * FEAT + i;
        if (dropout_mode == EAV_DROPOUT_MASK) s = mask2[e] ? s * inv_keep : 0.f;
        else if (dropout_mode == EAV_DROPOUT_PHILOX) s = philox_keep(seed, step, 2u, (uint64_t)e, p_drop) ? s * inv_keep : 0.f;
        else if (dropout_mode == EAV_DROPOUT_PHILOX_2D) s = philox_keep(seed, step, 18u, (uint64_t)((int64_t)n * F2 + i / T32), p_drop) ? s * inv_keep : 0.f;
        dfeat_s[i] = s * (1.f / (float)P2);
    }
    __syncthreads();
    const int warp = tid >> 5, lane = tid & 31;
    auto finish_row = [&](int o, const float4 st, float s1, float s2) {
        s1 = warp_sum(s1);
        s2 = warp_sum(s2);
        if (lane == 0) {
            part[((int64_t)n * F2 + o) * 2] = s1;
            part[((int64_t)n * F2 + o) * 2 + 1] = s2;
        }
    };
    if (T4 <= 128) {
        // four rows per pass: 16 independent loads in flight per lane before the first dependent use
        for (int o0 = warp * 4; o0 < F2; o0 += (blockDim.x >> 5) * 4) {
            float yv[4][4];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const float *src = y3 + ((int64_t)n * F2 + min(o0 + r, F2 - 1)) * T4;
#pragma unroll
                for (int q = 0; q < 4; ++q) yv[r][q] = (lane + 32 * q < T4) ? src[lane + 32 * q] : 0.f;
            }
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int o = o0 + r;
                if (o >= F2) break;
                const float4 st = bn3[(int64_t)m * F2 + o];
                float *dst = dz3 + ((int64_t)n * F2 + o) * T4;
                float s1 = 0.f, s2 = 0.f;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int u = lane + 32 * q;
                    if (u < T4) {
                        const int v = u / P2;
                        const float y = yv[r][q];
                        const float g = (v < T32) ? dfeat_s[o * T32 + v] * elu_grad_from_pre(fmaf(y, st.z, st.w)) : 0.f;
                        dst[u] = fold_bn ? g * st.z : g;
                        s1 += g;
                        s2 = fmaf(g, (y - st.x) * st.y, s2);
                    }
                }
                finish_row(o, st, s1, s2);
            }
        }
        return;
    }
    for (int o = warp; o < F2; o += blockDim.x >> 5) {
        const float4 st = bn3[(int64_t)m * F2 + o];
        const float *src = y3 + ((int64_t)n * F2 + o) * T4;
        float *dst = dz3 + ((int64_t)n * F2 + o) * T4;
        float s1 = 0.f, s2 = 0.f;
        for (int u = lane; u < T4; u += 32) {
            int v = u / P2;
            float y = src[u];
            float g = (v < T32) ? dfeat_s[o * T32 + v] * elu_grad_from_pre(fmaf(y, st.z, st.w)) : 0.f;
            dst[u] = fold_bn ? g * st.z : g;
            s1 += g;
            s2 = fmaf(g, (y - st.x) * st.y, s2);
        }
        finish_row(o, st, s1, s2);
    }
}

bool tail_bwd_folds_bn3(const NetDims &d) { return !d.bn_train && d.variant == EAV_VARIANT_TOR; }

int launch_tail_bwd(const NetDims &d, const float *dout, const float *probs, const float *params,
                    const float *y3, const float4 *bn3, const uint8_t *mask2, float *dz, float *dz3,
                    float *part, cudaStream_t st) {
    size_t smem = (size_t)(d.FEAT + d.NC) * sizeof(float);
    tail_bwd_kernel<<<d.N, 128, smem, st>>>(dout, probs, params, d.pstride, d.oWd, y3, bn3, mask2, d.B, d.F2,
                                            d.T4, d.T32, d.P2, d.NC, d.variant == EAV_VARIANT_TOR,
                                            d.dropout_mode, d.p_drop, d.seed, d.step, d.step_ptr, dz, dz3, part, tail_bwd_folds_bn3(d) ? 1 : 0);
    EAV_CUDA_LAUNCH_CHECK("tail_bwd");
    return 0;
}

// dWd[m,j,i] = sum_b dz[n,j] * feat[n,i];  dbd[m,j] = sum_b dz[n,j]
__global__ void dense_bwd_w_kernel(const float *__restrict__ feat, const float *__restrict__ dz, int B,
                                   int FEAT, int NC, int64_t pstride, int64_t oWd, int64_t obd,
                                   float *__restrict__ grads) {
    const int m = blockIdx.z, j = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < FEAT) {
        float s = 0.f;
        for (int b = 0; b < B; ++b) {
            int64_t n = (int64_t)m * B + b;
            s = fmaf(dz[n * NC + j], feat[n * FEAT + i], s);
        }
        grads[(int64_t)m * pstride + oWd + (int64_t)j * FEAT + i] = s;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        float s = 0.f;
        for (int b = 0; b < B; ++b) s += dz[((int64_t)m * B + b) * NC + j];
        grads[(int64_t)m * pstride + obd + j] = s;
    }
}

int launch_dense_bwd_w(const NetDims &d, const float *feat, const float *dz, float *grads, cudaStream_t st) {
    dense_bwd_w_kernel<<<dim3(cdiv(d.FEAT, 128), d.NC, d.M), 128, 0, st>>>(feat, dz, d.B, d.FEAT, d.NC, d.pstride,
                                                                          d.oWd, d.obd, grads);
    EAV_CUDA_LAUNCH_CHECK("dense_bwd_w");
    return 0;
}

// =================================================================================
// BatchNorm backward statistics.  part: [M][rows_per_model][ch][2] = (sum dz, sum dz*xhat).
// Writes d(gamma) = sum dz*xhat, d(beta) = sum dz and the {k, c1, c2} block.
// =================================================================================
__global__ void bn_bwd_finalize_kernel(const float *__restrict__ part, int rows_per_model, int ch,
                                       double count, const double *__restrict__ sums, double grad_scale,
                                       const float *__restrict__ params, int64_t pstride,
                                       int64_t og, int64_t ob, const float4 *__restrict__ bnf, int bn_train,
                                       float4 *__restrict__ bnb, float *__restrict__ grads) {
    const int c = blockIdx.x, m = blockIdx.y, lane = threadIdx.x;
    double s1 = 0.0, s2 = 0.0;
    if (sums != nullptr) {              // already reduced (and all-reduced across replicas)
        s1 = sums[((int64_t)m * ch + c) * 2];
        s2 = sums[((int64_t)m * ch + c) * 2 + 1];
    } else {
        const float *p = part + ((int64_t)m * rows_per_model) * ch * 2;
        for (int r = lane; r < rows_per_model; r += 32) {
            s1 += (double)p[((int64_t)r * ch + c) * 2];
            s2 += (double)p[((int64_t)r * ch + c) * 2 + 1];
        }
        s1 = warp_sum(s1);
        s2 = warp_sum(s2);
    }
    if (lane == 0) {
        // grad_scale = 1/dp_world when the sums are global: the caller's gradient all-reduce (sum)
        // then restores exactly the global d(gamma), d(beta)
        grads[(int64_t)m * pstride + og + c] = (float)(s2 * grad_scale);
        grads[(int64_t)m * pstride + ob + c] = (float)(s1 * grad_scale);
        float k = params[(int64_t)m * pstride + og + c] * bnf[(int64_t)m * ch + c].y;
        float c1 = bn_train ? (float)(s1 / count) : 0.f;
        float c2 = bn_train ? (float)(s2 / count) : 0.f;
        bnb[(int64_t)m * ch + c] = make_float4(k, c1, c2, 0.f);
    }
}

int launch_bn_bwd_finalize(const NetDims &d, int layer, const float *part, int rows_per_model,
                           double count, const double *sums, const float *params, const float4 *bnf, float4 *bnb,
                           float *grads, cudaStream_t st) {
    int ch;
    int64_t og, ob;
    if (layer == 1) { ch = d.F1; og = d.og1; ob = d.ob1; }
    else if (layer == 2) { ch = d.G; og = d.og2; ob = d.ob2; }
    else { ch = d.F2; og = d.og3; ob = d.ob3; }
    bn_bwd_finalize_kernel<<<dim3(ch, d.M), 32, 0, st>>>(part, rows_per_model, ch, count, sums,
                                                         sums != nullptr ? 1.0 / d.dp_world : 1.0, params, d.pstride, og,
                                                         ob, bnf, d.bn_train, bnb, grads);
    EAV_CUDA_LAUNCH_CHECK("bn_bwd_finalize");
    return 0;
}

// =================================================================================
// M6 weight gradient:  dW3[m,o,g,k] = sum_{b,u} dy3[n,o,u] * d1[n,g,u+k-7]
// GEMM view M=F2, N=G*16, K=B*T4.  CTA = (model, 16 input channels, sample split);
// thread = 8 o x 8 k, walking u four at a time (8 broadcast LDS.128 + 3 LDS.128 per
// 256 FFMA).  Split partials are reduced by reduce_partials in a fixed order.
// =================================================================================
constexpr int SW_GC = 16;       // input channels per CTA
constexpr int SW_UP = 128;      // padded positions (T4 <= 128 per pass)
constexpr int SW_XS = 148;      // d1 row: 7 left pad + 128 + right pad (>= 128+15+4)

__device__ __forceinline__ void cp_async4_ca(void *smem_dst, const void *gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}

// dy3 (the gradient w.r.t. the conv OUTPUT, BatchNorm-3 backward already applied by
// bn_bwd_apply) and d1 of sample b+1 stream into the second half of a double buffer with
// cp.async while sample b feeds the FFMA loop.
__global__ void __launch_bounds__(256, 2)
sepconv_bwd_dw_kernel(const float *__restrict__ dy3, const float *__restrict__ d1, int B, int F2, int G, int L,
                      int padl, int splits, float *__restrict__ part) {
    extern __shared__ __align__(16) float smem[];
    const int stage_floats = F2 * SW_UP + SW_GC * SW_XS;
    const int m = blockIdx.z, split = blockIdx.y, g0 = blockIdx.x * SW_GC;
    const int tid = threadIdx.x;
    const int kq = tid & 1, gl = (tid >> 1) & (SW_GC - 1), og = tid >> 5;  // og: group of 8 output channels
    const int b_lo = (int)((int64_t)B * split / splits), b_hi = (int)((int64_t)B * (split + 1) / splits);
    const int n_og = F2 / 8;  // output-channel groups; threads with og >= n_og idle (F2 <= 64)
    const int n_ub = (L + SW_UP - 1) / SW_UP;

    auto stage = [&](int step, float *buf) {          // step = (b - b_lo) * n_ub + u-block
        const int b = b_lo + step / n_ub, u_base = (step % n_ub) * SW_UP;
        const int64_t n = (int64_t)m * B + b;
        float *dys = buf, *xs = buf + F2 * SW_UP;
        for (int i = tid; i < F2 * SW_UP; i += 256) {
            const int o = i / SW_UP, j = i - o * SW_UP;
            const int u = u_base + j;
            if (u < L) cp_async4_ca(dys + i, dy3 + (n * F2 + o) * (int64_t)L + u);
            else dys[i] = 0.f;
        }
        for (int i = tid; i < SW_GC * SW_XS; i += 256) {
            const int gg = i / SW_XS, j = i - gg * SW_XS;
            const int u = u_base - padl + j;
            if (g0 + gg < G && u >= 0 && u < L) cp_async4_ca(xs + i, d1 + (n * G + g0 + gg) * (int64_t)L + u);
            else xs[i] = 0.f;
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    float acc[8][8];
#pragma unroll
    for (int oo = 0; oo < 8; ++oo)
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) acc[oo][kk] = 0.f;

    const int n_steps = (b_hi - b_lo) * n_ub;
    if (n_steps > 0) stage(0, smem);
    for (int step = 0; step < n_steps; ++step) {
        float *buf = smem + (step & 1) * stage_floats;
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();                               // tile `step` visible; the other buffer is free
        if (step + 1 < n_steps) stage(step + 1, smem + ((step + 1) & 1) * stage_floats);
        if (og < n_og) {
            const float *dr = buf + og * 8 * SW_UP;
            const float *xr = buf + F2 * SW_UP + gl * SW_XS + 8 * kq;
#pragma unroll 2
            for (int u = 0; u < SW_UP; u += 4) {
                float xw[12];
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                    float4 v = *reinterpret_cast<const float4 *>(xr + u + 4 * q);
                    xw[4 * q] = v.x; xw[4 * q + 1] = v.y; xw[4 * q + 2] = v.z; xw[4 * q + 3] = v.w;
                }
#pragma unroll
                for (int oo = 0; oo < 8; ++oo) {
                    float4 d4 = *reinterpret_cast<const float4 *>(dr + oo * SW_UP + u);
                    const float dv[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
                    for (int uu = 0; uu < 4; ++uu)
#pragma unroll
                        for (int kk = 0; kk < 8; ++kk)
                            acc[oo][kk] = fmaf(dv[uu], xw[uu + kk], acc[oo][kk]);
                }
            }
        }
    }
    if (og < n_og && g0 + gl < G) {
        // part[m][split][o][g][k]
        float *dst = part + (((int64_t)m * splits + split) * F2) * (int64_t)G * 16;
#pragma unroll
        for (int oo = 0; oo < 8; ++oo) {
            float *p = dst + (((int64_t)(og * 8 + oo)) * G + g0 + gl) * 16 + 8 * kq;
            reinterpret_cast<float4 *>(p)[0] = make_float4(acc[oo][0], acc[oo][1], acc[oo][2], acc[oo][3]);
            reinterpret_cast<float4 *>(p)[1] = make_float4(acc[oo][4], acc[oo][5], acc[oo][6], acc[oo][7]);
        }
    }
}

// In place: dz (gradient w.r.t. the BatchNorm OUTPUT) -> gradient w.r.t. the BatchNorm INPUT,
//   dy = k * (dz - c1 - xhat * c2)   (c1 = c2 = 0 in eval mode).  Tiny (43 MB at 1344 samples); it lets
// the two block-2 conv gradient kernels stage their operand with plain asynchronous copies.
// One warp per (sample, channel) row: the BN constants are loaded once per row, no per-element index division.
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(float *__restrict__ dz, const float *__restrict__ y,
                                                           const float4 *__restrict__ bnf,
                                                           const float4 *__restrict__ bnb, int bn_train, int B, int ch,
                                                           int L, int64_t rows) {
    const int lane = threadIdx.x & 31;
    const int64_t row = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int c = (int)(row % ch);
    const int64_t n = row / ch;
    const float4 kb = bnb[(n / B) * ch + c];
    const float4 kf = bnf[(n / B) * ch + c];
    float *d = dz + row * L;
    const float *yr = y + row * L;
    const float a = kb.x, b = kb.x * kf.y * kb.z;          // dy = a (dz - c1) - b (y - mean)
#pragma unroll 4
    for (int u = lane; u < L; u += 32) {
        const float v = d[u];
        d[u] = bn_train ? fmaf(-b, yr[u] - kf.x, a * (v - kb.y)) : a * v;
    }
}

int launch_bn_bwd_apply(const NetDims &d, float *dz3, const float *y3, const float4 *bnf3, const float4 *bnb3,
                        cudaStream_t st) {
    const int64_t rows = (int64_t)d.N * d.F2;
    bn_bwd_apply_kernel<<<(unsigned)cdiv64(rows, 8), 256, 0, st>>>(dz3, y3, bnf3, bnb3, d.bn_train, d.B, d.F2, d.T4, rows);
    EAV_CUDA_LAUNCH_CHECK("bn_bwd_apply");
    return 0;
}

int sepconv_dw_splits(const NetDims &d) {
    if (sepconv_use_tc(d)) return sepconv_dw_tc_splits(d);
    // Split the batch into `s` groups: CTAs = M * ceil(G/16) * s, each walking ceil(B/s) samples.
    const int per_model = cdiv(d.G, SW_GC);
    const int slots = 148 * 2;
    static int forced = -1;
    if (forced < 0) { const char *e = getenv("EAV_SEPDW_SPLITS"); forced = e ? atoi(e) : 0; }
    if (forced > 0) return forced < d.B ? forced : d.B;
    // >= ~4 waves of CTAs (measured on B200: 8 splits 0.643 ms, 4: 0.651, 6: 0.665, 2: 0.81 at M=42, B=32)
    int s = 1;
    while (s < d.B && (int64_t)d.M * per_model * s < 4 * slots) s *= 2;
    return s < d.B ? s : d.B;
}

int launch_sepconv_bwd_dw(const NetDims &d, const float *dz3, const float *y3, const float4 *bnf3,
                          const float4 *bnb3, const float *d1, float *part, float *grads, cudaStream_t st) {
    EAV_REQUIRE(d.F2 % 8 == 0 && d.F2 <= 64, EAV_ERR_UNSUPPORTED, "sepconv_dw: F2=%d unsupported (multiple of 8, <= 64)", d.F2);
    EAV_REQUIRE(d.K2 == 16, EAV_ERR_UNSUPPORTED, "sepconv_dw: kernel length %d unsupported", d.K2);
    const int splits = sepconv_dw_splits(d);
    if (sepconv_use_tc(d))      // tcgen05 path (sepconv_tc.cu); dz3 already holds dy3 (bn_bwd_apply stage)
        return launch_sepconv_dw_tc(d, dz3, d1, part, grads, splits, st);
    size_t smem = (size_t)2 * (d.F2 * SW_UP + SW_GC * SW_XS) * sizeof(float);
    static PerDevice<bool> attr_set_pd(false);
    bool &attr_set = attr_set_pd.here();
    if (!attr_set) {
        cudaFuncSetAttribute(sepconv_bwd_dw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        attr_set = true;
    }
    dim3 grid(cdiv(d.G, SW_GC), splits, d.M);
    (void)y3; (void)bnf3; (void)bnb3;      // dz3 already holds dy3 (bn_bwd_apply stage)
    sepconv_bwd_dw_kernel<<<grid, 256, smem, st>>>(dz3, d1, d.B, d.F2, d.G, d.T4, d.pad2l, splits, part);
    EAV_CUDA_LAUNCH_CHECK("sepconv_bwd_dw");
    return launch_reduce_partials(part, splits, (int64_t)d.F2 * d.G * 16, d.M, d.pstride, grads + d.oW3, st);
}

// dst[m][i] = sum_j part[m][j][i]  (fixed order)
__global__ void reduce_partials_kernel(const float *__restrict__ part, int n_part, int64_t len,
                                       int64_t dst_stride, float *__restrict__ dst) {
    const int m = blockIdx.y;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < len; i += (int64_t)gridDim.x * blockDim.x) {
        const float *p = part + (int64_t)m * n_part * len + i;
        float s = 0.f;
        for (int j = 0; j < n_part; ++j) s += p[(int64_t)j * len];
        dst[(int64_t)m * dst_stride + i] = s;
    }
}

int launch_reduce_partials(const float *part, int n_part, int64_t len, int n_models, int64_t dst_stride,
                           float *dst, cudaStream_t st) {
    int bx = (int)std::min<int64_t>(cdiv64(len, 256), 1024);
    reduce_partials_kernel<<<dim3(bx, n_models), 256, 0, st>>>(part, n_part, len, dst_stride, dst);
    EAV_CUDA_LAUNCH_CHECK("reduce_partials");
    return 0;
}

// =================================================================================
// Variant 1 block-2 backward: pointwise conv then depthwise temporal conv.
// =================================================================================
// dy3d[n,g,u] = sum_o W3p[o,g] * dy3[n,o,u];   per-sample partial dW3p[n][o][g] = sum_u dy3[n,o,u]*y3d[n,g,u]
//
// Fast path (F2, G <= 64, T4 <= 128: the EAV shape is 64 x 64 x 125).  One CTA per sample stages dy3 (BatchNorm-3
// backward applied on load), its transpose, y3d and the 16 KB weight matrix in shared memory once; both small GEMMs
// (2 x 0.5 MFLOP per sample) then run from shared memory with register tiles:
//   phase 1  warp = 8 channels g, lane = 4 positions u (acc[8][4]); per o: 2 broadcast LDS.128 + 4 LDS per 32 FFMA
//   phase 2  thread = (channel g, 16 outputs o) (acc[16]); per u: 1 LDS + 4 broadcast LDS.128 per 16 FFMA
// Row pitch 129 keeps the lane-varying reads conflict free.  (The first version read y3d straight from global memory
// with a stride of T4 floats between lanes and re-staged nothing: 2.36 ms per step at 1344 samples, now ~0.1 ms.)
constexpr int PWB_P = 129, PWB_W = 64, PWB_THREADS = 256;

__global__ void __launch_bounds__(PWB_THREADS)
pw_bwd_tiled_kernel(const float *__restrict__ dz3, const float *__restrict__ y3, const float4 *__restrict__ bnf,
                    const float4 *__restrict__ bnb, int bn_train, const float *__restrict__ y3d,
                    const float *__restrict__ params, int64_t pstride, int64_t oW3p, int B, int G, int F2, int L,
                    float *__restrict__ dy3d, float *__restrict__ part) {
    extern __shared__ __align__(16) float sm[];
    float *dys = sm;                         // [64][PWB_P]   dy3, rows >= F2 / columns >= L zero
    float *as = dys + 64 * PWB_P;            // [64][PWB_P]   y3d
    float *dyT = as + 64 * PWB_P;            // [128][PWB_W]  dy3 transposed
    float *wsh = dyT + 128 * PWB_W;          // [64][PWB_W]   W3p[o][g]
    const int n = blockIdx.x, m = n / B, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 64 * PWB_P; i += PWB_THREADS) { dys[i] = 0.f; as[i] = 0.f; }
    for (int i = tid; i < 128 * PWB_W; i += PWB_THREADS) dyT[i] = 0.f;
    for (int i = tid; i < 64 * PWB_W; i += PWB_THREADS) wsh[i] = 0.f;
    __syncthreads();
    for (int i = tid; i < F2 * L; i += PWB_THREADS) {
        const int o = i / L, u = i - o * L;
        const int64_t idx = (int64_t)n * F2 * L + i;
        float v = dz3[idx];
        const float4 kb = bnb[(int64_t)m * F2 + o];
        if (bn_train) {
            const float4 kf = bnf[(int64_t)m * F2 + o];
            v = kb.x * (v - kb.y - (y3[idx] - kf.x) * kf.y * kb.z);
        } else {
            v = kb.x * v;
        }
        dys[o * PWB_P + u] = v;
        dyT[u * PWB_W + o] = v;
    }
    for (int i = tid; i < G * L; i += PWB_THREADS) {
        const int g = i / L, u = i - g * L;
        as[g * PWB_P + u] = y3d[(int64_t)n * G * L + i];
    }
    const float *W = params + (int64_t)m * pstride + oW3p;
    for (int i = tid; i < F2 * G; i += PWB_THREADS) wsh[(i / G) * PWB_W + (i % G)] = W[i];
    __syncthreads();
    {   // phase 1: dy3d[g][u] = sum_o W[o][g] * dy[o][u]
        float acc[8][4];
#pragma unroll
        for (int a = 0; a < 8; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
        const int g0 = 8 * warp;
        for (int o = 0; o < F2; ++o) {
            const float4 w0 = *reinterpret_cast<const float4 *>(wsh + o * PWB_W + g0);
            const float4 w1 = *reinterpret_cast<const float4 *>(wsh + o * PWB_W + g0 + 4);
            const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
            float dv[4];
#pragma unroll
            for (int b = 0; b < 4; ++b) dv[b] = dys[o * PWB_P + lane + 32 * b];
#pragma unroll
            for (int a = 0; a < 8; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(w[a], dv[b], acc[a][b]);
        }
#pragma unroll
        for (int a = 0; a < 8; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const int g = g0 + a, u = lane + 32 * b;
                if (g < G && u < L) dy3d[((int64_t)n * G + g) * L + u] = acc[a][b];
            }
    }
    {   // phase 2: part[n][o][g] = sum_u dy[o][u] * y3d[g][u]
        const int g = tid & 63, o0 = (tid >> 6) * 16;
        float acc[16];
#pragma unroll
        for (int a = 0; a < 16; ++a) acc[a] = 0.f;
        for (int u = 0; u < L; ++u) {
            const float av = as[g * PWB_P + u];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float4 d4 = *reinterpret_cast<const float4 *>(dyT + u * PWB_W + o0 + 4 * q);
                acc[4 * q] = fmaf(d4.x, av, acc[4 * q]);
                acc[4 * q + 1] = fmaf(d4.y, av, acc[4 * q + 1]);
                acc[4 * q + 2] = fmaf(d4.z, av, acc[4 * q + 2]);
                acc[4 * q + 3] = fmaf(d4.w, av, acc[4 * q + 3]);
            }
        }
        if (g < G) {
#pragma unroll
            for (int a = 0; a < 16; ++a)
                if (o0 + a < F2) part[((int64_t)n * F2 + o0 + a) * G + g] = acc[a];
        }
    }
}

__global__ void __launch_bounds__(128)
pw_bwd_kernel(const float *__restrict__ dz3, const float *__restrict__ y3, const float4 *__restrict__ bnf,
              const float4 *__restrict__ bnb, int bn_train, const float *__restrict__ y3d,
              const float *__restrict__ params, int64_t pstride, int64_t oW3p, int B, int G, int F2, int L,
              float *__restrict__ dy3d, float *__restrict__ part) {
    extern __shared__ float sm[];   // dys[F2][L] + w[F2][G]
    float *dys = sm, *wsh = sm + F2 * L;
    const int n = blockIdx.x, m = n / B, tid = threadIdx.x;
    for (int i = tid; i < F2 * L; i += blockDim.x) {
        int o = i / L;
        int64_t idx = (int64_t)n * F2 * L + i;
        float v = dz3[idx];
        const float4 kb = bnb[(int64_t)m * F2 + o];
        if (bn_train) {
            const float4 kf = bnf[(int64_t)m * F2 + o];
            v = kb.x * (v - kb.y - (y3[idx] - kf.x) * kf.y * kb.z);
        } else {
            v = kb.x * v;
        }
        dys[i] = v;
    }
    const float *W = params + (int64_t)m * pstride + oW3p;
    for (int i = tid; i < F2 * G; i += blockDim.x) wsh[i] = W[i];
    __syncthreads();
    for (int i = tid; i < G * L; i += blockDim.x) {
        int g = i / L, u = i - g * L;
        float s = 0.f;
        for (int o = 0; o < F2; ++o) s = fmaf(wsh[o * G + g], dys[o * L + u], s);
        dy3d[(int64_t)n * G * L + i] = s;
    }
    for (int i = tid; i < F2 * G; i += blockDim.x) {
        int o = i / G, g = i - o * G;
        const float *a = y3d + ((int64_t)n * G + g) * L;
        float s = 0.f;
        for (int u = 0; u < L; ++u) s = fmaf(dys[o * L + u], a[u], s);
        part[(int64_t)n * F2 * G + i] = s;
    }
}

int launch_pw_bwd(const NetDims &d, const float *dz3, const float *y3, const float4 *bnf3,
                  const float4 *bnb3, const float *y3d, const float *params, float *dy3d, float *part,
                  float *grads, cudaStream_t st) {
    size_t smem = (size_t)(d.F2 * d.T4 + d.F2 * d.G) * sizeof(float);
    EAV_REQUIRE(smem <= 200 * 1024, EAV_ERR_UNSUPPORTED, "pw_bwd: F2*T4 too large");
    static PerDevice<bool> attr_set_pd(false);
    bool &attr_set = attr_set_pd.here();
    if (!attr_set) {
        cudaFuncSetAttribute(pw_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        attr_set = true;
    }
    if (d.F2 <= 64 && d.G <= 64 && d.T4 <= 128) {
        const size_t tsm = (size_t)(2 * 64 * PWB_P + 128 * PWB_W + 64 * PWB_W) * sizeof(float);
        static PerDevice<bool> tiled_attr_pd(false);
        bool &tiled_attr = tiled_attr_pd.here();
        if (!tiled_attr) {
            cudaFuncSetAttribute(pw_bwd_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tsm);
            tiled_attr = true;
        }
        pw_bwd_tiled_kernel<<<d.N, PWB_THREADS, tsm, st>>>(dz3, y3, bnf3, bnb3, d.bn_train, y3d, params, d.pstride,
                                                          d.oW3p, d.B, d.G, d.F2, d.T4, dy3d, part);
    } else {
        pw_bwd_kernel<<<d.N, 128, smem, st>>>(dz3, y3, bnf3, bnb3, d.bn_train, y3d, params, d.pstride, d.oW3p, d.B,
                                              d.G, d.F2, d.T4, dy3d, part);
    }
    EAV_CUDA_LAUNCH_CHECK("pw_bwd");
    return launch_reduce_partials(part, d.B, (int64_t)d.F2 * d.G, d.M, d.pstride, grads + d.oW3p, st);
}

// dd1[n,g,u'] = sum_k W3d[g,k] * dy3d[n,g,u'-k+padl];  per-sample dW3d[n][g][k] = sum_u dy3d[n,g,u]*d1[n,g,u+k-padl]
__global__ void __launch_bounds__(128)
dwt_bwd_kernel(const float *__restrict__ dy3d, const float *__restrict__ d1, const float *__restrict__ params,
               int64_t pstride, int64_t oW3, int B, int G, int L, int K2, int padl, float *__restrict__ dd1,
               float *__restrict__ part) {
    const int g = blockIdx.x, n = blockIdx.y, m = n / B;
    const float *w = params + (int64_t)m * pstride + oW3 + (int64_t)g * K2;
    const float *dy = dy3d + ((int64_t)n * G + g) * L;
    const float *x = d1 + ((int64_t)n * G + g) * L;
    for (int u = threadIdx.x; u < L; u += blockDim.x) {
        float s = 0.f;
        for (int k = 0; k < K2; ++k) {
            int uo = u - k + padl;
            if (uo >= 0 && uo < L) s = fmaf(w[k], dy[uo], s);
        }
        dd1[((int64_t)n * G + g) * L + u] = s;
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int k = warp; k < K2; k += blockDim.x >> 5) {
        float s = 0.f;
        for (int u = lane; u < L; u += 32) {
            int ui = u + k - padl;
            if (ui >= 0 && ui < L) s = fmaf(dy[u], x[ui], s);
        }
        s = warp_sum(s);
        if (lane == 0) part[((int64_t)n * G + g) * K2 + k] = s;
    }
}

int launch_dwt_bwd(const NetDims &d, const float *dy3d, const float *d1, const float *params, float *dd1,
                   float *part, float *grads, cudaStream_t st) {
    dwt_bwd_kernel<<<dim3(d.G, d.N), 128, 0, st>>>(dy3d, d1, params, d.pstride, d.oW3, d.B, d.G, d.T4, d.K2,
                                                   d.pad2l, dd1, part);
    EAV_CUDA_LAUNCH_CHECK("dwt_bwd");
    return launch_reduce_partials(part, d.B, (int64_t)d.G * d.K2, d.M, d.pstride, grads + d.oW3, st);
}

// =================================================================================
// M5 backward: dropout, AvgPool(1,P1), ELU' -> dz2 + BN2 partials.  One warp per (n,g) row.
// =================================================================================
__global__ void __launch_bounds__(256)
pool1_bwd_kernel(const float *__restrict__ dd1, const float *__restrict__ y2, const float4 *__restrict__ bnf2,
                 const uint8_t *__restrict__ mask1, int B, int G, int T, int T4, int P1, int dropout_mode,
                 float p_drop, uint64_t seed, uint64_t step, const unsigned long long *__restrict__ step_ptr, int64_t rows, float *__restrict__ dz2,
                 float *__restrict__ part) {
    const int lane = threadIdx.x & 31;
    const int64_t row = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    if (step_ptr) step = *step_ptr;
    const int g = (int)(row % G), n = (int)(row / G);
    const float4 st = bnf2[(int64_t)(n / B) * G + g];
    const float inv_keep = (dropout_mode != EAV_DROPOUT_NONE && p_drop < 1.f) ? 1.f / (1.f - p_drop) : 1.f;
    const float sc = inv_keep / (float)P1;
    float s1 = 0.f, s2 = 0.f;
    auto upstream = [&](int u) -> float {       // d(loss)/d(pooled activation) after dropout, per input sample
        if (u >= T4) return 0.f;
        const int64_t e = row * T4 + u;
        bool keep = true;
        if (dropout_mode == EAV_DROPOUT_MASK) keep = mask1[e] != 0;
        else if (dropout_mode == EAV_DROPOUT_PHILOX) keep = philox_keep(seed, step, 1u, (uint64_t)e, p_drop);
        else if (dropout_mode == EAV_DROPOUT_PHILOX_2D) keep = philox_keep(seed, step, 17u, (uint64_t)row, p_drop);
        return keep ? dd1[e] * sc : 0.f;
    };
    if (P1 == 4 && (T & 3) == 0) {              // one pooling window == one aligned float4
        // lane = block of four consecutive pooled elements = one Philox counter (see pool1_fwd_kernel)
        const int64_t e0 = row * T4, eb0 = e0 & ~(int64_t)3;
        const int nblk = (int)((e0 - eb0 + T4 + 3) >> 2);
        for (int blk = lane; blk < nblk; blk += 32) {
            const int64_t eb = eb0 + 4 * (int64_t)blk;
            float4 yv[4];
            float up[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int64_t u = eb + j - e0;
                const bool ok = u >= 0 && u < T4;
                yv[j] = ok ? *reinterpret_cast<const float4 *>(y2 + row * T + 4 * u) : make_float4(0.f, 0.f, 0.f, 0.f);
                up[j] = ok ? dd1[eb + j] * sc : 0.f;
            }
            uint32_t keep = 0xFu;
            if (dropout_mode == EAV_DROPOUT_PHILOX) keep = philox_keep4(seed, step, 1u, (uint64_t)(eb >> 2), p_drop);
            else if (dropout_mode == EAV_DROPOUT_PHILOX_2D) keep = philox_keep(seed, step, 17u, (uint64_t)row, p_drop) ? 0xFu : 0u;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int64_t u = eb + j - e0;
                if (u < 0 || u >= T4) continue;
                bool kp = ((keep >> j) & 1u) != 0;
                if (dropout_mode == EAV_DROPOUT_MASK) kp = mask1[eb + j] != 0;
                const float upj = kp ? up[j] : 0.f;
                const float4 y = yv[j];
                float4 o;
                o.x = upj * elu_grad_from_pre(fmaf(y.x, st.z, st.w));
                o.y = upj * elu_grad_from_pre(fmaf(y.y, st.z, st.w));
                o.z = upj * elu_grad_from_pre(fmaf(y.z, st.z, st.w));
                o.w = upj * elu_grad_from_pre(fmaf(y.w, st.z, st.w));
                *reinterpret_cast<float4 *>(dz2 + row * T + 4 * u) = o;
                s1 += (o.x + o.y) + (o.z + o.w);
                s2 = fmaf(o.x, (y.x - st.x) * st.y, s2);
                s2 = fmaf(o.y, (y.y - st.x) * st.y, s2);
                s2 = fmaf(o.z, (y.z - st.x) * st.y, s2);
                s2 = fmaf(o.w, (y.w - st.x) * st.y, s2);
            }
        }
    } else {
        for (int t = lane; t < T; t += 32) {
            const float y = y2[row * T + t];
            const float gval = upstream(t / P1) * elu_grad_from_pre(fmaf(y, st.z, st.w));
            dz2[row * T + t] = gval;
            s1 += gval;
            s2 = fmaf(gval, (y - st.x) * st.y, s2);
        }
    }
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    if (lane == 0) { part[row * 2] = s1; part[row * 2 + 1] = s2; }
}

int launch_pool1_bwd(const NetDims &d, const float *dd1, const float *y2, const float4 *bnf2,
                     const uint8_t *mask1, float *dz2, float *part, cudaStream_t st) {
    int64_t rows = (int64_t)d.N * d.G;
    pool1_bwd_kernel<<<(unsigned)cdiv64(rows, 8), 256, 0, st>>>(dd1, y2, bnf2, mask1, d.B, d.G, d.T, d.T4, d.P1,
                                                               d.dropout_mode, d.p_drop, d.seed, d.step, d.step_ptr, rows,
                                                               dz2, part);
    EAV_CUDA_LAUNCH_CHECK("pool1_bwd");
    return 0;
}

// =================================================================================
// M4 backward (+ BN2 bwd on load, ELU1', BN1 partials).  CTA = (sample n, temporal filter f):
//   dW2[fD+d, c] += sum_t dy2[n,fD+d,t] * a1[n,f,c,t]          (per-CTA partial, reduced later)
//   dz1[n,f,c,t]  = act'(.) * sum_d W2new[fD+d,c] * dy2[n,fD+d,t]
// a1 = act(BN1(y1)) is recomputed from the saved raw conv output.
// =================================================================================
constexpr int DB_THREADS = 256;
constexpr int DB_TS = 16;   // t-slices for the dW2 reduction

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

// smem layout (floats): y1s [C][TSa] | dys [8][TSa] | w2s [8][C] | red [max(DB_TS*D*C, 8*TSa)];  TSa = roundup4(T) + 4
// The raw conv output y1, dz2 (and y2 in train mode, parked in `red`) are staged with 16-byte
// cp.async (LDGSTS): every load of the CTA is in flight at once and no register is tied up.
// a1 = act(BN1(y1)) and act' are recomputed from the staged y1, so y1 is read from HBM once.
// CT / TT / DT: compile-time Chans / Samples / D (0 = use the runtime value).  The dataset shape
// (30, 500, 8) gets its own instantiation: with constant trip counts the staging and phase loops
// unroll and the address arithmetic folds (the generic build spent ~40 % of its issue slots on
// integer/branch instructions).
template <int CT, int TT, int DT>
__global__ void __launch_bounds__(DB_THREADS, 2)
dw_bwd_kernel(const float *__restrict__ dz2, const float *__restrict__ y2, const float4 *__restrict__ bnf2,
              const float4 *__restrict__ bnb2, const float *__restrict__ y1, const float4 *__restrict__ bnf1,
              const float *__restrict__ params, int64_t pstride, int64_t oW2, int bn_train, int elu1, int B,
              int F1, int D_rt, int C_rt, int T_rt, float *__restrict__ dz1, float *__restrict__ part_w,
              float *__restrict__ part_bn) {
    extern __shared__ __align__(16) float smem[];
    const int C = CT ? CT : C_rt, T = TT ? TT : T_rt, D = DT ? DT : D_rt;
    const int TSa = ((T + 3) & ~3) + 4;
    float *y1s = smem;                    // [C][TSa]   raw conv output, zero padded past T
    float *dys = y1s + C * TSa;           // [8][TSa]   dL/d(depthwise output) (BN2 backward applied), rows >= D zero
    float *w2s = dys + 8 * TSa;           // [8][C]
    float *red = w2s + 8 * C;             // [DB_TS][D*C]  (first holds the staged y2 rows in train mode)
    const int n = blockIdx.x / F1, f = blockIdx.x - n * F1, m = n / B;
    const int G = F1 * D, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr int NW = DB_THREADS / 32;
    const float4 s1 = bnf1[(int64_t)m * F1 + f];
    const bool vec = (T & 3) == 0;
    const float *y1r = y1 + (((int64_t)n * F1 + f) * C) * (int64_t)T;

    // ---- staging: one warp per row, lanes along time
    for (int r = warp; r < C + 16; r += NW) {
        const float *src;
        float *dst;
        if (r < C) { src = y1r + (int64_t)r * T; dst = y1s + r * TSa; }
        else if (r < C + 8) {
            const int dd = r - C;
            dst = dys + dd * TSa;
            src = (dd < D) ? dz2 + ((int64_t)n * G + f * D + dd) * T : nullptr;
        } else {
            const int dd = r - C - 8;
            dst = red + dd * TSa;
            src = (bn_train && dd < D) ? y2 + ((int64_t)n * G + f * D + dd) * T : nullptr;
            if (src == nullptr) continue;
        }
        if (src == nullptr) {
            for (int t = lane; t < TSa; t += 32) dst[t] = 0.f;
            continue;
        }
        if (vec) for (int t = 4 * lane; t < T; t += 128) cp_async16(dst + t, src + t);
        else for (int t = lane; t < T; t += 32) dst[t] = src[t];
        for (int t = T + lane; t < TSa; t += 32) dst[t] = 0.f;
    }
    const float *W2 = params + (int64_t)m * pstride + oW2 + (int64_t)f * D * C;
    for (int i = tid; i < 8 * C; i += DB_THREADS) w2s[i] = (i < D * C) ? W2[i] : 0.f;
    cp_async_wait_all();
    __syncthreads();
    // BatchNorm-2 backward in place on the staged rows: dy2 = k (dz2 - c1 - xhat2 c2)
    for (int dd = warp; dd < D; dd += NW) {
        const float4 kb = bnb2[(int64_t)m * G + f * D + dd];
        const float4 kf = bnf2[(int64_t)m * G + f * D + dd];
        float *row = dys + dd * TSa;
        const float *yrow = red + dd * TSa;
        for (int t = lane; t < T; t += 32) {
            float v = row[t];
            if (bn_train) v = kb.x * (v - kb.y - (yrow[t] - kf.x) * kf.y * kb.z);
            else v = kb.x * v;
            row[t] = v;
        }
    }
    __syncthreads();

    // ---- phase 1: dW2 partial.  work item = (channel pair, t-slice); four time steps per smem load.
    {
        const int n_cp = (C + 1) / 2;
        const int per = ((((T + 3) >> 2) + DB_TS - 1) / DB_TS) << 2;      // slice length, multiple of 4
        for (int item = tid; item < n_cp * DB_TS; item += DB_THREADS) {
            const int cp = item % n_cp, ts = item / n_cp;
            float acc[2][8];
#pragma unroll
            for (int q = 0; q < 2; ++q)
#pragma unroll
                for (int dd = 0; dd < 8; ++dd) acc[q][dd] = 0.f;
            const int c0 = 2 * cp, c1 = (2 * cp + 1 < C) ? 2 * cp + 1 : c0;
            const int t_lo = min(TSa - 4, ts * per), t_hi = min(TSa - 4, t_lo + per);
            for (int t = t_lo; t < t_hi; t += 4) {
                float4 a0 = *reinterpret_cast<const float4 *>(y1s + c0 * TSa + t);
                float4 a1v = *reinterpret_cast<const float4 *>(y1s + c1 * TSa + t);
                // a = act(BN1(y1)); the zero pads past T meet zero dy2, so their value is irrelevant
                a0.x = fmaf(a0.x, s1.z, s1.w); a0.y = fmaf(a0.y, s1.z, s1.w); a0.z = fmaf(a0.z, s1.z, s1.w); a0.w = fmaf(a0.w, s1.z, s1.w);
                a1v.x = fmaf(a1v.x, s1.z, s1.w); a1v.y = fmaf(a1v.y, s1.z, s1.w); a1v.z = fmaf(a1v.z, s1.z, s1.w); a1v.w = fmaf(a1v.w, s1.z, s1.w);
                if (elu1) {
                    // exp(x) - 1 straight from MUFU.EX2: its ~1e-7 ABSOLUTE error near zero is irrelevant for a weight
                    // gradient that sums dy * a over 500 samples (the forward pass keeps the polynomial form, whose
                    // relative accuracy matters there); 4 instructions instead of 14 per element.
                    a0.x = elu_bwd_act(a0.x); a0.y = elu_bwd_act(a0.y); a0.z = elu_bwd_act(a0.z); a0.w = elu_bwd_act(a0.w);
                    a1v.x = elu_bwd_act(a1v.x); a1v.y = elu_bwd_act(a1v.y); a1v.z = elu_bwd_act(a1v.z); a1v.w = elu_bwd_act(a1v.w);
                }
#pragma unroll
                for (int dd = 0; dd < 8; ++dd) {
                    const float4 dv = *reinterpret_cast<const float4 *>(dys + dd * TSa + t);
                    acc[0][dd] = fmaf(dv.x, a0.x, fmaf(dv.y, a0.y, fmaf(dv.z, a0.z, fmaf(dv.w, a0.w, acc[0][dd]))));
                    acc[1][dd] = fmaf(dv.x, a1v.x, fmaf(dv.y, a1v.y, fmaf(dv.z, a1v.z, fmaf(dv.w, a1v.w, acc[1][dd]))));
                }
            }
            // `red` still holds the staged y2 rows until every thread is past the transform above
            // (guaranteed by the __syncthreads), so it can be overwritten now.
#pragma unroll
            for (int dd = 0; dd < 8; ++dd) {
                if (dd < D) {
                    red[(ts * D + dd) * C + c0] = acc[0][dd];
                    if (2 * cp + 1 < C) red[(ts * D + dd) * C + c1] = acc[1][dd];
                }
            }
        }
        __syncthreads();
        for (int i = tid; i < D * C; i += DB_THREADS) {
            float s = 0.f;
#pragma unroll
            for (int q = 0; q < DB_TS; ++q) s += red[q * D * C + i];
            part_w[((int64_t)n * F1 + f) * (D * C) + i] = s;   // [n][f][d][c] == [n][g][c]
        }
    }

    // ---- phase 2: dz1 = act'(.) * (W2new^T dy2) and BN1 partial sums.
    float p1 = 0.f, p2 = 0.f;
    if (vec) {
        // thread = (group of 4 time steps, channel range): all 256 threads busy for T = 500
        const int n_t4 = T >> 2;
        const int n_cr = max(1, DB_THREADS / n_t4);            // channel ranges processed concurrently
        const int c_per = (C + n_cr - 1) / n_cr;
        for (int item = tid; item < n_t4 * n_cr; item += DB_THREADS) {
            const int t = (item % n_t4) * 4, cr = item / n_t4;
            float4 dv[8];
#pragma unroll
            for (int dd = 0; dd < 8; ++dd) dv[dd] = *reinterpret_cast<const float4 *>(dys + dd * TSa + t);
            const int c_hi = min(C, (cr + 1) * c_per);
            for (int c = cr * c_per; c < c_hi; ++c) {
                float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int dd = 0; dd < 8; ++dd) {           // rows >= D of w2s are zero
                    const float w = w2s[dd * C + c];
                    s.x = fmaf(w, dv[dd].x, s.x); s.y = fmaf(w, dv[dd].y, s.y);
                    s.z = fmaf(w, dv[dd].z, s.z); s.w = fmaf(w, dv[dd].w, s.w);
                }
                const float4 y = *reinterpret_cast<const float4 *>(y1s + c * TSa + t);
                if (elu1) {
                    s.x *= elu_grad_from_pre(fmaf(y.x, s1.z, s1.w)); s.y *= elu_grad_from_pre(fmaf(y.y, s1.z, s1.w));
                    s.z *= elu_grad_from_pre(fmaf(y.z, s1.z, s1.w)); s.w *= elu_grad_from_pre(fmaf(y.w, s1.z, s1.w));
                }
                *reinterpret_cast<float4 *>(dz1 + (((int64_t)n * F1 + f) * C + c) * (int64_t)T + t) = s;
                p1 += (s.x + s.y) + (s.z + s.w);
                p2 = fmaf(s.x, (y.x - s1.x) * s1.y, p2); p2 = fmaf(s.y, (y.y - s1.x) * s1.y, p2);
                p2 = fmaf(s.z, (y.z - s1.x) * s1.y, p2); p2 = fmaf(s.w, (y.w - s1.x) * s1.y, p2);
            }
        }
    } else {
        for (int t = tid; t < T; t += DB_THREADS) {
            float dv[8];
#pragma unroll
            for (int dd = 0; dd < 8; ++dd) dv[dd] = dys[dd * TSa + t];
            for (int c = 0; c < C; ++c) {
                float s = 0.f;
#pragma unroll
                for (int dd = 0; dd < 8; ++dd) s = fmaf(w2s[dd * C + c], dv[dd], s);
                const float y = y1s[c * TSa + t];
                const float g = s * (elu1 ? elu_grad_from_pre(fmaf(y, s1.z, s1.w)) : 1.f);
                dz1[(((int64_t)n * F1 + f) * C + c) * (int64_t)T + t] = g;
                p1 += g;
                p2 = fmaf(g, (y - s1.x) * s1.y, p2);
            }
        }
    }
    __shared__ float redb[DB_THREADS / 32][2];
    p1 = warp_sum(p1);
    p2 = warp_sum(p2);
    if (lane == 0) { redb[warp][0] = p1; redb[warp][1] = p2; }
    __syncthreads();
    if (tid < 2) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < DB_THREADS / 32; ++w) s += redb[w][tid];
        part_bn[((int64_t)n * F1 + f) * 2 + tid] = s;
    }
}

int launch_dw_bwd(const NetDims &d, const float *dz2, const float *y2, const float4 *bnf2,
                  const float4 *bnb2, const float *y1, const float4 *bnf1, const float *params,
                  float *dz1, float *part_w, float *part_bn, float *grads, cudaStream_t st) {
    EAV_REQUIRE(d.D <= 8, EAV_ERR_UNSUPPORTED, "dw_bwd: D=%d > 8 unsupported", d.D);
    const int TSa = ((d.T + 3) & ~3) + 4;
    size_t redn = (size_t)DB_TS * d.D * d.C;
    if (redn < (size_t)8 * TSa) redn = (size_t)8 * TSa;
    size_t fl = (size_t)d.C * TSa + (size_t)8 * TSa + 8 * d.C + redn;
    size_t smem = fl * sizeof(float);
    EAV_REQUIRE(smem <= 200 * 1024, EAV_ERR_UNSUPPORTED, "dw_bwd: Chans*Samples too large for one CTA");
    static PerDevice<bool> attr_set_pd(false);
    bool &attr_set = attr_set_pd.here();
    if (!attr_set) {
        cudaFuncSetAttribute(dw_bwd_kernel<0, 0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        cudaFuncSetAttribute(dw_bwd_kernel<30, 500, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        attr_set = true;
    }
    if (d.C == 30 && d.T == 500 && d.D == 8)
        dw_bwd_kernel<30, 500, 8><<<d.N * d.F1, DB_THREADS, smem, st>>>(dz2, y2, bnf2, bnb2, y1, bnf1, params, d.pstride, d.oW2,
                                                                       d.bn_train, d.variant == EAV_VARIANT_TOR, d.B, d.F1,
                                                                       d.D, d.C, d.T, dz1, part_w, part_bn);
    else
        dw_bwd_kernel<0, 0, 0><<<d.N * d.F1, DB_THREADS, smem, st>>>(dz2, y2, bnf2, bnb2, y1, bnf1, params, d.pstride, d.oW2,
                                                                    d.bn_train, d.variant == EAV_VARIANT_TOR, d.B, d.F1, d.D,
                                                                    d.C, d.T, dz1, part_w, part_bn);
    EAV_CUDA_LAUNCH_CHECK("dw_bwd");
    // part_w is [N][G*C]: reduce over the B samples of each model
    return launch_reduce_partials(part_w, d.B, (int64_t)d.G * d.C, d.M, d.pstride, grads + d.oW2, st);
}

// =================================================================================
// M1 weight gradient:  dW1[m,f,k] = sum_{b,c,t} dy1[n,f,c,t] * x[n,c,t+k-pad1l]
// (72.0 MFLOP per epoch; the reference spends 54 % of its step here).
// One warp per (n,c) row: lane owns RK consecutive lags k for all F1 filters
// (F1*RK accumulators) and walks t four at a time: 8 broadcast LDS.128 (dy1 as [t][f]) +
// RK/2+2 LDS.64 (x window) per 4*F1*RK FFMA.  A CTA (4 warps) walks a contiguous range
// of rows of ONE model and writes one partial; reduce_partials sums them in order.
// =================================================================================
constexpr int TW_WARPS = 4;
constexpr int TW_TCH = 128;     // time steps of dy1 staged per pass

template <int F1, int RK, int MINB>
__global__ void __launch_bounds__(TW_WARPS * 32, MINB)
tconv_bwd_dw_kernel(const float *__restrict__ x, const int32_t *__restrict__ x_index,
                    const float *__restrict__ dz1, const float *__restrict__ y1,
                    const float4 *__restrict__ bnf1, const float4 *__restrict__ bnb1, int bn_train, int B,
                    int C, int T, int K1, int padl, int ctas_per_model, float *__restrict__ part) {
    extern __shared__ __align__(16) float smem[];
    const int Tp = (T + 3) & ~3;
    const int XS = (Tp + 32 * RK + 8 + 3) & ~3;      // x row incl. both paddings
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int DYS = TW_TCH * F1 + (TW_TCH / 4) * 4;   // [t][F1] rows, +4 floats after every 4 rows (bank spread)
    float *xs = smem + warp * (XS + DYS);            // per-warp private staging
    float *dys = xs + XS;                            // dys[t*F1 + (t/4)*4 + f]
    const int m = blockIdx.x / ctas_per_model, j = blockIdx.x - m * ctas_per_model;
    const int rows = B * C;                          // rows of this model
    const int r_lo = (int)((int64_t)rows * j / ctas_per_model), r_hi = (int)((int64_t)rows * (j + 1) / ctas_per_model);

    // Packed fp32 (FFMA2): float2 accumulators over filter PAIRS; dy1 pairs come out of the [t][f]
    // LDS.128, the x operand is the scalar-broadcast form.
    float2 acc2[F1 / 2][RK];
#pragma unroll
    for (int p = 0; p < F1 / 2; ++p)
#pragma unroll
        for (int q = 0; q < RK; ++q) acc2[p][q] = make_float2(0.f, 0.f);

    const bool vec_ok = (T & 3) == 0;
    for (int r = r_lo + warp; r < r_hi; r += TW_WARPS) {
        const int b = r / C, c = r - b * C;
        const int64_t n = (int64_t)m * B + b;
        const int64_t xrow = x_index ? (int64_t)x_index[n] : n;
        __syncwarp();
        // xs[i] = x[i - padl], zero padded on both sides
        const float *xsrc = x + (xrow * C + c) * (int64_t)T;
        if (vec_ok) {
            for (int i = lane; i < padl; i += 32) xs[i] = 0.f;
            for (int i = padl + T + lane; i < XS; i += 32) xs[i] = 0.f;
#pragma unroll 4
            for (int t = 4 * lane; t < T; t += 128) {
                float4 v = *reinterpret_cast<const float4 *>(xsrc + t);
                float *d = xs + padl + t;
                d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
            }
        } else {
            for (int i = lane; i < XS; i += 32) {
                int t = i - padl;
                xs[i] = (t >= 0 && t < T) ? xsrc[t] : 0.f;
            }
        }
        for (int tc = 0; tc < Tp; tc += TW_TCH) {
            const int tn = min(TW_TCH, Tp - tc);     // multiple of 4
            __syncwarp();
            // dys[f][t] = dL/d(conv output): BatchNorm-1 backward applied while staging.
            if (vec_ok) {
                const int t = 4 * lane;
                const bool act = t < tn;              // T % 4 == 0: the whole float4 is in range
#pragma unroll
                for (int fh = 0; fh < F1; fh += 2) {       // two filters at a time keeps the staging registers low
                    float4 dzv[2], yv[2], o[2];
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        const int64_t base = ((n * F1 + fh + q) * C + c) * (int64_t)T + tc + t;
                        dzv[q] = act ? *reinterpret_cast<const float4 *>(dz1 + base) : make_float4(0.f, 0.f, 0.f, 0.f);
                        if (bn_train) yv[q] = act ? *reinterpret_cast<const float4 *>(y1 + base) : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        const float4 kb = bnb1[(int64_t)m * F1 + fh + q];
                        if (bn_train) {
                            const float4 kf = bnf1[(int64_t)m * F1 + fh + q];
                            o[q].x = kb.x * (dzv[q].x - kb.y - (yv[q].x - kf.x) * kf.y * kb.z);
                            o[q].y = kb.x * (dzv[q].y - kb.y - (yv[q].y - kf.x) * kf.y * kb.z);
                            o[q].z = kb.x * (dzv[q].z - kb.y - (yv[q].z - kf.x) * kf.y * kb.z);
                            o[q].w = kb.x * (dzv[q].w - kb.y - (yv[q].w - kf.x) * kf.y * kb.z);
                        } else {
                            o[q] = make_float4(kb.x * dzv[q].x, kb.x * dzv[q].y, kb.x * dzv[q].z, kb.x * dzv[q].w);
                        }
                    }
                    if (act) {      // register transpose: one float2 over 2 filters per time step
                        float *d = dys + t * F1 + t + fh;          // t*F1 + (t/4)*4 with t % 4 == 0
                        *reinterpret_cast<float2 *>(d) = make_float2(o[0].x, o[1].x);
                        *reinterpret_cast<float2 *>(d + F1) = make_float2(o[0].y, o[1].y);
                        *reinterpret_cast<float2 *>(d + 2 * F1) = make_float2(o[0].z, o[1].z);
                        *reinterpret_cast<float2 *>(d + 3 * F1) = make_float2(o[0].w, o[1].w);
                    }
                }
            } else {
#pragma unroll
                for (int f = 0; f < F1; ++f) {
                    const float4 kb = bnb1[(int64_t)m * F1 + f];
                    const float4 kf = bnf1[(int64_t)m * F1 + f];
                    const int64_t base = ((n * F1 + f) * C + c) * (int64_t)T + tc;
                    for (int t = lane; t < tn; t += 32) {
                        float v = 0.f;
                        if (tc + t < T) {
                            v = dz1[base + t];
                            if (bn_train) v = kb.x * (v - kb.y - (y1[base + t] - kf.x) * kf.y * kb.z);
                            else v = kb.x * v;
                        }
                        dys[t * F1 + (t >> 2) * 4 + f] = v;
                    }
                }
            }
            __syncwarp();
            const float *xr = xs + lane * RK + tc;
#pragma unroll 1
            for (int t = 0; t < tn; t += 4) {
                float xw[RK + 4];
#pragma unroll
                for (int q = 0; q < (RK + 4) / 2; ++q) {
                    float2 v = *reinterpret_cast<const float2 *>(xr + t + 2 * q);
                    xw[2 * q] = v.x; xw[2 * q + 1] = v.y;
                }
                const float *dr = dys + t * F1 + t;        // rows t..t+3 are contiguous (pad comes after them)
#pragma unroll
                for (int tt = 0; tt < 4; ++tt) {
                    float2 dv2[F1 / 2];
#pragma unroll
                    for (int h = 0; h < F1 / 4; ++h) {
                        const float4 v = *reinterpret_cast<const float4 *>(dr + tt * F1 + 4 * h);
                        dv2[2 * h] = make_float2(v.x, v.y);
                        dv2[2 * h + 1] = make_float2(v.z, v.w);
                    }
#pragma unroll
                    for (int q = 0; q < RK; ++q) {
                        const float2 xx = make_float2(xw[tt + q], xw[tt + q]);
#pragma unroll
                        for (int p = 0; p < F1 / 2; ++p) acc2[p][q] = __ffma2_rn(dv2[p], xx, acc2[p][q]);
                    }
                }
            }
        }
    }
    // cross-warp reduction through smem (reuse the staging area), then one partial per CTA
    __syncthreads();
    float *red = smem;   // [TW_WARPS][F1][32*RK]
#pragma unroll
    for (int p = 0; p < F1 / 2; ++p)
#pragma unroll
        for (int q = 0; q < RK; ++q) {
            red[(warp * F1 + 2 * p) * (32 * RK) + lane * RK + q] = acc2[p][q].x;
            red[(warp * F1 + 2 * p + 1) * (32 * RK) + lane * RK + q] = acc2[p][q].y;
        }
    __syncthreads();
    for (int i = threadIdx.x; i < F1 * K1; i += blockDim.x) {
        int f = i / K1, k = i - f * K1;
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < TW_WARPS; ++w) s += red[(w * F1 + f) * (32 * RK) + k];
        part[((int64_t)m * ctas_per_model + j) * (F1 * K1) + i] = s;
    }
}

template <int RK>
static size_t tconv_dw_smem(int T) {
    const int Tp = (T + 3) & ~3;
    const int XS = (Tp + 32 * RK + 8 + 3) & ~3;
    size_t stage = (size_t)TW_WARPS * (XS + TW_TCH * 8 + (TW_TCH / 4) * 4) * sizeof(float);
    size_t redb = (size_t)TW_WARPS * 8 * 32 * RK * sizeof(float);
    return stage > redb ? stage : redb;
}

static int tconv_rk(int K1) { return K1 <= 64 ? 2 : K1 <= 128 ? 4 : K1 <= 256 ? 8 : K1 <= 320 ? 10 : 16; }

int tconv_dw_ctas_per_model(const NetDims &d) {
    if (tconv_bwd_dw_use_tc(d)) return tconv_bwd_dw_tc_splits(d);
    // ~two full waves of resident CTAs (148 SMs x up to 3 CTAs) when the work allows it
    int64_t rows = (int64_t)d.B * d.C;
    static int per_sm = -1;
    if (per_sm < 0) { const char *e = getenv("EAV_TW_MINB"); per_sm = e ? atoi(e) : 3; }
    int64_t want = (148 * per_sm * 2) / d.M;
    int64_t cap = cdiv64(rows, TW_WARPS);   // at least one row per warp
    int64_t c = want < cap ? want : cap;
    if (c < 1) c = 1;
    return (int)c;
}

template <int RK>
static int launch_tconv_bwd_dw_rk(const NetDims &d, const float *x, const int32_t *x_index, const float *dz1,
                                  const float *y1, const float4 *bnf1, const float4 *bnb1, float *part,
                                  cudaStream_t st, int cpm) {
    size_t smem = tconv_dw_smem<RK>(d.T);
    EAV_REQUIRE(smem <= 200 * 1024, EAV_ERR_UNSUPPORTED, "tconv_bwd_dw: Samples=%d too large", d.T);
    static PerDevice<bool> attr_set_pd(false);
    bool &attr_set = attr_set_pd.here();   // one flag per RK instantiation
    if (!attr_set) {
        cudaFuncSetAttribute(tconv_bwd_dw_kernel<8, RK, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        cudaFuncSetAttribute(tconv_bwd_dw_kernel<8, RK, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        attr_set = true;
    }
    static int minb = -1;
    if (minb < 0) { const char *e = getenv("EAV_TW_MINB"); minb = e ? atoi(e) : 3; }
    if (minb == 2)
        tconv_bwd_dw_kernel<8, RK, 2><<<d.M * cpm, TW_WARPS * 32, smem, st>>>(x, x_index, dz1, y1, bnf1, bnb1, d.bn_train,
                                                                              d.B, d.C, d.T, d.K1, d.pad1l, cpm, part);
    else
        tconv_bwd_dw_kernel<8, RK, 3><<<d.M * cpm, TW_WARPS * 32, smem, st>>>(x, x_index, dz1, y1, bnf1, bnb1, d.bn_train,
                                                                              d.B, d.C, d.T, d.K1, d.pad1l, cpm, part);
    return 0;
}

int launch_tconv_bwd_dw(const NetDims &d, const float *x, const int32_t *x_index, const float *dz1,
                        const float *y1, const float4 *bnf1, const float4 *bnb1, float *part, float *grads,
                        cudaStream_t st) {
    EAV_REQUIRE(d.F1 == 8, EAV_ERR_UNSUPPORTED, "tconv_bwd_dw: F1=%d unsupported (only 8)", d.F1);
    EAV_REQUIRE(d.K1 <= 512, EAV_ERR_UNSUPPORTED, "tconv_bwd_dw: kernLength=%d > 512 unsupported", d.K1);
    const int cpm = tconv_dw_ctas_per_model(d);
    if (tconv_bwd_dw_use_tc(d)) {   // tcgen05 path (tconv_tc.cu); EAV_TCONV=ffma selects the CUDA-core kernel below
        int rc = launch_tconv_bwd_dw_tc(d, x, x_index, dz1, y1, bnf1, bnb1, part, cpm, st);
        if (rc) return rc;
        return launch_reduce_partials(part, cpm, (int64_t)d.F1 * d.K1, d.M, d.pstride, grads + d.oW1, st);
    }
    int rc;
    switch (tconv_rk(d.K1)) {
        case 2: rc = launch_tconv_bwd_dw_rk<2>(d, x, x_index, dz1, y1, bnf1, bnb1, part, st, cpm); break;
        case 4: rc = launch_tconv_bwd_dw_rk<4>(d, x, x_index, dz1, y1, bnf1, bnb1, part, st, cpm); break;
        case 8: rc = launch_tconv_bwd_dw_rk<8>(d, x, x_index, dz1, y1, bnf1, bnb1, part, st, cpm); break;
        case 10: rc = launch_tconv_bwd_dw_rk<10>(d, x, x_index, dz1, y1, bnf1, bnb1, part, st, cpm); break;
        default: rc = launch_tconv_bwd_dw_rk<16>(d, x, x_index, dz1, y1, bnf1, bnb1, part, st, cpm); break;
    }
    if (rc) return rc;
    EAV_CUDA_LAUNCH_CHECK("tconv_bwd_dw");
    return launch_reduce_partials(part, cpm, (int64_t)d.F1 * d.K1, d.M, d.pstride, grads + d.oW1, st);
}

}  // namespace eav

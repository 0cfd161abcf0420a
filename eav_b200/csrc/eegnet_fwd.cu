// Forward kernels of the EEGNet path (CNN_torch/EEGNet_tor.py:50-67, CNN_torch/CNN_EEG.py:57-67)
// for sm_100a.  Data layout: activations [N][channels][time] fp32, N = M*B samples of M
// independent models (sample n belongs to model n / B); parameters are read from the
// flat per-model arena (NetDims offsets).
#include "eegnet_kernels.cuh"

namespace eav {

// =================================================================================
// M1  temporal convolution, register-blocked sliding window on the CUDA cores.
//   y1[n,f,c,t] = sum_k W1[f,k] * x[n,c,t+k-pad1l]          (EEGNet_tor.py:24,51)
// CTA = 4 channel rows x 512 time outputs; 64 threads per row, each thread keeps
// F1 x 8 accumulators and walks the taps 4 at a time: per 4 taps it issues 3 LDS.128 for
// the x window, 8 broadcast LDS.128 for the weights and 256 FFMA (95.9 % FFMA issue).
// Algorithmic work: 2*K1*F1 flop per output sample -> 72.0 MFLOP per EEG epoch.
// =================================================================================
constexpr int TC_TT = 512;      // time outputs per work item (threads per row x outputs per thread)

__device__ __forceinline__ void cp_async4(void *smem_dst, const void *gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Persistent CTAs: the grid is sized to the resident capacity of the GPU and every CTA walks a
// CONTIGUOUS range of work items (item = sample n, 3-row channel group, 512-sample time tile),
// so there is no wave-quantisation tail and the filter bank is re-staged only when the model
// changes.  The x tile of the NEXT item is fetched with cp.async into the other half of a
// double buffer while the FFMA loop of the current item runs.
template <int F1, int TC_ROWS, int KU, int MINB, int TC_R>
__global__ void __launch_bounds__(TC_ROWS * (TC_TT / TC_R), MINB)
tconv_fwd_kernel(const float *__restrict__ x, const int32_t *__restrict__ x_index,
                 const float *__restrict__ params, int64_t pstride, int64_t oW1, int B, int C,
                 int T, int K1, int padl, int n_items, int groups, int tiles, float *__restrict__ y1,
                 float *__restrict__ part) {
    extern __shared__ __align__(16) float smem[];
    constexpr int TC_TPR = TC_TT / TC_R;             // threads per row
    constexpr int TC_THREADS = TC_ROWS * TC_TPR;
    const int K1p = (K1 + KU - 1) / KU * KU;
    const int XS = TC_TT + K1p + 4;  // multiple of 4
    float *ws = smem;                // [K1p][F1]
    float *xs0 = smem + K1p * F1;    // 2 x [TC_ROWS][XS]
    __shared__ float red[TC_THREADS / 32][2 * F1];
    const int tid = threadIdx.x;
    const int it_lo = (int)((int64_t)n_items * blockIdx.x / gridDim.x);
    const int it_hi = (int)((int64_t)n_items * (blockIdx.x + 1) / gridDim.x);

    auto decode = [&](int item, int &n, int &c0, int &tile0) {
        const int per_n = groups * tiles;
        n = item / per_n;
        const int rem = item - n * per_n;
        const int g = rem / tiles;
        c0 = g * TC_ROWS;
        tile0 = (rem - g * tiles) * TC_TT;
    };
    auto stage = [&](int item, float *xs) {     // asynchronous: returns with the copies in flight
        int n, c0, tile0;
        decode(item, n, c0, tile0);
        const int64_t xrow = x_index ? (int64_t)x_index[n] : (int64_t)n;
        const int lim = TC_TT + K1p;
        for (int r = 0; r < TC_ROWS; ++r) {
            const int c = c0 + r;
            const float *src = x + (xrow * C + c) * (int64_t)T;
            float *dst = xs + r * XS;
            for (int j = tid; j < XS; j += TC_THREADS) {
                const int t = tile0 - padl + j;
                if (c < C && j < lim && t >= 0 && t < T) cp_async4(dst + j, src + t);
                else dst[j] = 0.f;
            }
        }
        cp_async_commit();
    };

    if (it_lo < it_hi) stage(it_lo, xs0);
    int cur_m = -1;
    for (int item = it_lo; item < it_hi; ++item) {
        float *xs = xs0 + ((item - it_lo) & 1) * (TC_ROWS * XS);
        int n, c0, tile0;
        decode(item, n, c0, tile0);
        const int m = n / B;
        if (m != cur_m) {                       // (re)stage the filter bank of this model
            __syncthreads();                    // nobody still reads the previous bank
            const float *W1 = params + (int64_t)m * pstride + oW1;
            for (int i = tid; i < K1p * F1; i += TC_THREADS) {
                int k = i / F1, f = i - k * F1;
                ws[i] = (k < K1) ? W1[f * K1 + k] : 0.f;
            }
            cur_m = m;
        }
        cp_async_wait<0>();
        __syncthreads();                        // tile `item` and the weights are visible to all
        if (item + 1 < it_hi) stage(item + 1, xs0 + ((item + 1 - it_lo) & 1) * (TC_ROWS * XS));

        const int r = tid / TC_TPR, j = tid - r * TC_TPR;
        const int t0 = j * TC_R;
        const float *xr = xs + r * XS + t0;
        // Packed fp32 (Blackwell FFMA2, fma.rn.f32x2): accumulators are float2 over filter PAIRS, the
        // weight pairs (w[2p], w[2p+1]) come straight out of the [k][f] LDS.128, the x operand is a
        // duplicated pair.  Two FMAs per instruction halve the register-file reads per FMA: an 8x8
        // register outer product measures 63.3 TFLOP/s this way vs 52.7 with scalar FFMA.
        float2 acc2[F1 / 2][TC_R];
#pragma unroll
        for (int p = 0; p < F1 / 2; ++p)
#pragma unroll
            for (int q = 0; q < TC_R; ++q) acc2[p][q] = make_float2(0.f, 0.f);

        for (int k0 = 0; k0 < K1p; k0 += KU) {
            float xw[TC_R + KU];
#pragma unroll
            for (int q = 0; q < (TC_R + KU) / 4; ++q) {
                float4 v = *reinterpret_cast<const float4 *>(xr + k0 + 4 * q);
                xw[4 * q + 0] = v.x; xw[4 * q + 1] = v.y; xw[4 * q + 2] = v.z; xw[4 * q + 3] = v.w;
            }
#pragma unroll
            for (int kk = 0; kk < KU; ++kk) {
                float2 w2[F1 / 2];
#pragma unroll
                for (int q = 0; q < F1 / 4; ++q) {
                    float4 v = *reinterpret_cast<const float4 *>(ws + (k0 + kk) * F1 + 4 * q);
                    w2[2 * q] = make_float2(v.x, v.y);
                    w2[2 * q + 1] = make_float2(v.z, v.w);
                }
#pragma unroll
                for (int q = 0; q < TC_R; ++q) {
                    const float2 xx = make_float2(xw[kk + q], xw[kk + q]);
#pragma unroll
                    for (int p = 0; p < F1 / 2; ++p) acc2[p][q] = __ffma2_rn(w2[p], xx, acc2[p][q]);
                }
            }
        }
        float acc[F1][TC_R];
#pragma unroll
        for (int p = 0; p < F1 / 2; ++p)
#pragma unroll
            for (int q = 0; q < TC_R; ++q) { acc[2 * p][q] = acc2[p][q].x; acc[2 * p + 1][q] = acc2[p][q].y; }

        const int c = c0 + r;
        const int tg = tile0 + t0;
        const bool row_ok = c < C;
        if (row_ok) {
#pragma unroll
            for (int f = 0; f < F1; ++f) {
                float *dst = y1 + (((int64_t)n * F1 + f) * C + c) * (int64_t)T + tg;
                if ((T & 3) == 0 && tg + TC_R <= T) {
#pragma unroll
                    for (int q4 = 0; q4 < TC_R / 4; ++q4)
                        reinterpret_cast<float4 *>(dst)[q4] = make_float4(acc[f][4 * q4], acc[f][4 * q4 + 1], acc[f][4 * q4 + 2], acc[f][4 * q4 + 3]);
                } else {
#pragma unroll
                    for (int q = 0; q < TC_R; ++q)
                        if (tg + q < T) dst[q] = acc[f][q];
                }
            }
        }
        if (part != nullptr) {
            // per-item partial sums for BatchNorm statistics (deterministic two-stage reduction)
            const int warp = tid >> 5, lane = tid & 31;
#pragma unroll
            for (int f = 0; f < F1; ++f) {
                float s = 0.f, q2 = 0.f;
#pragma unroll
                for (int q = 0; q < TC_R; ++q) {
                    float v = (row_ok && tg + q < T) ? acc[f][q] : 0.f;
                    s += v;
                    q2 = fmaf(v, v, q2);
                }
                s = warp_sum(s);
                q2 = warp_sum(q2);
                if (lane == 0) { red[warp][2 * f] = s; red[warp][2 * f + 1] = q2; }
            }
            __syncthreads();
            if (tid < 2 * F1) {
                float s = 0.f;
#pragma unroll
                for (int w = 0; w < TC_THREADS / 32; ++w) s += red[w][tid];
                part[(int64_t)item * (2 * F1) + tid] = s;      // item-major == model-major rows
            }
        }
    }
}

// variant table (EAV_TC_VARIANT env var for A/B runs; default = the fastest measured)
static int tc_variant() {
    static int v = -1;
    if (v < 0) { const char *e = getenv("EAV_TC_VARIANT"); v = e ? atoi(e) : 3; }   // measured on B200 with FFMA2: 3 (4 rows, 8 taps/iter) 1.76 ms, 2: 1.83, 6: 1.84, 5: 1.85, 4: 1.90
    return v;
}
static int tc_rows() { int v = tc_variant(); return (v == 0 || v == 1) ? 3 : (v == 5 ? 8 : 4); }

int tconv_fwd_rows_per_sample(const NetDims &d) {
    if (tconv_fwd_use_tc(d)) return tconv_fwd_tc_rows_per_sample(d);
    return cdiv(d.C, tc_rows()) * cdiv(d.T, TC_TT);
}

template <int ROWS, int KU, int MINB, int R>
static int launch_tconv_fwd_v(const NetDims &d, const float *x, const int32_t *x_index, const float *params,
                              float *y1, float *part, int *part_rows, cudaStream_t st) {
    constexpr int THREADS = ROWS * (TC_TT / R);
    const int K1p = (d.K1 + KU - 1) / KU * KU;
    const int XS = TC_TT + K1p + 4;
    size_t smem = (size_t)(K1p * d.F1 + 2 * ROWS * XS) * sizeof(float);
    EAV_REQUIRE(smem <= 200 * 1024, EAV_ERR_UNSUPPORTED, "tconv_fwd: kernLength %d too large", d.K1);
    const int groups = cdiv(d.C, ROWS), tiles = cdiv(d.T, TC_TT);
    const int64_t n_items = (int64_t)d.N * groups * tiles;
    static PerDevice<int> resident_pd(0);          // CTAs the device holds at once (SMs x occupancy)
    static PerDevice<size_t> resident_smem_pd(0);
    int &resident = resident_pd.here();
    size_t &resident_smem = resident_smem_pd.here();
    if (resident == 0 || resident_smem != smem) {
        cudaFuncSetAttribute(tconv_fwd_kernel<8, ROWS, KU, MINB, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        int dev = 0, sms = 148, per_sm = 1;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, tconv_fwd_kernel<8, ROWS, KU, MINB, R>, THREADS, smem);
        resident = sms * (per_sm > 0 ? per_sm : 1);
        resident_smem = smem;
    }
    const int grid = (int)(n_items < resident ? n_items : resident);
    tconv_fwd_kernel<8, ROWS, KU, MINB, R><<<grid, THREADS, smem, st>>>(x, x_index, params, d.pstride, d.oW1, d.B, d.C, d.T,
                                                                       d.K1, d.pad1l, (int)n_items, groups, tiles, y1, part);
    EAV_CUDA_LAUNCH_CHECK("tconv_fwd");
    if (part_rows) *part_rows = d.B * groups * tiles;
    return 0;
}

int launch_tconv_fwd(const NetDims &d, const float *x, const int32_t *x_index, const float *params,
                     float *wt_scratch, float *y1, float *part, int *part_rows, cudaStream_t st) {
    if (tconv_fwd_use_tc(d))   // tcgen05 path (tconv_tc.cu); EAV_TCONV=ffma selects the CUDA-core kernel below
        return launch_tconv_fwd_tc(d, x, x_index, params, wt_scratch, y1, part, part_rows, st);
    EAV_REQUIRE(d.F1 == 8, EAV_ERR_UNSUPPORTED, "tconv_fwd: F1=%d unsupported (only 8)", d.F1);
    switch (tc_variant()) {
        case 0: return launch_tconv_fwd_v<3, 4, 3, 8>(d, x, x_index, params, y1, part, part_rows, st);
        case 1: return launch_tconv_fwd_v<3, 8, 3, 8>(d, x, x_index, params, y1, part, part_rows, st);
        case 3: return launch_tconv_fwd_v<4, 8, 2, 8>(d, x, x_index, params, y1, part, part_rows, st);
        case 4: return launch_tconv_fwd_v<4, 4, 3, 16>(d, x, x_index, params, y1, part, part_rows, st);   // 128 thr, 16 outputs/thread
        case 5: return launch_tconv_fwd_v<8, 4, 1, 16>(d, x, x_index, params, y1, part, part_rows, st);   // 256 thr, 16 outputs/thread
        case 6: return launch_tconv_fwd_v<4, 8, 3, 16>(d, x, x_index, params, y1, part, part_rows, st);
        default: return launch_tconv_fwd_v<4, 4, 2, 8>(d, x, x_index, params, y1, part, part_rows, st);
    }
}

// =================================================================================
// BatchNorm statistics (nn.BatchNorm2d, eps/momentum as the reference: EEGNet_tor.py:25,29,38)
// part: [M][rows_per_model][ch][2] fp32 partial (sum, sum of squares); one warp per
// (channel, model) reduces them in fp64 in a fixed order.
// =================================================================================
__global__ void bn_reduce_kernel(const float *__restrict__ part, int rows_per_model, int ch,
                                 double *__restrict__ sums) {
    const int c = blockIdx.x, m = blockIdx.y, lane = threadIdx.x;
    double s = 0.0, q = 0.0;
    const float *p = part + ((int64_t)m * rows_per_model) * ch * 2;
    for (int r = lane; r < rows_per_model; r += 32) {
        s += (double)p[((int64_t)r * ch + c) * 2];
        q += (double)p[((int64_t)r * ch + c) * 2 + 1];
    }
    s = warp_sum(s);
    q = warp_sum(q);
    if (lane == 0) { sums[((int64_t)m * ch + c) * 2] = s; sums[((int64_t)m * ch + c) * 2 + 1] = q; }
}

static int layer_channels(const NetDims &d, int layer) { return layer == 1 ? d.F1 : layer == 2 ? d.G : d.F2; }

int launch_bn_reduce(const NetDims &d, int layer, const float *part, int rows_per_model, double *sums,
                     cudaStream_t st) {
    const int ch = layer_channels(d, layer);
    bn_reduce_kernel<<<dim3(ch, d.M), 32, 0, st>>>(part, rows_per_model, ch, sums);
    EAV_CUDA_LAUNCH_CHECK("bn_reduce");
    return 0;
}

__global__ void bn_finalize_kernel(const float *__restrict__ part, int rows_per_model, int ch,
                                   double count, const double *__restrict__ sums,
                                   const float *__restrict__ params, int64_t pstride,
                                   int64_t og, int64_t ob, float *__restrict__ bn_state,
                                   int64_t bnstride, int64_t orm, int64_t orv, int bn_train,
                                   float eps, float momentum, float4 *__restrict__ stats) {
    const int c = blockIdx.x, m = blockIdx.y, lane = threadIdx.x;
    float *rm = bn_state + (int64_t)m * bnstride + orm;
    float *rv = bn_state + (int64_t)m * bnstride + orv;
    float mean, var;
    if (bn_train) {
        double s = 0.0, q = 0.0;
        if (sums != nullptr) {          // already reduced (and all-reduced across replicas)
            s = sums[((int64_t)m * ch + c) * 2];
            q = sums[((int64_t)m * ch + c) * 2 + 1];
        } else {
            const float *p = part + ((int64_t)m * rows_per_model) * ch * 2;
            for (int r = lane; r < rows_per_model; r += 32) {
                s += (double)p[((int64_t)r * ch + c) * 2];
                q += (double)p[((int64_t)r * ch + c) * 2 + 1];
            }
            s = warp_sum(s);
            q = warp_sum(q);
        }
        double mu = s / count;
        double vb = q / count - mu * mu;
        if (vb < 0.0) vb = 0.0;
        mean = (float)mu;
        var = (float)vb;
        if (lane == 0) {
            double unbiased = count > 1.0 ? vb * count / (count - 1.0) : vb;
            rm[c] = (1.f - momentum) * rm[c] + momentum * mean;
            rv[c] = (1.f - momentum) * rv[c] + momentum * (float)unbiased;
        }
    } else {
        mean = rm[c];
        var = rv[c];
    }
    if (lane == 0) {
        float invstd = 1.0f / sqrtf(var + eps);
        float g = params[(int64_t)m * pstride + og + c];
        float b = params[(int64_t)m * pstride + ob + c];
        float scale = g * invstd;
        stats[(int64_t)m * ch + c] = make_float4(mean, invstd, scale, b - mean * scale);
    }
}

// Eval mode: the three BatchNorm layers are per-channel affines of the running statistics, known before any
// activation exists -> ONE launch at the start of the forward instead of three small ones.
__global__ void bn_eval_finalize_all_kernel(const float *__restrict__ params, int64_t pstride, const float *__restrict__ bn_state,
                                            int64_t bnstride, int F1, int G, int F2, int64_t og1, int64_t ob1, int64_t og2,
                                            int64_t ob2, int64_t og3, int64_t ob3, int64_t orm1, int64_t orv1, int64_t orm2,
                                            int64_t orv2, int64_t orm3, int64_t orv3, float eps, float4 *__restrict__ s1,
                                            float4 *__restrict__ s2, float4 *__restrict__ s3) {
    const int m = blockIdx.x;
    for (int i = threadIdx.x; i < F1 + G + F2; i += blockDim.x) {
        int c, ch;
        int64_t og, ob, orm, orv;
        float4 *dst;
        if (i < F1) { c = i; ch = F1; og = og1; ob = ob1; orm = orm1; orv = orv1; dst = s1; }
        else if (i < F1 + G) { c = i - F1; ch = G; og = og2; ob = ob2; orm = orm2; orv = orv2; dst = s2; }
        else { c = i - F1 - G; ch = F2; og = og3; ob = ob3; orm = orm3; orv = orv3; dst = s3; }
        const float mean = bn_state[(int64_t)m * bnstride + orm + c], var = bn_state[(int64_t)m * bnstride + orv + c];
        const float invstd = 1.0f / sqrtf(var + eps);
        const float scale = params[(int64_t)m * pstride + og + c] * invstd;
        dst[(int64_t)m * ch + c] = make_float4(mean, invstd, scale, params[(int64_t)m * pstride + ob + c] - mean * scale);
    }
}

int launch_bn_eval_finalize_all(const NetDims &d, const float *params, const float *bn_state, float4 *s1, float4 *s2,
                                float4 *s3, cudaStream_t st) {
    bn_eval_finalize_all_kernel<<<d.M, 160, 0, st>>>(params, d.pstride, bn_state, d.bnstride, d.F1, d.G, d.F2, d.og1, d.ob1,
                                                     d.og2, d.ob2, d.og3, d.ob3, d.orm1, d.orv1, d.orm2, d.orv2, d.orm3,
                                                     d.orv3, d.eps, s1, s2, s3);
    EAV_CUDA_LAUNCH_CHECK("bn_eval_finalize_all");
    return 0;
}

int launch_bn_finalize(const NetDims &d, int layer, const float *part, int rows_per_model,
                       double count, const double *sums, const float *params, float *bn_state, float4 *stats,
                       cudaStream_t st) {
    int ch;
    int64_t og, ob, orm, orv;
    if (layer == 1) { ch = d.F1; og = d.og1; ob = d.ob1; orm = d.orm1; orv = d.orv1; }
    else if (layer == 2) { ch = d.G; og = d.og2; ob = d.ob2; orm = d.orm2; orv = d.orv2; }
    else { ch = d.F2; og = d.og3; ob = d.ob3; orm = d.orm3; orv = d.orv3; }
    bn_finalize_kernel<<<dim3(ch, d.M), 32, 0, st>>>(part, rows_per_model, ch, count, sums, params, d.pstride,
                                                     og, ob, bn_state, d.bnstride, orm, orv, d.bn_train,
                                                     d.eps, d.momentum, stats);
    EAV_CUDA_LAUNCH_CHECK("bn_finalize");
    return 0;
}

// =================================================================================
// M2+M3+M4  BN1 (+ELU for variant 0) + depthwise spatial conv over the electrodes.
//   y2[n, f*D+d, t] = sum_c W2[f*D+d, c] * act(y1[n,f,c,t]*scale1[f] + shift1[f])
// HBM/L2-bound: reads y1 once (F1*C*T floats per sample), writes G*T.
// =================================================================================
constexpr int DW_THREADS = 128;
constexpr int DMAX = 8;

// VEC = 4: each thread owns 4 consecutive time samples (128-bit loads/stores, needs T % 4 == 0);
// VEC = 1: scalar fallback for any T.
template <int VEC, int UNR>
__global__ void __launch_bounds__(DW_THREADS)
dw_fwd_kernel(const float *__restrict__ y1, const float *__restrict__ params, int64_t pstride,
              int64_t oW2, const float4 *__restrict__ bn1, int B, int F1, int D, int C, int T,
              int elu1, float *__restrict__ y2, float *__restrict__ part, const float4 *__restrict__ bn2_pool,
              float *__restrict__ d1) {
    extern __shared__ __align__(16) float w2s[];  // [C][DMAX]: the D weights of one electrode are two 128-bit reads
    __shared__ float red[DW_THREADS / 32][2 * DMAX];
    const int n = blockIdx.z, f = blockIdx.y, m = n / B;
    const int t = (blockIdx.x * DW_THREADS + threadIdx.x) * VEC;
    const int G = F1 * D;
    const float *W2 = params + (int64_t)m * pstride + oW2 + (int64_t)f * D * C;
    for (int i = threadIdx.x; i < DMAX * C; i += DW_THREADS) {
        const int c = i / DMAX, dd = i - c * DMAX;
        w2s[i] = dd < D ? W2[dd * C + c] : 0.f;
    }
    const float4 st = bn1[(int64_t)m * F1 + f];
    __syncthreads();
    float acc[DMAX][VEC];
#pragma unroll
    for (int dd = 0; dd < DMAX; ++dd)
#pragma unroll
        for (int e = 0; e < VEC; ++e) acc[dd][e] = 0.f;
    if (t < T) {
        const float *src = y1 + (((int64_t)n * F1 + f) * C) * (int64_t)T + t;
        // UNR electrode rows are fetched back to back (UNR independent 128-bit loads in flight per
        // thread) before any of them is consumed: the kernel is HBM-latency bound otherwise.
        for (int c0 = 0; c0 < C; c0 += UNR) {
            float a[UNR][VEC];
#pragma unroll
            for (int u = 0; u < UNR; ++u) {
                const int c = c0 + u;
                if (c < C) {
                    if (VEC == 4) {
                        float4 v = *reinterpret_cast<const float4 *>(src + (int64_t)c * T);
                        a[u][0] = v.x; a[u][1 % VEC] = v.y; a[u][2 % VEC] = v.z; a[u][3 % VEC] = v.w;
                    } else {
                        a[u][0] = src[(int64_t)c * T];
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < UNR; ++u) {
                const int c = c0 + u;
                if (c < C) {
#pragma unroll
                    for (int e = 0; e < VEC; ++e) {
                        float v = fmaf(a[u][e], st.z, st.w);
                        a[u][e] = elu1 ? elu_fast(v) : v;
                    }
                    const float4 wa = *reinterpret_cast<const float4 *>(&w2s[c * DMAX]);
                    const float4 wb = *reinterpret_cast<const float4 *>(&w2s[c * DMAX + 4]);
                    const float wv[DMAX] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
                    for (int dd = 0; dd < DMAX; ++dd)
                        if (dd < D) {
                            if (VEC == 4) {         // packed fp32: two FFMA2 instead of four FFMA, same rounding per element
                                const float2 w2 = make_float2(wv[dd], wv[dd]);
                                const float2 r0 = __ffma2_rn(w2, make_float2(a[u][0], a[u][1 % VEC]), make_float2(acc[dd][0], acc[dd][1 % VEC]));
                                const float2 r1 = __ffma2_rn(w2, make_float2(a[u][2 % VEC], a[u][3 % VEC]), make_float2(acc[dd][2 % VEC], acc[dd][3 % VEC]));
                                acc[dd][0] = r0.x; acc[dd][1 % VEC] = r0.y; acc[dd][2 % VEC] = r1.x; acc[dd][3 % VEC] = r1.y;
                            } else {
#pragma unroll
                                for (int e = 0; e < VEC; ++e) acc[dd][e] = fmaf(wv[dd], a[u][e], acc[dd][e]);
                            }
                        }
                }
            }
        }
#pragma unroll
        for (int dd = 0; dd < DMAX; ++dd)
            if (dd < D) {
                float *dst = y2 + ((int64_t)n * G + f * D + dd) * (int64_t)T + t;
                if (VEC == 4) *reinterpret_cast<float4 *>(dst) = make_float4(acc[dd][0], acc[dd][1 % VEC], acc[dd][2 % VEC], acc[dd][3 % VEC]);
                else dst[0] = acc[dd][0];
                // Eval-mode fusion of M5 (EEGNet_tor.py:55-57): the thread's four time steps are exactly one AvgPool(1,4)
                // window, BatchNorm-2 is a known per-channel affine and there is no dropout, so d1 comes straight out of
                // the registers and pool1_fwd (a 172 MB re-read of y2) is not launched.
                if (VEC == 4 && bn2_pool != nullptr) {
                    const float4 s2 = bn2_pool[(int64_t)m * G + f * D + dd];
                    d1[((int64_t)n * G + f * D + dd) * (int64_t)(T >> 2) + (t >> 2)] =
                        0.25f * (elu_fast(fmaf(acc[dd][0], s2.z, s2.w)) + elu_fast(fmaf(acc[dd][1 % VEC], s2.z, s2.w)) +
                                 elu_fast(fmaf(acc[dd][2 % VEC], s2.z, s2.w)) + elu_fast(fmaf(acc[dd][3 % VEC], s2.z, s2.w)));
                }
            }
    }
    if (part != nullptr) {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
        for (int dd = 0; dd < DMAX; ++dd) {
            float s = 0.f, q = 0.f;
            if (t < T && dd < D) {
#pragma unroll
                for (int e = 0; e < VEC; ++e) { s += acc[dd][e]; q = fmaf(acc[dd][e], acc[dd][e], q); }
            }
            s = warp_sum(s);
            q = warp_sum(q);
            if (lane == 0) { red[warp][2 * dd] = s; red[warp][2 * dd + 1] = q; }
        }
        __syncthreads();
        if (threadIdx.x < 2 * D) {
            float s = 0.f;
#pragma unroll
            for (int w = 0; w < DW_THREADS / 32; ++w) s += red[w][threadIdx.x];
            int64_t row = (int64_t)n * gridDim.x + blockIdx.x;
            int dd = threadIdx.x >> 1, which = threadIdx.x & 1;
            part[(row * G + f * D + dd) * 2 + which] = s;
        }
    }
}

int dw_fwd_tiles(const NetDims &d) { return (d.T & 3) == 0 ? cdiv(d.T, DW_THREADS * 4) : cdiv(d.T, DW_THREADS); }

bool dw_fwd_fuses_pool(const NetDims &d) {
    return !d.bn_train && d.dropout_mode == EAV_DROPOUT_NONE && d.P1 == 4 && (d.T & 3) == 0 && tc_path_enabled("EAV_FUSE_FWD");
}

int launch_dw_fwd(const NetDims &d, const float *y1, const float *params, const float4 *bn1,
                  float *y2, float *part, int *part_rows, const float4 *bn2_pool, float *d1, cudaStream_t st) {
    EAV_REQUIRE(d.D <= DMAX, EAV_ERR_UNSUPPORTED, "dw_fwd: D=%d > %d unsupported", d.D, DMAX);
    dim3 grid(dw_fwd_tiles(d), d.F1, d.N);
    const size_t smem = (size_t)DMAX * d.C * sizeof(float);
    const int elu1 = d.variant == EAV_VARIANT_TOR;
    static int unr = -1;
    if (unr < 0) { const char *e = getenv("EAV_DWF_UNR"); unr = e ? atoi(e) : 5; }   // measured on B200: 5 -> 0.195 ms, 3: 0.203, 6: 0.209, 10: 0.256, 15: 0.434
    if ((d.T & 3) == 0) {
        if (unr == 3) dw_fwd_kernel<4, 3><<<grid, DW_THREADS, smem, st>>>(y1, params, d.pstride, d.oW2, bn1, d.B, d.F1, d.D, d.C, d.T, elu1, y2, part, bn2_pool, d1);
        else if (unr == 6) dw_fwd_kernel<4, 6><<<grid, DW_THREADS, smem, st>>>(y1, params, d.pstride, d.oW2, bn1, d.B, d.F1, d.D, d.C, d.T, elu1, y2, part, bn2_pool, d1);
        else if (unr == 15) dw_fwd_kernel<4, 15><<<grid, DW_THREADS, smem, st>>>(y1, params, d.pstride, d.oW2, bn1, d.B, d.F1, d.D, d.C, d.T, elu1, y2, part, bn2_pool, d1);
        else if (unr == 10) dw_fwd_kernel<4, 10><<<grid, DW_THREADS, smem, st>>>(y1, params, d.pstride, d.oW2, bn1, d.B, d.F1, d.D, d.C, d.T, elu1, y2, part, bn2_pool, d1);
        else dw_fwd_kernel<4, 5><<<grid, DW_THREADS, smem, st>>>(y1, params, d.pstride, d.oW2, bn1, d.B, d.F1, d.D, d.C, d.T, elu1, y2, part, bn2_pool, d1);
    } else {
        dw_fwd_kernel<1, 8><<<grid, DW_THREADS, smem, st>>>(y1, params, d.pstride, d.oW2, bn1, d.B, d.F1, d.D, d.C, d.T, elu1, y2, part, bn2_pool, d1);
    }
    EAV_CUDA_LAUNCH_CHECK("dw_fwd");
    if (part_rows) *part_rows = d.B * grid.x;
    return 0;
}

// =================================================================================
// M5  BN2 + ELU + AvgPool2d((1,P1)) + dropout -> d1[n,g,u]
// =================================================================================
// One warp per (n, g) row: no per-element index divisions, lanes walk the pooled positions.
__global__ void __launch_bounds__(256)
pool1_fwd_kernel(const float *__restrict__ y2, const float4 *__restrict__ bn2,
                 const uint8_t *__restrict__ mask1, int B, int G, int T, int T4, int P1,
                 int dropout_mode, float p_drop, uint64_t seed, uint64_t step, const unsigned long long *__restrict__ step_ptr,
                 int64_t rows, float *__restrict__ d1) {
    const int lane = threadIdx.x & 31;
    const int64_t row = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    if (step_ptr) step = *step_ptr;
    const float inv_keep = (dropout_mode != EAV_DROPOUT_NONE && p_drop < 1.f) ? 1.f / (1.f - p_drop) : 1.f;
    const int g = (int)(row % G), n = (int)(row / G);
    const float4 st = bn2[(int64_t)(n / B) * G + g];
    const float *src = y2 + row * (int64_t)T;
    const bool vec = P1 == 4 && (T & 3) == 0;      // one aligned 128-bit load per pooling window
    const float invp = 1.f / (float)P1;
    if (vec) {
        // Lane = one block of four consecutive output elements e = 4*blk .. 4*blk+3 (global element index, so the
        // block is exactly one Philox counter): four independent 128-bit loads in flight and one Philox call per
        // four dropout decisions.
        const int64_t e0 = row * T4, eb0 = e0 & ~(int64_t)3;
        const int nblk = (int)((e0 - eb0 + T4 + 3) >> 2);
        for (int blk = lane; blk < nblk; blk += 32) {
            const int64_t eb = eb0 + 4 * (int64_t)blk;
            float4 v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int64_t u = eb + j - e0;
                v[j] = (u >= 0 && u < T4) ? *reinterpret_cast<const float4 *>(src + 4 * u) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            uint32_t keep = 0xFu;
            if (dropout_mode == EAV_DROPOUT_PHILOX) keep = philox_keep4(seed, step, 1u, (uint64_t)(eb >> 2), p_drop);
            else if (dropout_mode == EAV_DROPOUT_PHILOX_2D) keep = philox_keep(seed, step, 17u, (uint64_t)row, p_drop) ? 0xFu : 0u;
            float o[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int64_t u = eb + j - e0;
                float sj = (elu_fast(fmaf(v[j].x, st.z, st.w)) + elu_fast(fmaf(v[j].y, st.z, st.w)) +
                            elu_fast(fmaf(v[j].z, st.z, st.w)) + elu_fast(fmaf(v[j].w, st.z, st.w))) * invp;
                if (dropout_mode == EAV_DROPOUT_MASK) sj = (u >= 0 && u < T4 && mask1[eb + j]) ? sj * inv_keep : 0.f;
                else if (dropout_mode >= EAV_DROPOUT_PHILOX) sj = ((keep >> j) & 1u) ? sj * inv_keep : 0.f;
                o[j] = sj;
            }
            if (eb >= e0 && eb + 3 < e0 + T4) {
                *reinterpret_cast<float4 *>(d1 + eb) = make_float4(o[0], o[1], o[2], o[3]);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (eb + j >= e0 && eb + j < e0 + T4) d1[eb + j] = o[j];
            }
        }
        return;
    }
    for (int u = lane; u < T4; u += 32) {
        float s = 0.f;
        for (int w = 0; w < P1; ++w) s += elu_fast(fmaf(src[u * P1 + w], st.z, st.w));
        s *= invp;
        const int64_t e = row * T4 + u;
        if (dropout_mode == EAV_DROPOUT_MASK) s = mask1[e] ? s * inv_keep : 0.f;
        else if (dropout_mode == EAV_DROPOUT_PHILOX) s = philox_keep(seed, step, 1u, (uint64_t)e, p_drop) ? s * inv_keep : 0.f;
        else if (dropout_mode == EAV_DROPOUT_PHILOX_2D) s = philox_keep(seed, step, 17u, (uint64_t)row, p_drop) ? s * inv_keep : 0.f;
        d1[e] = s;
    }
}

int launch_pool1_fwd(const NetDims &d, const float *y2, const float4 *bn2, const uint8_t *mask1,
                     float *d1, cudaStream_t st) {
    const int64_t rows = (int64_t)d.N * d.G;
    pool1_fwd_kernel<<<(unsigned)cdiv64(rows, 8), 256, 0, st>>>(y2, bn2, mask1, d.B, d.G, d.T, d.T4, d.P1, d.dropout_mode,
                                                               d.p_drop, d.seed, d.step, d.step_ptr, rows, d1);
    EAV_CUDA_LAUNCH_CHECK("pool1_fwd");
    return 0;
}

// =================================================================================
// M6  (1,16) 'same' convolution over all input channels (EEGNet_tor.py:37,59) and its
// input gradient (MODE 1), as a register-blocked implicit GEMM on the CUDA cores:
//   MODE 0: out[n,o,u] = sum_g sum_k W3[o,g,k] * in[n,g,u+k-7]          in = d1
//   MODE 1: out[n,g,u] = sum_o sum_k W3[o,g,15-k] * in[n,o,u+k-8]       in = dy3 (BN3 bwd applied on load)
// CTA = 2 samples of one model x 64 output channels x 128 positions; thread = 8 channels x
// 4 positions x 2 samples (64 accumulators); weights streamed through smem 4 input
// channels at a time.  16.4 MFLOP per epoch per direction.
// =================================================================================
constexpr int SC_K = 16;       // kernel length (fixed by the reference)
constexpr int SC_SC = 2;       // samples per CTA
constexpr int SC_GC = 4;       // input channels per weight chunk
constexpr int SC_UT = 128;     // positions per CTA
constexpr int SC_XS = SC_UT + SC_K;  // 144: padded input row
constexpr int SC_CO = 64;      // output channels per CTA

__device__ __forceinline__ void cp_async16_cg(void *smem_dst, const void *gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}

// Work items: `n_pairs` two-sample items first, then single-sample items for what is left, so
// that the last (partial) wave of CTAs runs half-size items instead of idling the GPU.
// Weight chunks (4 input channels x 64 outputs x 16 taps = 16 KB) are double-buffered with
// 16-byte cp.async: chunk i+1 streams in while chunk i feeds the FFMA loop.  In MODE 1 the
// weights stay unflipped in smem ([o'][g][k] is contiguous in global) and the flip is done by
// indexing the register window backwards.
template <int MODE>
__global__ void __launch_bounds__(256, 2)
sepconv_kernel(const float *__restrict__ in, const float *__restrict__ yraw,
               const float4 *__restrict__ bnf, const float4 *__restrict__ bnb, int bn_train,
               const float *__restrict__ params, int64_t pstride, int64_t oW3, int B, int Cin, int Cout,
               int L, int padl, int n_pairs, int pairs_per_model, int singles_per_model,
               float *__restrict__ out, float *__restrict__ part) {
    extern __shared__ __align__(16) float smem[];
    float *ws0 = smem;                               // 2 x [SC_GC][SC_CO][SC_K]
    float *xs = smem + 2 * SC_GC * SC_CO * SC_K;     // [SC_SC][Cin][SC_XS]
    // decode the work item -> (model, first sample, number of samples)
    int m, b0, ns, part_row;
    {
        const int it = blockIdx.z;
        if (it < n_pairs) {
            m = it / pairs_per_model;
            const int p = it - m * pairs_per_model;
            b0 = p * SC_SC;
            ns = min(SC_SC, B - b0);
            part_row = m * (pairs_per_model + singles_per_model) + p;
        } else {
            const int j = it - n_pairs;                 // only reached when singles_per_model > 0
            m = j / max(singles_per_model, 1);
            const int q = j - m * singles_per_model;
            b0 = pairs_per_model * SC_SC + q;
            ns = 1;
            part_row = m * (pairs_per_model + singles_per_model) + pairs_per_model + q;
        }
    }
    const int o_base = blockIdx.y * SC_CO;
    const int u0 = blockIdx.x * SC_UT;
    const int tid = threadIdx.y * 32 + threadIdx.x;
    const float *W3 = params + (int64_t)m * pstride + oW3;
    const bool w_vec = (reinterpret_cast<uintptr_t>(W3) & 15) == 0;    // arena slices are 16-byte aligned

    auto stage_w = [&](int g0, float *ws) {
        // ws[gl][o][k]
        if (w_vec) {
            for (int i = tid; i < SC_GC * SC_CO * (SC_K / 4); i += 256) {
                const int gl = i / (SC_CO * 4), rem = i - gl * SC_CO * 4;
                const int o = rem >> 2, k4 = (rem & 3) * 4;
                const int gi = g0 + gl, oc = o_base + o;
                float *dst = ws + (gl * SC_CO + o) * SC_K + k4;
                if (gi < Cin && oc < Cout) {
                    const float *src = (MODE == 0) ? W3 + ((int64_t)oc * Cin + gi) * SC_K + k4      // W3[o][g][k]
                                                   : W3 + ((int64_t)gi * Cout + oc) * SC_K + k4;    // W3[o'=gi][g=oc][k]
                    cp_async16_cg(dst, src);
                } else {
                    *reinterpret_cast<float4 *>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
        } else {
            for (int i = tid; i < SC_GC * SC_CO * SC_K; i += 256) {
                const int gl = i / (SC_CO * SC_K), rem = i - gl * SC_CO * SC_K;
                const int o = rem / SC_K, k = rem - o * SC_K;
                const int gi = g0 + gl, oc = o_base + o;
                float w = 0.f;
                if (gi < Cin && oc < Cout)
                    w = (MODE == 0) ? W3[((int64_t)oc * Cin + gi) * SC_K + k] : W3[((int64_t)gi * Cout + oc) * SC_K + k];
                ws[i] = w;
            }
        }
        cp_async_commit();
    };

    stage_w(0, ws0);
    // stage the input rows (with the BatchNorm-backward transform in MODE 1)
    for (int i = tid; i < SC_SC * Cin * SC_XS; i += 256) {
        int s = i / (Cin * SC_XS);
        int rem = i - s * Cin * SC_XS;
        int ch = rem / SC_XS, j = rem - ch * SC_XS;
        int b = b0 + s;
        int u = u0 - padl + j;
        float v = 0.f;
        if (s < ns && u >= 0 && u < L) {
            int64_t idx = (((int64_t)(m * B + b)) * Cin + ch) * (int64_t)L + u;
            v = in[idx];
            if (MODE == 1 && bnb != nullptr) {   // (unused since bn_bwd_apply materialises dy3; kept for callers that skip it)
                const float4 kb = bnb[(int64_t)m * Cin + ch];
                if (bn_train) {
                    const float4 kf = bnf[(int64_t)m * Cin + ch];
                    float xhat = (yraw[idx] - kf.x) * kf.y;
                    v = kb.x * (v - kb.y - xhat * kb.z);
                } else {
                    v = kb.x * v;
                }
            }
        }
        xs[i] = v;
    }

    float acc[SC_SC][8][4];
#pragma unroll
    for (int s = 0; s < SC_SC; ++s)
#pragma unroll
        for (int oo = 0; oo < 8; ++oo)
#pragma unroll
            for (int uu = 0; uu < 4; ++uu) acc[s][oo][uu] = 0.f;

    int buf = 0;
    for (int g0 = 0; g0 < Cin; g0 += SC_GC, buf ^= 1) {
        cp_async_wait<0>();
        __syncthreads();                               // chunk g0 (and, first time, xs) visible; other buffer free
        if (g0 + SC_GC < Cin) stage_w(g0 + SC_GC, ws0 + (buf ^ 1) * (SC_GC * SC_CO * SC_K));
        const float *ws = ws0 + buf * (SC_GC * SC_CO * SC_K);
#pragma unroll 1
        for (int gl = 0; gl < SC_GC; ++gl) {
            if (g0 + gl >= Cin) break;
            float xw[SC_SC][20];
#pragma unroll
            for (int s = 0; s < SC_SC; ++s) {
                const float *xr = xs + ((int64_t)s * Cin + g0 + gl) * SC_XS + 4 * threadIdx.x;
#pragma unroll
                for (int q = 0; q < 5; ++q) {
                    float4 v = *reinterpret_cast<const float4 *>(xr + 4 * q);
                    xw[s][4 * q] = v.x; xw[s][4 * q + 1] = v.y; xw[s][4 * q + 2] = v.z; xw[s][4 * q + 3] = v.w;
                }
            }
            const float *wr = ws + (gl * SC_CO + 8 * threadIdx.y) * SC_K;
#pragma unroll
            for (int k = 0; k < SC_K; k += 4) {
#pragma unroll
                for (int oo = 0; oo < 8; ++oo) {
                    float4 w4 = *reinterpret_cast<const float4 *>(wr + oo * SC_K + k);
                    const float wv[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk)
#pragma unroll
                        for (int s = 0; s < SC_SC; ++s)
#pragma unroll
                            for (int uu = 0; uu < 4; ++uu) {
                                // MODE 0: in[u + k - padl];  MODE 1 (flipped kernel): in[u - k + (K-1) - padl']
                                const int wi = (MODE == 0) ? (k + kk + uu) : (SC_K - 1 - (k + kk) + uu);
                                acc[s][oo][uu] = fmaf(wv[kk], xw[s][wi], acc[s][oo][uu]);
                            }
                }
            }
        }
    }

    const int ub = u0 + 4 * threadIdx.x;
#pragma unroll
    for (int oo = 0; oo < 8; ++oo) {
        const int oc = o_base + 8 * threadIdx.y + oo;
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int s = 0; s < SC_SC; ++s) {
            const int b = b0 + s;
            if (s < ns && oc < Cout) {
                float *dst = out + (((int64_t)(m * B + b)) * Cout + oc) * (int64_t)L + ub;
#pragma unroll
                for (int uu = 0; uu < 4; ++uu)
                    if (ub + uu < L) {
                        dst[uu] = acc[s][oo][uu];
                        s1 += acc[s][oo][uu];
                        s2 = fmaf(acc[s][oo][uu], acc[s][oo][uu], s2);
                    }
            }
        }
        if (MODE == 0 && part != nullptr) {
            s1 = warp_sum(s1);
            s2 = warp_sum(s2);
            if (threadIdx.x == 0 && oc < Cout) {
                int64_t row = (int64_t)part_row * gridDim.x + blockIdx.x;
                part[(row * Cout + oc) * 2] = s1;
                part[(row * Cout + oc) * 2 + 1] = s2;
            }
        }
    }
}

// item plan shared by both modes: full waves of two-sample items, the remainder as singles
struct SepPlan { int pairs_per_model, singles_per_model; };
static SepPlan sepconv_plan(const NetDims &d, int ctas_xy) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int slots = sms * 2 / (ctas_xy > 0 ? ctas_xy : 1);      // 2 resident CTAs per SM
    const int pairs_all = d.B / SC_SC;                            // per model, whole pairs only
    int singles = d.B - pairs_all * SC_SC;                        // odd batch: one single anyway
    int pairs = pairs_all;
    const int64_t total_pairs = (int64_t)d.M * pairs_all;
    static int split_tail = -1;
    if (split_tail < 0) { const char *e = getenv("EAV_SEP_SPLIT_TAIL"); split_tail = e ? atoi(e) : 0; }   // measured: splitting the tail wave is slower (0.55 vs 0.47 ms)
    if (split_tail && slots > 0 && total_pairs > slots) {
        // pairs beyond the last full wave are converted to singles, spread evenly over the models
        const int64_t tail = total_pairs % slots;
        int conv = (int)cdiv64(tail, d.M);                        // per model
        if (conv > pairs) conv = pairs;
        if (tail > 0 && tail * 2 <= slots) { pairs -= conv; singles += conv * SC_SC; }
    }
    return {pairs, singles};
}
int sepconv_fwd_rows_per_model(const NetDims &d) {
    if (sepconv_use_tc(d)) return sepconv_tc_rows_per_model(d);
    SepPlan p = sepconv_plan(d, cdiv(d.T4, SC_UT) * cdiv(d.F2, SC_CO));
    return (p.pairs_per_model + p.singles_per_model) * cdiv(d.T4, SC_UT);
}

static size_t sepconv_smem(int Cin) {
    return (size_t)(2 * SC_GC * SC_CO * SC_K + SC_SC * Cin * SC_XS) * sizeof(float);
}

int launch_sepconv_fwd(const NetDims &d, const float *d1, const float *params, float *wt_scratch, float *y3,
                       float *part, int *part_rows, cudaStream_t st) {
    if (sepconv_use_tc(d)) return launch_sepconv_tc(d, 0, d1, params, wt_scratch, y3, part, part_rows, st);
    EAV_REQUIRE(d.K2 == SC_K, EAV_ERR_UNSUPPORTED, "sepconv: kernel length %d unsupported (only 16)", d.K2);
    size_t smem = sepconv_smem(d.G);
    EAV_REQUIRE(smem <= 200 * 1024, EAV_ERR_UNSUPPORTED, "sepconv: F1*D=%d too large", d.G);
    static PerDevice<bool> attr_set_pd(false);
    bool &attr_set = attr_set_pd.here();
    if (!attr_set) {
        cudaFuncSetAttribute(sepconv_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        attr_set = true;
    }
    const SepPlan pl = sepconv_plan(d, cdiv(d.T4, SC_UT) * cdiv(d.F2, SC_CO));
    dim3 grid(cdiv(d.T4, SC_UT), cdiv(d.F2, SC_CO), d.M * (pl.pairs_per_model + pl.singles_per_model));
    sepconv_kernel<0><<<grid, dim3(32, 8), smem, st>>>(d1, nullptr, nullptr, nullptr, 0, params, d.pstride,
                                                       d.oW3, d.B, d.G, d.F2, d.T4, d.pad2l,
                                                       d.M * pl.pairs_per_model, pl.pairs_per_model,
                                                       pl.singles_per_model, y3, part);
    EAV_CUDA_LAUNCH_CHECK("sepconv_fwd");
    if (part_rows) *part_rows = (pl.pairs_per_model + pl.singles_per_model) * grid.x;
    return 0;
}

int launch_sepconv_bwd_dx(const NetDims &d, const float *dz3, const float *y3, const float4 *bnf3,
                          const float4 *bnb3, const float *params, float *wt_scratch, float *dd1, cudaStream_t st) {
    if (sepconv_use_tc(d)) return launch_sepconv_tc(d, 1, dz3, params, wt_scratch, dd1, nullptr, nullptr, st);
    size_t smem = sepconv_smem(d.F2);
    EAV_REQUIRE(smem <= 200 * 1024, EAV_ERR_UNSUPPORTED, "sepconv_dx: F2=%d too large", d.F2);
    static PerDevice<bool> attr_set_pd(false);
    bool &attr_set = attr_set_pd.here();
    if (!attr_set) {
        cudaFuncSetAttribute(sepconv_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        attr_set = true;
    }
    const SepPlan pl = sepconv_plan(d, cdiv(d.T4, SC_UT) * cdiv(d.G, SC_CO));
    dim3 grid(cdiv(d.T4, SC_UT), cdiv(d.G, SC_CO), d.M * (pl.pairs_per_model + pl.singles_per_model));
    // flipped kernel: left padding K-1-pad2l
    sepconv_kernel<1><<<grid, dim3(32, 8), smem, st>>>(dz3, y3, bnf3, nullptr, d.bn_train, params, d.pstride,
                                                       d.oW3, d.B, d.F2, d.G, d.T4, d.K2 - 1 - d.pad2l,
                                                       d.M * pl.pairs_per_model, pl.pairs_per_model,
                                                       pl.singles_per_model, dd1,
                                                       nullptr);
    EAV_CUDA_LAUNCH_CHECK("sepconv_bwd_dx");
    return 0;
}

// =================================================================================
// Variant 1 (CNN_EEG.py:35-37) block 2: depthwise temporal conv, then pointwise conv.
// =================================================================================
__global__ void dwt_fwd_kernel(const float *__restrict__ d1, const float *__restrict__ params,
                               int64_t pstride, int64_t oW3, int B, int G, int L, int K2, int padl,
                               int64_t total, float *__restrict__ y3d) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        int u = (int)(i % L);
        int64_t ng = i / L;
        int g = (int)(ng % G), n = (int)(ng / G);
        const float *w = params + (int64_t)(n / B) * pstride + oW3 + (int64_t)g * K2;
        const float *src = d1 + ng * (int64_t)L;
        float s = 0.f;
        for (int k = 0; k < K2; ++k) {
            int uu = u + k - padl;
            if (uu >= 0 && uu < L) s = fmaf(w[k], src[uu], s);
        }
        y3d[i] = s;
    }
}

int launch_dwt_fwd(const NetDims &d, const float *d1, const float *params, float *y3d, cudaStream_t st) {
    int64_t total = (int64_t)d.N * d.G * d.T4;
    int blocks = (int)std::min<int64_t>(cdiv64(total, 256), 148 * 16);
    dwt_fwd_kernel<<<blocks, 256, 0, st>>>(d1, params, d.pstride, d.oW3, d.B, d.G, d.T4, d.K2, d.pad2l, total, y3d);
    EAV_CUDA_LAUNCH_CHECK("dwt_fwd");
    return 0;
}

// pointwise: y3[n,o,u] = sum_g W3p[o,g] * y3d[n,g,u]; one warp-row per (n,o): lanes over u.
__global__ void __launch_bounds__(128)
pw_fwd_kernel(const float *__restrict__ y3d, const float *__restrict__ params, int64_t pstride,
              int64_t oW3p, int B, int G, int F2, int L, float *__restrict__ y3, float *__restrict__ part) {
    extern __shared__ float wsh[];  // [G]
    const int n = blockIdx.y, o = blockIdx.x, m = n / B;
    const float *w = params + (int64_t)m * pstride + oW3p + (int64_t)o * G;
    for (int i = threadIdx.x; i < G; i += blockDim.x) wsh[i] = w[i];
    __syncthreads();
    float s1 = 0.f, s2 = 0.f;
    for (int u = threadIdx.x; u < L; u += blockDim.x) {
        const float *src = y3d + ((int64_t)n * G) * L + u;
        float s = 0.f;
        for (int g = 0; g < G; ++g) s = fmaf(wsh[g], src[(int64_t)g * L], s);
        y3[((int64_t)n * F2 + o) * L + u] = s;
        s1 += s;
        s2 = fmaf(s, s, s2);
    }
    if (part != nullptr) {
        __shared__ float red[4][2];
        s1 = warp_sum(s1);
        s2 = warp_sum(s2);
        if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5][0] = s1; red[threadIdx.x >> 5][1] = s2; }
        __syncthreads();
        if (threadIdx.x < 2) {
            float s = red[0][threadIdx.x] + red[1][threadIdx.x] + red[2][threadIdx.x] + red[3][threadIdx.x];
            part[((int64_t)n * F2 + o) * 2 + threadIdx.x] = s;   // one partial row per sample
        }
    }
}

// Tiled variant for F2, G <= 64 and T4 <= 128 (the EAV shape): one CTA per sample stages y3d once (the kernel above
// makes every (sample, output channel) CTA re-read the sample's 32 KB of y3d: 64x the L2 traffic); warp = 8 output
// channels, lane = 4 positions, acc[8][4]; per input channel 2 broadcast LDS.128 + 4 LDS per 32 FFMA.
constexpr int PWF_P = 129, PWF_W = 64;
__global__ void __launch_bounds__(256)
pw_fwd_tiled_kernel(const float *__restrict__ y3d, const float *__restrict__ params, int64_t pstride,
                    int64_t oW3p, int B, int G, int F2, int L, float *__restrict__ y3, float *__restrict__ part) {
    extern __shared__ __align__(16) float sm[];
    float *as = sm;                  // [64][PWF_P]  y3d, zero padded
    float *wt = as + 64 * PWF_P;     // [64][PWF_W]  W3p transposed: wt[g][o]
    const int n = blockIdx.x, m = n / B, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 64 * PWF_P; i += 256) as[i] = 0.f;
    for (int i = tid; i < 64 * PWF_W; i += 256) wt[i] = 0.f;
    __syncthreads();
    for (int i = tid; i < G * L; i += 256) {
        const int g = i / L, u = i - g * L;
        as[g * PWF_P + u] = y3d[(int64_t)n * G * L + i];
    }
    const float *W = params + (int64_t)m * pstride + oW3p;
    for (int i = tid; i < F2 * G; i += 256) wt[(i % G) * PWF_W + (i / G)] = W[i];
    __syncthreads();
    float acc[8][4];
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
    const int o0 = 8 * warp;
    for (int g = 0; g < G; ++g) {
        const float4 w0 = *reinterpret_cast<const float4 *>(wt + g * PWF_W + o0);
        const float4 w1 = *reinterpret_cast<const float4 *>(wt + g * PWF_W + o0 + 4);
        const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
        float xv[4];
#pragma unroll
        for (int b = 0; b < 4; ++b) xv[b] = as[g * PWF_P + lane + 32 * b];
#pragma unroll
        for (int a = 0; a < 8; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(w[a], xv[b], acc[a][b]);
    }
#pragma unroll
    for (int a = 0; a < 8; ++a) {
        const int o = o0 + a;
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int u = lane + 32 * b;
            if (o < F2 && u < L) {
                y3[((int64_t)n * F2 + o) * L + u] = acc[a][b];
                s1 += acc[a][b];
                s2 = fmaf(acc[a][b], acc[a][b], s2);
            }
        }
        if (part != nullptr) {
            s1 = warp_sum(s1);
            s2 = warp_sum(s2);
            if (lane == 0 && o < F2) {
                part[((int64_t)n * F2 + o) * 2] = s1;
                part[((int64_t)n * F2 + o) * 2 + 1] = s2;
            }
        }
    }
}

int launch_pw_fwd(const NetDims &d, const float *y3d, const float *params, float *y3, float *part,
                  int *part_rows, cudaStream_t st) {
    if (d.F2 <= 64 && d.G <= 64 && d.T4 <= 128) {
        const size_t tsm = (size_t)(64 * PWF_P + 64 * PWF_W) * sizeof(float);
        static PerDevice<bool> attr_pd(false);
        bool &attr = attr_pd.here();
        if (!attr) {
            cudaFuncSetAttribute(pw_fwd_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tsm);
            attr = true;
        }
        pw_fwd_tiled_kernel<<<d.N, 256, tsm, st>>>(y3d, params, d.pstride, d.oW3p, d.B, d.G, d.F2, d.T4, y3, part);
    } else {
        pw_fwd_kernel<<<dim3(d.F2, d.N), 128, (size_t)d.G * sizeof(float), st>>>(y3d, params, d.pstride, d.oW3p,
                                                                               d.B, d.G, d.F2, d.T4, y3, part);
    }
    EAV_CUDA_LAUNCH_CHECK("pw_fwd");
    if (part_rows) *part_rows = d.B;
    return 0;
}

// =================================================================================
// M7+M8  BN3 + ELU + AvgPool2d((1,P2)) + dropout + flatten + dense (+ softmax).
// One CTA per sample.  The max-norm hook on the dense weight runs AFTER this kernel
// (separate launch) so that the forward uses W_old.
// =================================================================================
__global__ void __launch_bounds__(128)
tail_fwd_kernel(const float *__restrict__ y3, const float4 *__restrict__ bn3,
                const uint8_t *__restrict__ mask2, const float *__restrict__ params, int64_t pstride,
                int64_t oWd, int64_t obd, int B, int F2, int T4, int T32, int P2, int NC, int softmax_out,
                int dropout_mode, float p_drop, uint64_t seed, uint64_t step, const unsigned long long *__restrict__ step_ptr,
                float *__restrict__ feat, float *__restrict__ out, float *__restrict__ probs_saved) {
    extern __shared__ float sm[];  // feat_s[FEAT] + z_s[NC]
    const int FEAT = F2 * T32;
    float *feat_s = sm, *z_s = sm + FEAT;
    const int n = blockIdx.x, m = n / B, tid = threadIdx.x;
    if (step_ptr) step = *step_ptr;
    const float inv_keep = (dropout_mode != EAV_DROPOUT_NONE && p_drop < 1.f) ? 1.f / (1.f - p_drop) : 1.f;
    for (int i = tid; i < FEAT; i += blockDim.x) {
        int o = i / T32, v = i - o * T32;
        const float4 st = bn3[(int64_t)m * F2 + o];
        const float *src = y3 + ((int64_t)n * F2 + o) * T4 + v * P2;
        float s = 0.f;
        for (int w = 0; w < P2; ++w) s += elu_fast(fmaf(src[w], st.z, st.w));
        s *= 1.f / (float)P2;
        int64_t e = (int64_t)n * FEAT + i;
        if (dropout_mode == EAV_DROPOUT_MASK) s = mask2[e] ? s * inv_keep : 0.f;
        else if (dropout_mode == EAV_DROPOUT_PHILOX) s = philox_keep(seed, step, 2u, (uint64_t)e, p_drop) ? s * inv_keep : 0.f;
        else if (dropout_mode == EAV_DROPOUT_PHILOX_2D) s = philox_keep(seed, step, 18u, (uint64_t)((int64_t)n * F2 + o), p_drop) ? s * inv_keep : 0.f;
        feat_s[i] = s;
        feat[e] = s;
    }
    __syncthreads();
    const int warp = tid >> 5, lane = tid & 31;
    const float *Wd = params + (int64_t)m * pstride + oWd;
    for (int j = warp; j < NC; j += blockDim.x >> 5) {
        float s = 0.f;
        for (int i = lane; i < FEAT; i += 32) s = fmaf(Wd[(int64_t)j * FEAT + i], feat_s[i], s);
        s = warp_sum(s);
        if (lane == 0) z_s[j] = s + params[(int64_t)m * pstride + obd + j];
    }
    __syncthreads();
    if (tid == 0) {
        if (softmax_out) {
            float mx = -INFINITY;
            for (int j = 0; j < NC; ++j) mx = fmaxf(mx, z_s[j]);
            float den = 0.f;
            for (int j = 0; j < NC; ++j) den += expf(z_s[j] - mx);
            for (int j = 0; j < NC; ++j) {
                float p = expf(z_s[j] - mx) / den;
                out[(int64_t)n * NC + j] = p;
                probs_saved[(int64_t)n * NC + j] = p;
            }
        } else {
            for (int j = 0; j < NC; ++j) {
                out[(int64_t)n * NC + j] = z_s[j];
                probs_saved[(int64_t)n * NC + j] = z_s[j];
            }
        }
    }
}

int launch_tail_fwd(const NetDims &d, const float *y3, const float4 *bn3, const uint8_t *mask2,
                    const float *params, float *feat, float *out, float *probs_saved, cudaStream_t st) {
    size_t smem = (size_t)(d.FEAT + d.NC) * sizeof(float);
    EAV_REQUIRE(smem <= 48 * 1024, EAV_ERR_UNSUPPORTED, "tail_fwd: feature size %d too large", d.FEAT);
    tail_fwd_kernel<<<d.N, 128, smem, st>>>(y3, bn3, mask2, params, d.pstride, d.oWd, d.obd, d.B, d.F2, d.T4,
                                            d.T32, d.P2, d.NC, d.variant == EAV_VARIANT_TOR, d.dropout_mode,
                                            d.p_drop, d.seed, d.step, d.step_ptr, feat, out, probs_saved);
    EAV_CUDA_LAUNCH_CHECK("tail_fwd");
    return 0;
}

}  // namespace eav

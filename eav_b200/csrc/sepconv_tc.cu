// Tensor-core (tcgen05 + TMEM) kernel for the block-2 (1,16) convolution over 64 channels
// (EEGNet_tor.py:37 `nn.Conv2d(F1*D, F2, (1,16), padding='same')`): forward and input gradient.
//
//   out[n][o][u] = sum_c sum_k Wk[o][c][k] * in[n][c][u + k - PL]          (64 x 64 x 16 taps, U = T/4 positions)
//     forward : in = d1,  Wk[o][c][k] = W3[o][c][k],          PL = pad2l
//     d(input): in = dy3, Wk[g][o][k] = W3[o][g][K2-1-k],     PL = K2-1-pad2l
//
// GEMM view per sample: M = 128 positions, N = 64 output channels, K = 8 input channels per MMA, one MMA per
// (tap, 8-channel group) = 128 of them.  The activations are staged CHANNEL-INTERLEAVED, act[c/4][row][c%4] with
// row = u + PL (zero rows around), so that for a fixed 4-channel chunk consecutive positions are 16 B apart: that is
// the no-swizzle K-major canonical layout {SBO = 128 B, LBO = chunk pitch}, and tap k is a start-address shift of
// k * 16 B -- no im2col, no shifted copies.  fp32 parity through the 3-term tf32 split:
//   D[:, 0:128] = a_hi * [w_hi | w_lo]   (N = 128)     D[:, 0:64] += a_lo * w_hi   (N = 64);  out = D[:, o] + D[:, 64+o].
// The packed weights of one model are 512 KB (hi and lo), more than shared memory: they are streamed from L2
// through a TMA ring while TWO samples (both resident in shared memory) consume every block.
// Roles: warps 0-3 epilogue, warp 4 MMA issue, warp 5 weight TMA, warps 6-13 activation producers.
#include <cstdlib>
#include <cstring>

#include "eav_common.cuh"
#include "eegnet_kernels.cuh"
#include "tc_common.cuh"

namespace eav {
namespace {

constexpr int SCT_PROD_WARPS = 8;         // 16 four-channel chunks / 2 per warp
constexpr int SCT_THREADS = (6 + SCT_PROD_WARPS) * 32;
constexpr int SCT_C = 64;                 // input channels == output channels
constexpr int SCT_K = 16;                 // taps
constexpr int SCT_ROWS = 144;             // activation rows: 128 positions + 16 taps
constexpr int SCT_CHUNK = SCT_ROWS * 4;   // floats per 4-channel chunk
constexpr int SCT_ACT = 16 * SCT_CHUNK;   // floats per activation buffer (one of hi / lo, one sample) = 36864 B
constexpr int SCT_BLOCK = 128 * 8;        // floats per weight block (tap, 8-channel group): [w_hi | w_lo] x 8
constexpr int SCT_NBLOCKS = SCT_K * (SCT_C / 8);       // 128
constexpr int SCT_BPS = 4;                // weight blocks per ring stage (16 KB)
constexpr int SCT_WSTAGES = 3;
constexpr int SCT_NITER = SCT_NBLOCKS / SCT_BPS;       // ring stages consumed per sample pair
constexpr size_t SCT_SMEM = ((size_t)4 * SCT_ACT + (size_t)SCT_WSTAGES * SCT_BPS * SCT_BLOCK) * 4;   // 196608 B

// out[m][blk = k*8 + c8][n/8][cc/4][n%8][cc%4]:  n < 64 -> hi(Wk[n][8 c8 + cc][k]),  n >= 64 -> lo
template <int MODE>
__global__ void sepconv_wt_pack_kernel(const float *__restrict__ params, int64_t pstride, int64_t oW3,
                                       float *__restrict__ out) {
    const int m = blockIdx.y, blk = blockIdx.x;
    const int k = blk >> 3, c8 = blk & 7;
    for (int idx = threadIdx.x; idx < SCT_BLOCK; idx += blockDim.x) {
        const int n = idx >> 3, cc = idx & 7;
        const int o = n & 63, c = c8 * 8 + cc;
        const float *W3 = params + (int64_t)m * pstride + oW3;           // [F2][G][K2]
        const float w = MODE == 0 ? W3[(o * SCT_C + c) * SCT_K + k] : W3[(c * SCT_C + o) * SCT_K + (SCT_K - 1 - k)];
        float hi, lo;
        tc::split_tf32(w, hi, lo);
        out[((int64_t)m * SCT_NBLOCKS + blk) * SCT_BLOCK + (n >> 3) * 64 + (cc >> 2) * 32 + (n & 7) * 4 + (cc & 3)] =
            n < 64 ? hi : lo;
    }
}

__global__ void __launch_bounds__(SCT_THREADS, 1)
sepconv_tc_kernel(const float *__restrict__ in, const float *__restrict__ wt_packed, float *__restrict__ out,
                  float *__restrict__ part, int M, int B, int U, int PL) {
    extern __shared__ __align__(128) float smem[];
    float *act = smem;                             // [2 samples][hi, lo][16 chunks][144 rows][4]
    float *wring = smem + 4 * SCT_ACT;             // [WSTAGES][BPS][BLOCK]
    __shared__ uint64_t bar_wfull[SCT_WSTAGES], bar_wempty[SCT_WSTAGES], bar_actfull, bar_actempty, bar_accfull[2],
        bar_accempty[2];
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ppm = (B + 1) / 2;                   // sample pairs per model
    const int n_pairs = M * ppm;
    const int p_lo = (int)((int64_t)n_pairs * blockIdx.x / gridDim.x);
    const int p_hi = (int)((int64_t)n_pairs * (blockIdx.x + 1) / gridDim.x);
    const int np = p_hi - p_lo;
    if (np <= 0) return;

    if (tid == 0) {
        for (int s = 0; s < SCT_WSTAGES; ++s) { tc::mbar_init(&bar_wfull[s], 1); tc::mbar_init(&bar_wempty[s], 1); }
        tc::mbar_init(&bar_actfull, SCT_PROD_WARPS);
        tc::mbar_init(&bar_actempty, 1);
        for (int b = 0; b < 2; ++b) { tc::mbar_init(&bar_accfull[b], 1); tc::mbar_init(&bar_accempty[b], 4); }
        tc::mbar_init_fence();
    }
    if (warp == 4) tc::tmem_alloc(&tmem_slot, 512);
    for (int i = tid; i < 4 * SCT_ACT; i += SCT_THREADS) act[i] = 0.f;     // padding rows stay zero
    tc::fence_proxy_async_smem();
    tc::tc_fence_before_sync();
    __syncthreads();
    tc::tc_fence_after_sync();
    const uint32_t tmem = tmem_slot;

    auto pair_samples = [&](int p, int &m, int &n0, int &cnt) {
        m = p / ppm;
        const int j = p - m * ppm;
        n0 = m * B + 2 * j;
        cnt = (2 * j + 1 < B) ? 2 : 1;
    };

    if (warp >= 6) {
        // ---------------- activation producers: [c][u] rows -> channel-interleaved hi / lo ----------------
        // The global loads of pair lp are issued BEFORE waiting for the MMAs of pair lp-1 to release the
        // buffers, so only the split + shared-memory stores sit between two pairs' MMA streams.
        const int pw = warp - 6;
        for (int lp = 0; lp < np; ++lp) {
            int m, n0, cnt;
            pair_samples(p_lo + lp, m, n0, cnt);
            float4 v[2][2][4];
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                const float *src = in + (int64_t)(n0 + (s < cnt ? s : 0)) * SCT_C * U;
#pragma unroll
                for (int ci = 0; ci < 2; ++ci) {
                    const float *r0 = src + (int64_t)(4 * (pw + SCT_PROD_WARPS * ci)) * U;
#pragma unroll
                    for (int it = 0; it < 4; ++it) {
                        const int u = lane + 32 * it;
                        v[s][ci][it] = u < U ? make_float4(__ldg(r0 + u), __ldg(r0 + U + u), __ldg(r0 + 2 * U + u),
                                                           __ldg(r0 + 3 * U + u))
                                             : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                }
            }
            if (lp > 0) tc::mbar_wait(&bar_actempty, (lp - 1) & 1);
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                if (s < cnt) {
                    float *hi = act + (size_t)s * 2 * SCT_ACT, *lo = hi + SCT_ACT;
#pragma unroll
                    for (int ci = 0; ci < 2; ++ci) {
                        const int c4 = pw + SCT_PROD_WARPS * ci;
#pragma unroll
                        for (int it = 0; it < 4; ++it) {
                            const int u = lane + 32 * it;
                            if (u < U) {
                                float4 h, l;
                                tc::split_tf32(v[s][ci][it].x, h.x, l.x);
                                tc::split_tf32(v[s][ci][it].y, h.y, l.y);
                                tc::split_tf32(v[s][ci][it].z, h.z, l.z);
                                tc::split_tf32(v[s][ci][it].w, h.w, l.w);
                                const int o = c4 * SCT_CHUNK + (u + PL) * 4;
                                *reinterpret_cast<float4 *>(hi + o) = h;
                                *reinterpret_cast<float4 *>(lo + o) = l;
                            }
                        }
                    }
                }
            }
            tc::fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&bar_actfull);
        }
    } else if (warp == 5) {
        // ---------------- weight stream: 16 KB bulk copies into the ring ----------------
        if (tc::elect_one()) {
            int g = 0;
            for (int lp = 0; lp < np; ++lp) {
                int m, n0, cnt;
                pair_samples(p_lo + lp, m, n0, cnt);
                const float *wsrc = wt_packed + (int64_t)m * SCT_NBLOCKS * SCT_BLOCK;
                for (int it = 0; it < SCT_NITER; ++it, ++g) {
                    const int st = g % SCT_WSTAGES, use = g / SCT_WSTAGES;
                    if (use > 0) tc::mbar_wait(&bar_wempty[st], (use - 1) & 1);
                    tc::mbar_expect_tx(&bar_wfull[st], SCT_BPS * SCT_BLOCK * 4);
                    tc::tma_load_1d(wring + (size_t)st * SCT_BPS * SCT_BLOCK, wsrc + (size_t)it * SCT_BPS * SCT_BLOCK,
                                    SCT_BPS * SCT_BLOCK * 4, &bar_wfull[st]);
                }
            }
        }
    } else if (warp == 4) {
        // ---------------- MMA issuer ----------------
        const bool leader = tc::elect_one();
        const uint32_t idesc128 = tc::idesc_tf32(128, 128, 0, 0), idesc64 = tc::idesc_tf32(128, 64, 0, 0);
        const uint32_t act_u32 = tc::smem_u32(act), ring_u32 = tc::smem_u32(wring);
        int g = 0;
        for (int lp = 0; lp < np; ++lp) {
            int m, n0, cnt;
            pair_samples(p_lo + lp, m, n0, cnt);
            const int buf = lp & 1;
            tc::mbar_wait(&bar_actfull, lp & 1);
            if (lp >= 2) tc::mbar_wait(&bar_accempty[buf], ((lp >> 1) - 1) & 1);
            for (int it = 0; it < SCT_NITER; ++it, ++g) {
                const int st = g % SCT_WSTAGES, use = g / SCT_WSTAGES;
                tc::mbar_wait(&bar_wfull[st], use & 1);
                tc::tc_fence_after_sync();
                if (leader) {
#pragma unroll
                    for (int bb = 0; bb < SCT_BPS; ++bb) {
                        const int blk = it * SCT_BPS + bb;
                        const int k = blk >> 3, c8 = blk & 7;
                        const uint64_t bd = tc::smem_desc(ring_u32 + (uint32_t)((st * SCT_BPS + bb) * SCT_BLOCK * 4), 128, 256);
                        const uint32_t aoff = (uint32_t)((2 * c8 * SCT_CHUNK + k * 4) * 4);
                        for (int s = 0; s < cnt; ++s) {
                            const uint32_t a_hi = act_u32 + (uint32_t)(s * 2 * SCT_ACT * 4) + aoff;
                            const uint64_t ah = tc::smem_desc(a_hi, SCT_CHUNK * 4, 128);
                            const uint64_t al = tc::smem_desc(a_hi + SCT_ACT * 4, SCT_CHUNK * 4, 128);
                            const uint32_t d = tmem + buf * 256 + s * 128;
                            tc::mma_tf32_ss(d, ah, bd, idesc128, blk > 0 ? 1u : 0u);
                            tc::mma_tf32_ss(d, al, bd, idesc64, 1u);
                        }
                    }
                    tc::mma_commit(&bar_wempty[st]);
                    if (it == SCT_NITER - 1) {
                        tc::mma_commit(&bar_actempty);
                        tc::mma_commit(&bar_accfull[buf]);
                    }
                }
                __syncwarp();
            }
        }
    } else {
        // ---------------- epilogue: TMEM -> out (+ BatchNorm-3 partial sums in forward mode) ----------------
        const int u = warp * 32 + lane;
        const bool valid = u < U;
        for (int lp = 0; lp < np; ++lp) {
            int m, n0, cnt;
            pair_samples(p_lo + lp, m, n0, cnt);
            const int buf = lp & 1;
            tc::mbar_wait(&bar_accfull[buf], (lp >> 1) & 1);
            tc::tc_fence_after_sync();
            for (int s = 0; s < cnt; ++s) {
                float *dst = out + (int64_t)(n0 + s) * SCT_C * U + u;
                float *prow = part ? part + ((int64_t)(n0 + s) * 4 + warp) * (2 * SCT_C) : nullptr;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    float v0[32], v1[32];
                    const uint32_t ta = tmem + ((uint32_t)(warp * 32) << 16) + buf * 256 + s * 128 + h * 32;
                    tc::tmem_ld32(ta, v0);
                    tc::tmem_ld32(ta + 64, v1);
                    tc::tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const float v = v0[j] + v1[j];
                        if (valid) dst[(int64_t)(h * 32 + j) * U] = v;
                        if (prow) {
                            const float vv = valid ? v : 0.f;
                            const float sm = warp_sum(vv), sq = warp_sum(vv * vv);
                            if (lane == 0) { prow[2 * (h * 32 + j)] = sm; prow[2 * (h * 32 + j) + 1] = sq; }
                        }
                    }
                }
            }
            tc::tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&bar_accempty[buf]);
        }
    }
    tc::tc_fence_before_sync();
    __syncthreads();
    if (warp == 4) tc::tmem_dealloc(tmem, 512);
}

}  // namespace

bool sepconv_use_tc(const NetDims &d) {
    if (!tc_path_enabled("EAV_SEPCONV")) return false;
    return d.variant == EAV_VARIANT_TOR && d.G == SCT_C && d.F2 == SCT_C && d.K2 == SCT_K && d.T4 >= 1 && d.T4 <= 128;
}
size_t sepconv_tc_scratch_floats(const NetDims &d) {
    return sepconv_use_tc(d) ? (size_t)2 * d.M * SCT_NBLOCKS * SCT_BLOCK : 0;     // forward pack, d(input) pack
}
int sepconv_tc_rows_per_model(const NetDims &d) { return d.B * 4; }

// mode 0: forward (in = d1, out = y3, BN-3 partials);  mode 1: input gradient (in = dy3, out = dd1)
int launch_sepconv_tc(const NetDims &d, int mode, const float *in, const float *params, float *wt_scratch, float *out,
                      float *part, int *part_rows, cudaStream_t st) {
    EAV_REQUIRE(wt_scratch != nullptr, EAV_ERR_BAD_ARG, "sepconv_tc: no weight scratch");
    float *wt = wt_scratch + (mode ? (size_t)d.M * SCT_NBLOCKS * SCT_BLOCK : 0);
    if (mode == 0) sepconv_wt_pack_kernel<0><<<dim3(SCT_NBLOCKS, d.M), 256, 0, st>>>(params, d.pstride, d.oW3, wt);
    else sepconv_wt_pack_kernel<1><<<dim3(SCT_NBLOCKS, d.M), 256, 0, st>>>(params, d.pstride, d.oW3, wt);
    EAV_CUDA_LAUNCH_CHECK("sepconv_wt_pack");
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(sepconv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SCT_SMEM);
        EAV_REQUIRE(e == cudaSuccess, (int)e, "sepconv_tc: %s", cudaGetErrorString(e));
        attr_set = true;
    }
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    }
    const int n_pairs = d.M * ((d.B + 1) / 2);
    const int grid = n_pairs < sms ? n_pairs : sms;
    const int PL = mode == 0 ? d.pad2l : d.K2 - 1 - d.pad2l;
    sepconv_tc_kernel<<<grid, SCT_THREADS, SCT_SMEM, st>>>(in, wt, out, mode == 0 ? part : nullptr, d.M, d.B, d.T4, PL);
    EAV_CUDA_LAUNCH_CHECK(mode == 0 ? "sepconv_fwd_tc" : "sepconv_bwd_dx_tc");
    if (part_rows) *part_rows = sepconv_tc_rows_per_model(d);
    return 0;
}

}  // namespace eav

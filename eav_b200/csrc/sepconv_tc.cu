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

// out[m][blk = k*8 + c8][n/8][cc/4][n%8][cc%4]:  n < 64 -> hi(Wk[n][8 c8 + cc][k]),  n >= 64 -> lo.
// Block (c8, m) stages the 64 x 8 x 16 weights it needs in shared memory with coalesced reads, then writes its
// 16 tap blocks with coalesced stores.
template <int MODE>
__global__ void __launch_bounds__(256) sepconv_wt_pack_kernel(const float *__restrict__ params, int64_t pstride,
                                                              int64_t oW3, float *__restrict__ out) {
    __shared__ float ws[64][8 * SCT_K + 1];          // [n][cc*16 + k]
    const int m = blockIdx.y, c8 = blockIdx.x;
    const float *W3 = params + (int64_t)m * pstride + oW3;           // [F2][G][K2]
    for (int i = threadIdx.x; i < 64 * 8 * SCT_K; i += blockDim.x) {
        // MODE 0: rows n = o, columns c = 8 c8 + cc: W3[n][8 c8 .. 8 c8 + 7][0..15] is 128 contiguous floats
        // MODE 1: rows n = g, columns c = o:           W3[8 c8 + cc][n][0..15]
        if (MODE == 0) {
            const int n = i >> 7, r = i & 127;
            ws[n][r] = W3[(n * SCT_C + c8 * 8) * SCT_K + r];
        } else {
            const int cc = i >> 10, r = i & 1023, n = r >> 4, k = r & 15;
            ws[n][cc * SCT_K + (SCT_K - 1 - k)] = W3[((c8 * 8 + cc) * SCT_C) * SCT_K + r];
        }
    }
    __syncthreads();
    for (int k = 0; k < SCT_K; ++k) {
        float *dst = out + ((int64_t)m * SCT_NBLOCKS + k * 8 + c8) * SCT_BLOCK;
        for (int idx = threadIdx.x; idx < SCT_BLOCK; idx += blockDim.x) {
            // idx = (n/8)*64 + (cc/4)*32 + (n%8)*4 + cc%4
            const int n = (idx >> 6) * 8 + ((idx >> 2) & 7), cc = ((idx >> 5) & 1) * 4 + (idx & 3);
            float hi, lo;
            tc::split_tf32(ws[n & 63][cc * SCT_K + k], hi, lo);
            dst[idx] = n < 64 ? hi : lo;
        }
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
            if (lp > 0) tc::mbar_wait_sleep(&bar_actempty, (lp - 1) & 1, 200);      // a whole pair's MMAs away: sleep, do not spin
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
                    if (use > 0) tc::mbar_wait_sleep(&bar_wempty[st], (use - 1) & 1, 40);
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
            tc::mbar_wait_sleep(&bar_actfull, lp & 1, 40);
            if (lp >= 2) tc::mbar_wait_sleep(&bar_accempty[buf], ((lp >> 1) - 1) & 1, 40);
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
            tc::mbar_wait_sleep(&bar_accfull[buf], (lp >> 1) & 1, 200);
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

// =============================================================================================
// Weight gradient of the block-2 conv on the tensor cores.
//
//   dW3[o][g][k] = sum_{samples} sum_u dy3[o][u] * d1[g][u + k - pad]
// With k = 4 k4 + 2 sb + sa and u' = u + sa:
//   D_k4[(sa,o)][(sb,g)] = sum_{u'} A[(sa,o)][u'] * B[(sb,g)][u' + 4 k4],
//   A[(sa,o)][u'] = dy3[o][u' - sa],   B[(sb,g)][rho] = d1[g][rho + 2 sb - pad]
// i.e. M = 128 = two row-shifted copies of dy3^T, N = 128 = two row-shifted copies of d1^T, K = positions, and the
// four accumulators D_0..D_3 (4 x 128 columns = all of TMEM) hold the 16 taps.  Operands are MN-major
// (rows = positions, 128 B = 32 channels per row, 128B/32B-atom swizzle); tap group k4 is a start-address shift of
// 4 rows.  A whole sample (dy3 and d1, 32 KB each) is bulk-copied into a staging buffer; eight producer warps
// transpose it 16 positions at a time (stride-U reads are conflict-free for odd U) into a 2-stage operand ring,
// applying the tf32 hi/lo split.  One work unit = a range of samples of one model; D stays in TMEM over the unit.
// =============================================================================================
constexpr int SDW_R = 16;                           // positions per chunk
constexpr int SDW_BROWS = SDW_R + 12;               // rows of the d1 operand a chunk touches (4 tap groups)
constexpr int SDW_A = 4 * SDW_R * 32;               // floats, one of hi / lo
constexpr int SDW_B = 4 * SDW_BROWS * 32;
constexpr int SDW_STAGE = 2 * SDW_A + 2 * SDW_B;    // 11264 floats = 45056 B
constexpr int SDW_STAGES = 2;
constexpr int SDW_STG = 64 * 128;                   // staging floats per tensor (U <= 128)
constexpr int SDW_PROD_WARPS = 8;
constexpr int SDW_THREADS = (6 + SDW_PROD_WARPS) * 32;
constexpr size_t SDW_SMEM = ((size_t)SDW_STAGES * SDW_STAGE + 4 * SDW_STG) * 4;    // 221184 B: ring + 2 staged samples

__device__ __forceinline__ uint32_t sdw_swz(uint32_t a) { return a ^ (((a >> 7) & 3u) << 5); }
__device__ __forceinline__ void sdw_sts4(uint32_t addr, float4 v) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ uint64_t sdw_desc(uint32_t saddr, uint32_t lbo) {
    return tc::smem_desc(saddr, lbo, 512) | ((uint64_t)1 << 61);     // MN-major, SWIZZLE_128B_BASE32B
}

__global__ void __launch_bounds__(SDW_THREADS, 1)
sepconv_dw_tc_kernel(const float *__restrict__ dy3, const float *__restrict__ d1, float *__restrict__ part, int M,
                     int B, int U, int pad, int S) {
    extern __shared__ __align__(1024) float smem[];
    float *ring = smem;                                   // [STAGES][A hi, A lo, B hi, B lo]
    float *stg = smem + SDW_STAGES * SDW_STAGE;           // 2 x [dy3 sample][d1 sample]
    __shared__ uint64_t bar_full[SDW_STAGES], bar_empty[SDW_STAGES], bar_stgfull[2], bar_stgempty[2], bar_accfull,
        bar_accempty;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_units = M * S;
    const int n_chunks = (U + 1 + SDW_R - 1) / SDW_R;

    if (tid == 0) {
        for (int s = 0; s < SDW_STAGES; ++s) { tc::mbar_init(&bar_full[s], SDW_PROD_WARPS); tc::mbar_init(&bar_empty[s], 1); }
        for (int b = 0; b < 2; ++b) { tc::mbar_init(&bar_stgfull[b], 1); tc::mbar_init(&bar_stgempty[b], SDW_PROD_WARPS); }
        tc::mbar_init(&bar_accfull, 1);
        tc::mbar_init(&bar_accempty, 4);
        tc::mbar_init_fence();
    }
    if (warp == 4) tc::tmem_alloc(&tmem_slot, 512);
    tc::tc_fence_before_sync();
    __syncthreads();
    tc::tc_fence_after_sync();
    const uint32_t tmem = tmem_slot;

    auto unit_range = [&](int u, int &m, int &b_lo, int &b_hi) {
        m = u / S;
        const int sp = u - m * S;
        b_lo = (int)((int64_t)B * sp / S);
        b_hi = (int)((int64_t)B * (sp + 1) / S);
    };

    if (warp >= 6) {
        // ---------------- producers: staging [ch][u] -> transposed, swizzled, hi / lo operand chunks ----------------
        // Item = (operand copy, channel half, 4 rows): lane (a = lane % 8, b = lane / 8) gathers channels 4a..4a+3 of
        // row 4 rq + b from the staged [ch][u] sample (4 conflict-free LDS for odd U) and writes one 16-byte piece of
        // the transposed row for each of hi / lo.
        const int pw = warp - 6;
        const int la = lane & 7, lb = lane >> 3;
        int g = 0, ns = 0;
        for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
            int m, b_lo, b_hi;
            unit_range(u, m, b_lo, b_hi);
            for (int b = b_lo; b < b_hi; ++b, ++ns) {
                const float *sdy = stg + (size_t)(ns & 1) * 2 * SDW_STG, *sd1 = sdy + SDW_STG;
                tc::mbar_wait_sleep(&bar_stgfull[ns & 1], (ns >> 1) & 1, 40);
                for (int c = 0; c < n_chunks; ++c, ++g) {
                    const int st = g % SDW_STAGES, use = g / SDW_STAGES;
                    if (use > 0) tc::mbar_wait(&bar_empty[st], (use - 1) & 1);
                    const uint32_t base = tc::smem_u32(ring + (size_t)st * SDW_STAGE);
                    constexpr int NA = 2 * 2 * (SDW_R / 4), NB = 2 * 2 * (SDW_BROWS / 4);
                    for (int item = pw; item < NA + NB; item += SDW_PROD_WARPS) {
                        const float *src;
                        int pos, rows, copy, h, rq;
                        uint32_t a, lo_off;
                        if (item < NA) {          // dy3 operand: A[(sa, o)][u'] = dy3[o][u' - sa]
                            copy = item / (2 * (SDW_R / 4));
                            const int rem = item - copy * 2 * (SDW_R / 4);
                            h = rem / (SDW_R / 4); rq = rem - h * (SDW_R / 4);
                            rows = SDW_R; src = sdy; lo_off = SDW_A * 4;
                            pos = c * SDW_R + 4 * rq + lb - copy;
                            a = base;
                        } else {                  // d1 operand: B[(sb, g)][rho] = d1[g][rho + 2 sb - pad]
                            const int ib = item - NA;
                            copy = ib / (2 * (SDW_BROWS / 4));
                            const int rem = ib - copy * 2 * (SDW_BROWS / 4);
                            h = rem / (SDW_BROWS / 4); rq = rem - h * (SDW_BROWS / 4);
                            rows = SDW_BROWS; src = sd1; lo_off = SDW_B * 4;
                            pos = c * SDW_R + 4 * rq + lb + 2 * copy - pad;
                            a = base + 2 * SDW_A * 4;
                        }
                        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (pos >= 0 && pos < U) {
                            const float *p = src + (h * 32 + 4 * la) * U + pos;
                            v = make_float4(p[0], p[U], p[2 * U], p[3 * U]);
                        }
                        float4 hi, lo;
                        tc::split_tf32(v.x, hi.x, lo.x);
                        tc::split_tf32(v.y, hi.y, lo.y);
                        tc::split_tf32(v.z, hi.z, lo.z);
                        tc::split_tf32(v.w, hi.w, lo.w);
                        a += (uint32_t)(((copy * 2 + h) * rows + 4 * rq + lb) * 128 + la * 16);
                        sdw_sts4(sdw_swz(a), hi);
                        sdw_sts4(sdw_swz(a + lo_off), lo);
                    }
                    tc::fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) tc::mbar_arrive(&bar_full[st]);
                }
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&bar_stgempty[ns & 1]);     // this warp no longer reads the staged sample
            }
        }
    } else if (warp == 5) {
        // ---------------- staging loads: one sample of dy3 and of d1 per bulk-copy pair ----------------
        if (tc::elect_one()) {
            const uint32_t bytes = (uint32_t)(64 * U * 4);
            int ns = 0;
            for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
                int m, b_lo, b_hi;
                unit_range(u, m, b_lo, b_hi);
                for (int b = b_lo; b < b_hi; ++b, ++ns) {
                    const int sb = ns & 1;
                    if (ns >= 2) tc::mbar_wait_sleep(&bar_stgempty[sb], ((ns >> 1) - 1) & 1, 100);
                    const int64_t n = (int64_t)m * B + b;
                    float *dst = stg + (size_t)sb * 2 * SDW_STG;
                    tc::mbar_expect_tx(&bar_stgfull[sb], 2 * bytes);
                    tc::tma_load_1d(dst, dy3 + n * 64 * U, bytes, &bar_stgfull[sb]);
                    tc::tma_load_1d(dst + SDW_STG, d1 + n * 64 * U, bytes, &bar_stgfull[sb]);
                }
            }
        }
    } else if (warp == 4) {
        // ---------------- MMA issuer ----------------
        const bool leader = tc::elect_one();
        const uint32_t idesc = tc::idesc_tf32(128, 128, 1, 1);
        int g = 0, nu = 0;
        for (int u = blockIdx.x; u < n_units; u += gridDim.x, ++nu) {
            int m, b_lo, b_hi;
            unit_range(u, m, b_lo, b_hi);
            if (nu > 0) tc::mbar_wait_sleep(&bar_accempty, (nu - 1) & 1, 100);
            const int total = (b_hi - b_lo) * n_chunks;
            for (int ci = 0; ci < total; ++ci, ++g) {
                const int st = g % SDW_STAGES, use = g / SDW_STAGES;
                tc::mbar_wait(&bar_full[st], use & 1);
                tc::tc_fence_after_sync();
                if (leader) {
                    const uint32_t base = tc::smem_u32(ring + (size_t)st * SDW_STAGE);
                    const uint32_t a_hi = base, a_lo = base + SDW_A * 4;
                    const uint32_t b_hi_ = base + 2 * SDW_A * 4, b_lo_ = b_hi_ + SDW_B * 4;
#pragma unroll
                    for (int k4 = 0; k4 < 4; ++k4) {
#pragma unroll
                        for (int ks = 0; ks < 2; ++ks) {
                            const uint32_t ao = ks * 1024, bo = k4 * 512 + ks * 1024;
                            const uint64_t ah = sdw_desc(a_hi + ao, SDW_R * 128), al = sdw_desc(a_lo + ao, SDW_R * 128);
                            const uint64_t bh = sdw_desc(b_hi_ + bo, SDW_BROWS * 128), bl = sdw_desc(b_lo_ + bo, SDW_BROWS * 128);
                            const uint32_t d = tmem + k4 * 128;
                            tc::mma_tf32_ss(d, ah, bh, idesc, (ci > 0 || ks > 0) ? 1u : 0u);
                            tc::mma_tf32_ss(d, ah, bl, idesc, 1u);
                            tc::mma_tf32_ss(d, al, bh, idesc, 1u);
                        }
                    }
                    tc::mma_commit(&bar_empty[st]);
                    if (ci == total - 1) tc::mma_commit(&bar_accfull);
                }
                __syncwarp();
            }
        }
    } else {
        // ---------------- epilogue: D_k4[(sa,o)][(sb,g)] -> part[unit][k][o][g] ----------------
        int nu = 0;
        const int m_idx = warp * 32 + lane;
        const int sa = m_idx >> 6, o = m_idx & 63;
        for (int u = blockIdx.x; u < n_units; u += gridDim.x, ++nu) {
            tc::mbar_wait_sleep(&bar_accfull, nu & 1, 400);     // a whole unit away (ncu: 1.7 M polls per launch while spinning)
            tc::tc_fence_after_sync();
            float *pu = part + (int64_t)u * 16 * 4096;
#pragma unroll 1
            for (int k4 = 0; k4 < 4; ++k4) {
#pragma unroll 1
                for (int q = 0; q < 4; ++q) {
                    float v[32];
                    tc::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + k4 * 128 + q * 32, v);
                    tc::tmem_ld_wait();
                    const int k = 4 * k4 + 2 * (q >> 1) + sa;
                    float4 *dst = reinterpret_cast<float4 *>(pu + ((int64_t)k * 64 + o) * 64 + (q & 1) * 32);
#pragma unroll
                    for (int j = 0; j < 8; ++j) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                }
            }
            tc::tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&bar_accempty);
        }
    }
    tc::tc_fence_before_sync();
    __syncthreads();
    if (warp == 4) tc::tmem_dealloc(tmem, 512);
}

// grads[m][oW3 + (o*64 + g)*16 + k] = sum_s part[(m*S + s)][k][o][g]   (fixed order).  Thread = (o, g): reads are
// coalesced along g for every (s, k), the 16 taps leave as one contiguous 64-byte store.
__global__ void __launch_bounds__(256) sepconv_dw_reduce_kernel(const float *__restrict__ part, int S, int64_t pstride,
                                                                 int64_t oW3, float *__restrict__ grads) {
    const int m = blockIdx.y;
    const int og = blockIdx.x * blockDim.x + threadIdx.x;      // o * 64 + g
    const float *p = part + (int64_t)m * S * 65536 + og;
    float acc[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) acc[k] = 0.f;
    for (int j = 0; j < S; ++j) {
#pragma unroll
        for (int k = 0; k < 16; ++k) acc[k] += p[(int64_t)j * 65536 + k * 4096];
    }
    float4 *dst = reinterpret_cast<float4 *>(grads + (int64_t)m * pstride + oW3 + (int64_t)og * 16);
    if (((pstride | oW3) & 3) == 0) {
#pragma unroll
        for (int q = 0; q < 4; ++q) dst[q] = make_float4(acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]);
    } else {
        float *d = reinterpret_cast<float *>(dst);
#pragma unroll
        for (int k = 0; k < 16; ++k) d[k] = acc[k];
    }
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
    if (mode == 0) sepconv_wt_pack_kernel<0><<<dim3(8, d.M), 256, 0, st>>>(params, d.pstride, d.oW3, wt);
    else sepconv_wt_pack_kernel<1><<<dim3(8, d.M), 256, 0, st>>>(params, d.pstride, d.oW3, wt);
    EAV_CUDA_LAUNCH_CHECK("sepconv_wt_pack");
    static PerDevice<bool> attr_set_pd(false);
    bool &attr_set = attr_set_pd.here();
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(sepconv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SCT_SMEM);
        EAV_REQUIRE(e == cudaSuccess, (int)e, "sepconv_tc: %s", cudaGetErrorString(e));
        attr_set = true;
    }
    const int sms = device_sm_count();
    const int n_pairs = d.M * ((d.B + 1) / 2);
    const int grid = n_pairs < sms ? n_pairs : sms;
    const int PL = mode == 0 ? d.pad2l : d.K2 - 1 - d.pad2l;
    sepconv_tc_kernel<<<grid, SCT_THREADS, SCT_SMEM, st>>>(in, wt, out, mode == 0 ? part : nullptr, d.M, d.B, d.T4, PL);
    EAV_CUDA_LAUNCH_CHECK(mode == 0 ? "sepconv_fwd_tc" : "sepconv_bwd_dx_tc");
    if (part_rows) *part_rows = sepconv_tc_rows_per_model(d);
    return 0;
}

// sample ranges per model: minimises  waves * (samples per unit * MMA time + epilogue)
int sepconv_dw_tc_splits(const NetDims &d) {
    const int n_chunks = (d.T4 + 1 + SDW_R - 1) / SDW_R;
    int best = 1;
    double best_cost = 1e30;
    for (int S = 1; S <= d.B && S <= 64; ++S) {
        const int64_t units = (int64_t)d.M * S;
        const double waves = (double)((units + 147) / 148);
        const double cost = waves * ((double)((d.B + S - 1) / S) * n_chunks * 24 * 64.0 + 8000.0);
        if (cost < best_cost) { best_cost = cost; best = S; }
    }
    return best;
}

int launch_sepconv_dw_tc(const NetDims &d, const float *dy3, const float *d1, float *part, float *grads, int S,
                         cudaStream_t st) {
    static PerDevice<bool> attr_set_pd(false);
    bool &attr_set = attr_set_pd.here();
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(sepconv_dw_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SDW_SMEM);
        EAV_REQUIRE(e == cudaSuccess, (int)e, "sepconv_dw_tc: %s", cudaGetErrorString(e));
        attr_set = true;
    }
    const int sms = device_sm_count();
    const int units = d.M * S;
    sepconv_dw_tc_kernel<<<units < sms ? units : sms, SDW_THREADS, SDW_SMEM, st>>>(dy3, d1, part, d.M, d.B, d.T4,
                                                                                   d.pad2l, S);
    EAV_CUDA_LAUNCH_CHECK("sepconv_bwd_dw_tc");
    sepconv_dw_reduce_kernel<<<dim3(4096 / 256, d.M), 256, 0, st>>>(part, S, d.pstride, d.oW3, grads);
    EAV_CUDA_LAUNCH_CHECK("sepconv_dw_reduce");
    return 0;
}

}  // namespace eav

// Tensor-core (tcgen05 + TMEM) kernels for the temporal convolution of EEGNet block 1
// (EEGNet_tor.py:24 `nn.Conv2d(1, F1, (1, kernLength), padding='same')` and its weight gradient).
//
// Formulation.  One (sample, electrode) row xs[0..] of the zero-padded signal is a plain
// fp32 array in shared memory.  With  X[col][i] = xs[4*col + i]  the 16-byte chunk (col + i/4) of the raw row IS
// the 4-element K-slice of matrix row `col`, so the no-swizzle UMMA descriptor {LBO = 16 B, SBO = 128 B}
// addresses the Toeplitz operand with no im2col copy at all:
//   forward   D[col][(f,j)] = sum_i X[col][i] * Wt[(f,j)][i],   Wt[(f,j)][i] = w[f][i-j]       (M=128, N=4*F1)
//   weight gr D'[i][(f,j)]  = sum_col X[col][i] * dY[col][(f,j)], dY[col][(f,j)] = dy[f][4col+j]
//             dW[f][k]      = sum_j D'[j+k][(f,j)]                                             (M=i, N=4*F1)
// fp32 accuracy comes from the 3-product tf32 split  hi*hi + hi*lo + lo*hi  accumulated in TMEM (fp32).
#include <cstdlib>
#include <cstring>

#include "eav_common.cuh"
#include "eegnet_kernels.cuh"
#include "tc_common.cuh"
#include "../../include/eav_b200.h"

namespace eav {

// ---------------------------------------------------------------------------------------------
// Probe: runs `reps` x `ksteps` tcgen05.mma.kind::tf32 on a caller-provided shared-memory image with
// caller-provided descriptors and returns the accumulator tile + the elapsed SM cycles.
// It exists so that the descriptor address maps in tc_common.cuh are verified on the device
// (scripts/tc_probe.py) rather than assumed.
// ---------------------------------------------------------------------------------------------
struct ProbeArgs {
    int image_floats, M, N, ksteps, reps, n_acc;
    int a_off, a_lbo, a_sbo, a_major, a_step;
    int b_off, b_lbo, b_sbo, b_major, b_step;
    int a_bits, b_bits;   // OR-ed into bits [32,64) of the A / B shared-memory descriptor (layout_type, base_offset)
};

__global__ void __launch_bounds__(128) tc_probe_kernel(const float *__restrict__ image, ProbeArgs a,
                                                       float *__restrict__ d_out, long long *__restrict__ cycles) {
    extern __shared__ __align__(1024) float sm[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < a.image_floats; i += 128) sm[i] = image[i];
    uint32_t ncols = 32;
    while ((int)ncols < a.N * a.n_acc) ncols <<= 1;
    if (warp == 0) tc::tmem_alloc(&tmem_slot, ncols);
    if (tid == 0) {
        tc::mbar_init(&bar, 1);
        tc::mbar_init_fence();
    }
    tc::fence_proxy_async_smem();
    tc::tc_fence_before_sync();
    __syncthreads();
    tc::tc_fence_after_sync();
    const uint32_t tmem = tmem_slot;
    long long t0 = 0, t1 = 0;
    if (warp == 0) {
        const uint32_t idesc = tc::idesc_tf32(a.M, a.N, a.a_major, a.b_major);
        const uint32_t base = tc::smem_u32(sm);
        const uint64_t ad0 = tc::smem_desc(base + a.a_off, a.a_lbo, a.a_sbo) | ((uint64_t)(uint32_t)a.a_bits << 32);
        const uint64_t bd0 = tc::smem_desc(base + a.b_off, a.b_lbo, a.b_sbo) | ((uint64_t)(uint32_t)a.b_bits << 32);
        const uint32_t astep = (uint32_t)a.a_step >> 4, bstep = (uint32_t)a.b_step >> 4;
        const bool leader = tc::elect_one();
        t0 = clock64();
        int acc = 0;
        uint32_t accumulate = 0;
        for (int r = 0; r < a.reps; ++r) {
            uint64_t ad = ad0, bd = bd0;
            for (int ks = 0; ks < a.ksteps; ++ks) {
                // independent TMEM column ranges break the accumulate dependency chain
                if (leader) tc::mma_tf32_ss(tmem + acc * a.N, ad, bd, idesc, accumulate);
                ad += astep;
                bd += bstep;
                if (++acc == a.n_acc) { acc = 0; accumulate = 1; }
            }
        }
        if (leader) tc::mma_commit(&bar);
        __syncwarp();
    }
    tc::mbar_wait(&bar, 0);
    if (tid == 0) {
        t1 = clock64();
        cycles[0] = t1 - t0;
    }
    tc::tc_fence_after_sync();
    for (int c0 = 0; c0 < a.N; c0 += 32) {
        float v[32];
        tc::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
        tc::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) d_out[(size_t)tid * a.N + c0 + j] = v[j];
    }
    tc::tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, ncols);
}

// =============================================================================================
// Forward temporal convolution on the tensor cores.
//
// Work item = one (sample n, electrode c) row of T samples.  Row buffer xs[PADL + t] = x[t], zeros elsewhere,
// PADL = padl = (K1-1)/2, so  y[f][4*col + j] = sum_i xs[4*col + i] * w[f][i - j]  with i = j + k in [0, K1+3).
//   A (M = 128 cols)  = the raw row buffer through the descriptor {K-major, LBO 16 B, SBO 128 B}; k-step ks
//                       advances the start address by 32 B.  Two copies: hi and lo halves of the tf32 split.
//   B (N = 64)        = packed Toeplitz weights, rows 0..31 = hi(w[f][i-j]) (row = 4f + j), rows 32..63 = lo.
//   D (TMEM, 64 cols) = x_hi * [w_hi | w_lo]  (N = 64)  +  x_lo * w_hi  (N = 32, accumulated into cols 0..31);
//                       the epilogue adds column n and n + 32.
// Roles (7 warps): warps 0-3 epilogue (TMEM lanes 32w..32w+31 = cols), warp 4 issues the MMAs and loads the
// weights with one bulk copy per model, warps 5-6 split x rows into the hi/lo ring.  Two CTAs per SM keep the
// tensor pipe busy across each other's per-MMA latency (tc_common.cuh).
// =============================================================================================
namespace {

constexpr int TCF_THREADS = 224;
constexpr int TCF_STAGES = 4;
constexpr int TCF_F1 = 8;
constexpr int TCF_WROW = 512;   // floats per k-step block of packed weights: 64 rows x 8

__host__ __device__ inline int tcf_ksteps(int K1) { return (K1 + 3 + 7) / 8; }
__host__ __device__ inline int tcf_xs_len(int K1, int T) {
    int need = 4 * 127 + 8 * tcf_ksteps(K1);           // highest element any MMA row touches, + 1
    int have = (K1 - 1) / 2 + T;
    int n = need > have ? need : have;
    return (n + 3) / 4 * 4;
}

// Packs the Toeplitz weight operand of every model: out[m][ks][n/8][k/4][n%8][k%4], n = 4f + j (+32 for lo).
__global__ void tconv_wt_pack_kernel(const float *__restrict__ params, int64_t pstride, int64_t oW1, int K1,
                                     int ksteps, float *__restrict__ out) {
    const int m = blockIdx.y, ks = blockIdx.x, idx = threadIdx.x;      // 512 threads
    const int n = idx >> 3, kk = idx & 7;
    const int r = n & 31, f = r >> 2, j = r & 3;
    const int k = ks * 8 + kk - j;
    float w = (k >= 0 && k < K1) ? params[(int64_t)m * pstride + oW1 + f * K1 + k] : 0.f;
    float hi, lo;
    tc::split_tf32(w, hi, lo);
    out[((int64_t)m * ksteps + ks) * TCF_WROW + (n >> 3) * 64 + (kk >> 2) * 32 + (n & 7) * 4 + (kk & 3)] =
        n < 32 ? hi : lo;
}

__global__ void __launch_bounds__(TCF_THREADS, 2)
tconv_fwd_tc_kernel(const float *__restrict__ x, const int32_t *__restrict__ x_index,
                    const float *__restrict__ wt_packed, float *__restrict__ y1, float *__restrict__ part,
                    int n_rows, int B, int C, int T, int padl, int ksteps, int xs_len) {
    extern __shared__ __align__(128) float smem[];
    float *wt = smem;                                  // [ksteps][512]
    float *xbuf = smem + (size_t)ksteps * TCF_WROW;    // [STAGES][2][xs_len]
    __shared__ uint64_t bar_xfull[TCF_STAGES], bar_xempty[TCF_STAGES], bar_accfull[2], bar_accempty[2], bar_w,
        bar_wfree;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int row_lo = (int)((int64_t)n_rows * blockIdx.x / gridDim.x);
    const int row_hi = (int)((int64_t)n_rows * (blockIdx.x + 1) / gridDim.x);
    const int nrows = row_hi - row_lo;
    if (nrows <= 0) return;

    if (tid == 0) {
        for (int s = 0; s < TCF_STAGES; ++s) { tc::mbar_init(&bar_xfull[s], 1); tc::mbar_init(&bar_xempty[s], 1); }
        for (int b = 0; b < 2; ++b) { tc::mbar_init(&bar_accfull[b], 1); tc::mbar_init(&bar_accempty[b], 4); }
        tc::mbar_init(&bar_w, 1);
        tc::mbar_init(&bar_wfree, 1);
        tc::mbar_init_fence();
    }
    if (warp == 4) tc::tmem_alloc(&tmem_slot, 128);
    for (int i = tid; i < TCF_STAGES * 2 * xs_len; i += TCF_THREADS) xbuf[i] = 0.f;   // halos stay zero
    tc::fence_proxy_async_smem();
    tc::tc_fence_before_sync();
    __syncthreads();
    tc::tc_fence_after_sync();
    const uint32_t tmem = tmem_slot;

    if (warp >= 5) {
        // ---------------- producers: x row -> tf32 hi / lo halves in the ring ----------------
        for (int li = warp - 5; li < nrows; li += 2) {
            const int s = li % TCF_STAGES, u = li / TCF_STAGES;
            const int row = row_lo + li;
            const int n = row / C, c = row - n * C;
            const int64_t xrow = x_index ? (int64_t)x_index[n] : (int64_t)n;
            const float *src = x + (xrow * C + c) * (int64_t)T;
            float v[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                const int e = lane + 32 * q;
                v[q] = e < T ? __ldg(src + e) : 0.f;
            }
            if (u > 0) tc::mbar_wait(&bar_xempty[s], (u - 1) & 1);
            float *hi = xbuf + (size_t)s * 2 * xs_len + padl;
            float *lo = hi + xs_len;
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                const int e = lane + 32 * q;
                if (e < T) {
                    float h, l;
                    tc::split_tf32(v[q], h, l);
                    hi[e] = h;
                    lo[e] = l;
                }
            }
            tc::fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&bar_xfull[s]);
        }
    } else if (warp == 4) {
        // ---------------- MMA issuer ----------------
        const bool leader = tc::elect_one();
        const uint32_t idesc64 = tc::idesc_tf32(128, 64, 0, 0), idesc32 = tc::idesc_tf32(128, 32, 0, 0);
        const uint64_t bdesc0 = tc::smem_desc(tc::smem_u32(wt), 128, 256);
        const uint32_t wbytes = (uint32_t)ksteps * TCF_WROW * 4;
        int cur_m = -1;
        uint32_t w_phase = 0, wfree_phase = 0;
        for (int li = 0; li < nrows; ++li) {
            const int row = row_lo + li;
            const int m = (row / C) / B;
            if (m != cur_m) {
                if (cur_m >= 0) {      // every MMA that reads the old weights must have retired
                    if (leader) tc::mma_commit(&bar_wfree);
                    tc::mbar_wait(&bar_wfree, wfree_phase);
                    wfree_phase ^= 1;
                }
                if (leader) {
                    tc::mbar_expect_tx(&bar_w, wbytes);
                    const float *src = wt_packed + (int64_t)m * ksteps * TCF_WROW;
                    for (uint32_t off = 0; off < wbytes; off += 16384) {
                        const uint32_t nb = wbytes - off < 16384 ? wbytes - off : 16384;
                        tc::tma_load_1d(reinterpret_cast<char *>(wt) + off, reinterpret_cast<const char *>(src) + off,
                                        nb, &bar_w);
                    }
                }
                tc::mbar_wait(&bar_w, w_phase);
                w_phase ^= 1;
                cur_m = m;
            }
            const int s = li % TCF_STAGES, u = li / TCF_STAGES, b = li & 1, ub = li >> 1;
            tc::mbar_wait(&bar_xfull[s], u & 1);
            if (ub > 0) tc::mbar_wait(&bar_accempty[b], (ub - 1) & 1);
            tc::tc_fence_after_sync();
            if (leader) {
                const float *xhi = xbuf + (size_t)s * 2 * xs_len;
                uint64_t ahi = tc::smem_desc(tc::smem_u32(xhi), 16, 128);
                uint64_t alo = tc::smem_desc(tc::smem_u32(xhi + xs_len), 16, 128);
                uint64_t bd = bdesc0;
                const uint32_t d = tmem + b * 64;
                for (int ks = 0; ks < ksteps; ++ks) {
                    tc::mma_tf32_ss(d, ahi, bd, idesc64, ks > 0 ? 1u : 0u);
                    tc::mma_tf32_ss(d, alo, bd, idesc32, 1u);
                    ahi += 2;      // 32 B: the next 8 taps
                    alo += 2;
                    bd += (TCF_WROW * 4) >> 4;
                }
                tc::mma_commit(&bar_xempty[s]);
                tc::mma_commit(&bar_accfull[b]);
            }
            __syncwarp();
        }
    } else {
        // ---------------- epilogue: TMEM -> y1 (+ BatchNorm-1 partial sums) ----------------
        const int col = warp * 32 + lane, t0 = 4 * col;
        const bool valid = t0 < T;
        for (int li = 0; li < nrows; ++li) {
            const int b = li & 1, ub = li >> 1;
            tc::mbar_wait(&bar_accfull[b], ub & 1);
            tc::tc_fence_after_sync();
            float v0[32], v1[32];
            const uint32_t ta = tmem + ((uint32_t)(warp * 32) << 16) + b * 64;
            tc::tmem_ld32(ta, v0);
            tc::tmem_ld32(ta + 32, v1);
            tc::tmem_ld_wait();
            tc::tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&bar_accempty[b]);
            const int row = row_lo + li;
            const int n = row / C, c = row - n * C;
#pragma unroll
            for (int q = 0; q < 32; ++q) v0[q] += v1[q];
            if (valid) {
#pragma unroll
                for (int f = 0; f < TCF_F1; ++f) {
                    float *dst = y1 + (((int64_t)n * TCF_F1 + f) * C + c) * (int64_t)T + t0;
                    *reinterpret_cast<float4 *>(dst) = make_float4(v0[4 * f], v0[4 * f + 1], v0[4 * f + 2], v0[4 * f + 3]);
                }
            }
            if (part != nullptr) {
                float *prow = part + ((int64_t)row * 4 + warp) * (2 * TCF_F1);
#pragma unroll
                for (int f = 0; f < TCF_F1; ++f) {
                    float sm = 0.f, sq = 0.f;
                    if (valid) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            sm += v0[4 * f + j];
                            sq = fmaf(v0[4 * f + j], v0[4 * f + j], sq);
                        }
                    }
                    sm = warp_sum(sm);
                    sq = warp_sum(sq);
                    if (lane == 0) { prow[2 * f] = sm; prow[2 * f + 1] = sq; }
                }
            }
        }
    }
    tc::tc_fence_before_sync();
    __syncthreads();
    if (warp == 4) tc::tmem_dealloc(tmem, 128);
}

bool tc_env_enabled() { return tc_path_enabled("EAV_TCONV"); }

}  // namespace

// EAV_TC=ffma switches every tensor-core path off, <var>=ffma one of them (A/B runs, tests).  Read on every call:
// the choice also fixes the workspace layout, so it must not change between eav_eegnet_workspace_bytes and the
// launches of one engine (tests flip it between engines).
bool tc_path_enabled(const char *var) {
    const char *names[2] = {"EAV_TC", var};
    for (const char *nm : names) {
        const char *e = getenv(nm);
        if (e && (!strcmp(e, "ffma") || !strcmp(e, "0"))) return false;
    }
    return true;
}

bool tconv_fwd_use_tc(const NetDims &d) {
    if (!tc_env_enabled()) return false;
    if (d.F1 != TCF_F1 || (d.T & 3) || d.T > 512 || d.T < 4) return false;
    const size_t smem = ((size_t)tcf_ksteps(d.K1) * TCF_WROW + (size_t)TCF_STAGES * 2 * tcf_xs_len(d.K1, d.T)) * 4;
    return smem <= 110 * 1024;
}
size_t tconv_fwd_tc_scratch_floats(const NetDims &d) {
    return tconv_fwd_use_tc(d) ? (size_t)d.M * tcf_ksteps(d.K1) * TCF_WROW : 0;
}
int tconv_fwd_tc_rows_per_sample(const NetDims &d) { return d.C * 4; }

int launch_tconv_fwd_tc(const NetDims &d, const float *x, const int32_t *x_index, const float *params,
                        float *wt_scratch, float *y1, float *part, int *part_rows, cudaStream_t st) {
    EAV_REQUIRE(wt_scratch != nullptr, EAV_ERR_BAD_ARG, "tconv_fwd_tc: no weight scratch");
    const int ksteps = tcf_ksteps(d.K1), xs_len = tcf_xs_len(d.K1, d.T);
    tconv_wt_pack_kernel<<<dim3(ksteps, d.M), 512, 0, st>>>(params, d.pstride, d.oW1, d.K1, ksteps, wt_scratch);
    EAV_CUDA_LAUNCH_CHECK("tconv_wt_pack");
    const size_t smem = ((size_t)ksteps * TCF_WROW + (size_t)TCF_STAGES * 2 * xs_len) * 4;
    static PerDevice<bool> attr_set_pd(false);
    bool &attr_set = attr_set_pd.here();
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(tconv_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024);
        EAV_REQUIRE(e == cudaSuccess, (int)e, "tconv_fwd_tc: %s", cudaGetErrorString(e));
        attr_set = true;
    }
    const int sms = device_sm_count();
    const int n_rows = d.N * d.C;
    int grid = 2 * sms;
    if (grid > n_rows) grid = n_rows;
    tconv_fwd_tc_kernel<<<grid, TCF_THREADS, smem, st>>>(x, x_index, wt_scratch, y1, part, n_rows, d.B, d.C, d.T,
                                                         d.pad1l, ksteps, xs_len);
    EAV_CUDA_LAUNCH_CHECK("tconv_fwd_tc");
    if (part_rows) *part_rows = d.B * tconv_fwd_tc_rows_per_sample(d);
    return 0;
}

// =============================================================================================
// Weight gradient of the temporal convolution on the tensor cores.
//
//   dW[f][k] = sum_{rows (b,c)} sum_t dy[f][t] * xs[t + k + DELTA],     xs[PADL + t] = x[t], DELTA = PADL - padl
// With t = 32 q + j:   D'[i][(f,j)] = sum_rows sum_q xs[32 q + i] * dy[f][32 q + j]   and
//                      dW[f][k]     = sum_{j<32} D'[k + j + DELTA][(f,j)].
// Both operands are the RAW row buffers, MN-major with the 128B/32B-atom swizzle (the only MN-major mode tf32 has):
//   A[i][q] = xs[32 q + i]    -> LBO 128 B (next 32 i), SBO 512 B (next 4 q); M tile mt starts 512 B further
//   B[(f,j)][q] = dy[f][32q+j]-> LBO = row pitch 2048 B (next f), SBO 512 B
// and one k-step (8 q = 256 samples) advances both start addresses by 1024 B.  The swizzle is a function of the
// absolute shared-memory address (scripts/tc_decode.py), so the producers store element e at swz(base + 4 e).
// D' (M = up to 384 rows in 3 tiles, N = 128 = 4 filters x 32) stays in TMEM for all rows of a work unit
// (model, filter half, row range); the epilogue sums the 32 diagonals through shared memory.
// Roles: warps 0-3 epilogue, warp 4 MMA issue, warps 5-11 producers, one ring stage each (BatchNorm-1 backward is applied to
// dz1 while staging, then the tf32 hi/lo split).
// =============================================================================================
namespace {

// One ring stage per producer warp: a warp then waits on every phase of its own stage's `empty` barrier in
// order.  (With more warps than stages a warp could test a parity two phases ahead, which mbarrier parity
// waits cannot distinguish from "already complete".)
constexpr int TCW_PROD_WARPS = 7;
constexpr int TCW_THREADS = (5 + TCW_PROD_WARPS) * 32;
constexpr int TCW_STAGES = TCW_PROD_WARPS;
// Row g of a CTA uses ring stage (g + TCW_FIRST_STAGE) % TCW_STAGES.  Any rotation is the same protocol (same barriers, same
// phases, same memory); 3 instead of 0 only because `compute-sanitizer --tool synccheck` (12.9) reports "Barrier error: Missing
// init" on the empty-barrier of whichever stage is index 0 when the ring STARTS there, and nothing when it starts elsewhere.
// The report followed that barrier through three different shared addresses and was independent of the order of the
// mbarrier.init calls and of the ring memory slot (profiles/r2_sanitizer.txt): a property of the tool, not of the kernel.
constexpr int TCW_FIRST_STAGE = 3;
constexpr int TCW_XS = 896;                       // floats per x buffer (>= 32*15 + 384), 3584 B
constexpr int TCW_DYROW = 512;                    // floats per dy row
constexpr int TCW_STAGE_FLOATS = 2 * TCW_XS + 2 * 4 * TCW_DYROW;   // x hi, x lo, dy hi[4], dy lo[4] = 23552 B
constexpr int TCW_SP = 33;                        // pitch of the diagonal-sum staging

__device__ __forceinline__ uint32_t swz128_32(uint32_t a) { return a ^ (((a >> 7) & 3u) << 5); }
__device__ __forceinline__ void sts_f4(uint32_t addr, float4 v) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void split4(float4 v, float4 &h, float4 &l) {
    tc::split_tf32(v.x, h.x, l.x);
    tc::split_tf32(v.y, h.y, l.y);
    tc::split_tf32(v.z, h.z, l.z);
    tc::split_tf32(v.w, h.w, l.w);
}
__device__ __forceinline__ uint64_t desc_mn_sw(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return tc::smem_desc(saddr, lbo, sbo) | ((uint64_t)1 << 61);     // layout_type 1 = SWIZZLE_128B_BASE32B
}
// Barrier among the four epilogue warps only (the other warps keep running): an mbarrier with 128 arrivals per phase.
__device__ __forceinline__ void epi_sync(uint64_t *bar, uint32_t &phase) {
    tc::mbar_arrive(bar);
    tc::mbar_wait(bar, phase);
    phase ^= 1u;
}

__global__ void __launch_bounds__(TCW_THREADS, 1)
tconv_bwd_dw_tc_kernel(const float *__restrict__ x, const int32_t *__restrict__ x_index,
                       const float *__restrict__ dz1, const float *__restrict__ y1,
                       const float4 *__restrict__ bnf1, const float4 *__restrict__ bnb1, int bn_train, int M, int B,
                       int C, int T, int K1, int padl, int S, float *__restrict__ part) {
    extern __shared__ __align__(1024) float smem[];
    float *ring = smem;                                        // [STAGES][STAGE_FLOATS]
    float *diag = smem + TCW_STAGES * TCW_STAGE_FLOATS;        // [384][33]
    __shared__ uint64_t bar_full[TCW_STAGES], bar_empty[TCW_STAGES], bar_accfull, bar_accempty, bar_epi;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int F1 = 8;
    const int PADL = (padl + 3) & ~3, DELTA = PADL - padl;
    const int mtiles = (K1 + 31 + DELTA + 127) / 128;
    const int ksteps = (T + 255) / 256;
    const int rows_m = B * C;
    const int n_units = 2 * M * S;

    if (tid == 0) {
        for (int s = 0; s < TCW_STAGES; ++s) { tc::mbar_init(&bar_full[s], 1); tc::mbar_init(&bar_empty[s], 1); }
        tc::mbar_init(&bar_accfull, 1);
        tc::mbar_init(&bar_accempty, 4);
        tc::mbar_init(&bar_epi, 128);
        tc::mbar_init_fence();
    }
    if (warp == 4) tc::tmem_alloc(&tmem_slot, 512);
    for (int i = tid; i < TCW_STAGES * TCW_STAGE_FLOATS; i += TCW_THREADS) ring[i] = 0.f;   // halos / tails stay zero
    tc::fence_proxy_async_smem();
    tc::tc_fence_before_sync();
    __syncthreads();
    tc::tc_fence_after_sync();
    const uint32_t tmem = tmem_slot;

    auto unit_rows = [&](int u, int &m, int &fg, int &r_lo, int &r_hi) {
        fg = u & 1;
        const int ms = u >> 1;
        m = ms / S;
        const int sp = ms - m * S;
        r_lo = (int)((int64_t)rows_m * sp / S);
        r_hi = (int)((int64_t)rows_m * (sp + 1) / S);
    };

    if (warp >= 5) {
        // ---------------- producers ----------------
        const int pw = warp - 5;
        int g = 0;                                     // running row counter of this CTA (all units)
        for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
            int m, fg, r_lo, r_hi;
            unit_rows(u, m, fg, r_lo, r_hi);
            for (int r = r_lo; r < r_hi; ++r, ++g) {
                if (g % TCW_PROD_WARPS != pw) continue;
                const int s = (g + TCW_FIRST_STAGE) % TCW_STAGES, use = g / TCW_STAGES;
                const int b = r / C, c = r - b * C;
                const int64_t n = (int64_t)m * B + b;
                const int64_t xrow = x_index ? (int64_t)x_index[n] : n;
                const float *xsrc = x + (xrow * C + c) * (int64_t)T;
                float4 xv[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int t = 4 * (lane + 32 * q);
                    xv[q] = t < T ? *reinterpret_cast<const float4 *>(xsrc + t) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
                if (use > 0) tc::mbar_wait(&bar_empty[s], (use - 1) & 1);
                const uint32_t sbase = tc::smem_u32(ring + (size_t)s * TCW_STAGE_FLOATS);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int t = 4 * (lane + 32 * q);
                    if (t < T) {
                        float4 h, l;
                        split4(xv[q], h, l);
                        const uint32_t a = sbase + 4u * (uint32_t)(PADL + t);
                        sts_f4(swz128_32(a), h);
                        sts_f4(swz128_32(a + 4u * TCW_XS), l);
                    }
                }
                const uint32_t dybase = sbase + 4u * 2 * TCW_XS;
#pragma unroll
                for (int fp = 0; fp < 4; fp += 2) {      // two filters at a time bounds the registers in flight
                    float4 dzv[2][4], yv[2][4];
#pragma unroll
                    for (int h2 = 0; h2 < 2; ++h2) {
                        const int f = fg * 4 + fp + h2;
                        const int64_t base = ((n * F1 + f) * C + c) * (int64_t)T;
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const int t = 4 * (lane + 32 * q);
                            const bool act = t < T;
                            dzv[h2][q] = act ? *reinterpret_cast<const float4 *>(dz1 + base + t) : make_float4(0.f, 0.f, 0.f, 0.f);
                            if (bn_train)
                                yv[h2][q] = act ? *reinterpret_cast<const float4 *>(y1 + base + t) : make_float4(0.f, 0.f, 0.f, 0.f);
                        }
                    }
#pragma unroll
                    for (int h2 = 0; h2 < 2; ++h2) {
                        const int f = fg * 4 + fp + h2;
                        const float4 kb = bnb1[(int64_t)m * F1 + f];
                        const float4 kf = bnf1[(int64_t)m * F1 + f];
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const int t = 4 * (lane + 32 * q);
                            if (t < T) {
                                float4 o;
                                if (bn_train) {
                                    o.x = kb.x * (dzv[h2][q].x - kb.y - (yv[h2][q].x - kf.x) * kf.y * kb.z);
                                    o.y = kb.x * (dzv[h2][q].y - kb.y - (yv[h2][q].y - kf.x) * kf.y * kb.z);
                                    o.z = kb.x * (dzv[h2][q].z - kb.y - (yv[h2][q].z - kf.x) * kf.y * kb.z);
                                    o.w = kb.x * (dzv[h2][q].w - kb.y - (yv[h2][q].w - kf.x) * kf.y * kb.z);
                                } else {
                                    o = make_float4(kb.x * dzv[h2][q].x, kb.x * dzv[h2][q].y, kb.x * dzv[h2][q].z, kb.x * dzv[h2][q].w);
                                }
                                float4 h, l;
                                split4(o, h, l);
                                const uint32_t a = dybase + 4u * (uint32_t)((fp + h2) * TCW_DYROW + t);
                                sts_f4(swz128_32(a), h);
                                sts_f4(swz128_32(a + 4u * 4 * TCW_DYROW), l);
                            }
                        }
                    }
                }
                tc::fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&bar_full[s]);
            }
        }
    } else if (warp == 4) {
        // ---------------- MMA issuer ----------------
        const bool leader = tc::elect_one();
        const uint32_t idesc = tc::idesc_tf32(128, 128, 1, 1);
        int g = 0, nu = 0;
        for (int u = blockIdx.x; u < n_units; u += gridDim.x, ++nu) {
            int m, fg, r_lo, r_hi;
            unit_rows(u, m, fg, r_lo, r_hi);
            if (nu > 0) tc::mbar_wait(&bar_accempty, (nu - 1) & 1);     // epilogue drained the previous unit
            for (int r = r_lo; r < r_hi; ++r, ++g) {
                const int s = (g + TCW_FIRST_STAGE) % TCW_STAGES, use = g / TCW_STAGES;
                tc::mbar_wait(&bar_full[s], use & 1);
                tc::tc_fence_after_sync();
                if (leader) {
                    const uint32_t sbase = tc::smem_u32(ring + (size_t)s * TCW_STAGE_FLOATS);
                    const uint32_t xhi = sbase, xlo = sbase + 4u * TCW_XS;
                    const uint32_t dyhi = sbase + 4u * 2 * TCW_XS, dylo = dyhi + 4u * 4 * TCW_DYROW;
                    const uint32_t first = (r == r_lo) ? 0u : 1u;
                    for (int mt = 0; mt < mtiles; ++mt) {
                        const uint32_t d = tmem + mt * 128;
                        for (int ks = 0; ks < ksteps; ++ks) {
                            const uint32_t ao = mt * 512 + ks * 1024, bo = ks * 1024;
                            const uint64_t ah = desc_mn_sw(xhi + ao, 128, 512), al = desc_mn_sw(xlo + ao, 128, 512);
                            const uint64_t bh = desc_mn_sw(dyhi + bo, 4 * TCW_DYROW, 512);
                            const uint64_t bl = desc_mn_sw(dylo + bo, 4 * TCW_DYROW, 512);
                            tc::mma_tf32_ss(d, ah, bh, idesc, (ks > 0) ? 1u : first);
                            tc::mma_tf32_ss(d, ah, bl, idesc, 1u);
                            tc::mma_tf32_ss(d, al, bh, idesc, 1u);
                        }
                    }
                    tc::mma_commit(&bar_empty[s]);
                    if (r == r_hi - 1) tc::mma_commit(&bar_accfull);
                }
                __syncwarp();
            }
        }
    } else {
        // ---------------- epilogue: D' -> diagonal sums -> partial dW ----------------
        int nu = 0;
        uint32_t epi_phase = 0;
        for (int u = blockIdx.x; u < n_units; u += gridDim.x, ++nu) {
            int m, fg, r_lo, r_hi;
            unit_rows(u, m, fg, r_lo, r_hi);
            const int sp = (u >> 1) - m * S;
            tc::mbar_wait(&bar_accfull, nu & 1);
            tc::tc_fence_after_sync();
            for (int fl = 0; fl < 4; ++fl) {
                for (int mt = 0; mt < mtiles; ++mt) {
                    float v[32];
                    tc::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + mt * 128 + fl * 32, v);
                    tc::tmem_ld_wait();
                    float *drow = diag + (size_t)(mt * 128 + warp * 32 + lane) * TCW_SP;
#pragma unroll
                    for (int j = 0; j < 32; ++j) drow[j] = v[j];
                }
                if (fl == 3) {           // all TMEM reads of this unit are done: the next unit may accumulate
                    tc::tc_fence_before_sync();
                    __syncwarp();
                    if (lane == 0) tc::mbar_arrive(&bar_accempty);
                }
                epi_sync(&bar_epi, epi_phase);
                float *dst = part + (((int64_t)m * S + sp) * F1 + fg * 4 + fl) * K1;
                for (int k = tid; k < K1; k += 128) {
                    float acc = 0.f;
                    const float *p = diag + (size_t)(k + DELTA) * TCW_SP;
#pragma unroll
                    for (int j = 0; j < 32; ++j) acc += p[j * (TCW_SP + 1)];
                    dst[k] = acc;
                }
                epi_sync(&bar_epi, epi_phase);
            }
        }
    }
    tc::tc_fence_before_sync();
    __syncthreads();
    if (warp == 4) tc::tmem_dealloc(tmem, 512);
}

}  // namespace

// =============================================================================================
// Eval-mode fused backward of block 1 (EEGNet_tor.py:51-57 backward; SURVEY 8a rows M1-M4, section 7: "[dz1] can be
// regenerated in-register inside the dW1 kernel").  Same MMA formulation as tconv_bwd_dw_tc_kernel, but the producers
// no longer READ the (N,8,30,500) gradient dz1 that dw_bwd used to write: they rebuild it from the 16x smaller dz2 and
// the saved conv output y1 while staging the MMA operand,
//     dz1[f][c][t] = E'(pre) * sum_d k2[g] W2[g][c] dz2[g][t],   g = 8 f + d,  pre = y1*scale1[f] + shift1[f],
// and fold in everything else dw_bwd did in the same pass over y1:
//     dW2[g][c] = k2[g] * sum_{b,t} dz2[g][t] * act(pre),    d(beta1)[f] = sum dz1,   d(gamma1)[f] = sum dz1 * xhat1.
// Eval-mode BatchNorm only (the reference's steady state): BN backward is then the per-channel scale
// k = gamma*invstd with no batch sums, so nothing has to be known before this kernel starts.  k1[f] is applied to the
// dW1 slab in the epilogue.  This deletes one 645 MB write, one 645 MB read and the dw_bwd launch per step.
//
// Producers: 8 warps = 4 time quarters (lane owns 4 consecutive t) x 2 filter pairs, ALL on the same row; a thread
// keeps dz2 of its 2 filters x 8 depth channels x 4 time steps in 64 registers for the 30 rows of a sample, so the
// per-row inputs are one 128-bit load of x and two of y1.  The 16 dW2 accumulators of a row are summed across the
// warp with a transposing butterfly (16 shuffles) and added to a per-(quarter, electrode) shared-memory slot owned
// by that warp: every sum has a fixed order (deterministic, like all other reductions of the library).
// =============================================================================================
namespace {

constexpr int TCX_PROD_WARPS = 8;
constexpr int TCX_THREADS = 16 * 32;     // 4 epilogue + 8 producer + MMA + loader warp + 2 idle warps completing the last warpgroup
constexpr int TCX_STAGES = 5;
constexpr int TCX_CMAX = 32;                               // electrodes the shared-memory tables hold
constexpr int TCX_NB = 4;                                  // input row buffers (TMA prefetch depth)
constexpr int TCX_INROW = 512;                             // floats per staged row (x, y1 of 4 filters): 5 rows per buffer
constexpr size_t TCX_SMEM_FLOATS = (size_t)TCX_STAGES * TCW_STAGE_FLOATS + 384 * TCW_SP + TCX_CMAX * 32 +
                                   4 * TCX_CMAX * 32 + TCX_PROD_WARPS * 4 + TCX_NB * 5 * TCX_INROW;

__device__ __forceinline__ void producers_sync() { asm volatile("bar.sync 1, %0;" ::"n"(TCX_PROD_WARPS * 32) : "memory"); }

// Sums v[0..15] over the 32 lanes; afterwards lane L holds the total of index 8*b16 + 4*b8 + 2*b4 + b2 (b = bits of L).
__device__ __forceinline__ float transpose_reduce16(float (&v)[16], int lane) {
#define EAV_TR_STEP(HALF, BIT)                                                         \
    {                                                                                  \
        const bool upper = (lane & BIT) != 0;                                          \
        _Pragma("unroll") for (int i = 0; i < HALF; ++i) {                             \
            const float send = upper ? v[i] : v[i + HALF];                             \
            const float keep = upper ? v[i + HALF] : v[i];                             \
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, BIT);                     \
        }                                                                              \
    }
    EAV_TR_STEP(8, 16)
    EAV_TR_STEP(4, 8)
    EAV_TR_STEP(2, 4)
    EAV_TR_STEP(1, 2)
#undef EAV_TR_STEP
    return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 1);
}

struct FusedBwdArgs {
    const float *dz2;          // [N][G][T]   gradient w.r.t. the BatchNorm-2 output
    const float *y1;           // [N][F1][C][T] saved raw conv output
    const float4 *bnf1, *bnf2; // {mean, invstd, scale, shift} per (model, channel)
    const float *params;
    int64_t pstride, oW2, og1, og2;
    float *part;               // dW1 partial slabs [M][S][F1][K1]
    float *partw2;             // dW2 partial sums  [M*S][2 fg][4 quarters][TCX_CMAX][32]
    float *partbn;             // BatchNorm-1 backward sums [M][S][F1][2]
    int elu1;                  // 1: ELU after BN1 (EEGNet_tor), 0: none (CNN_EEG)
};

// Register budget: the producers need ~150 registers (64 for the resident dz2 alone), the epilogue and the MMA warp
// far fewer, and 13 warps x 150 do not fit an SM (a sub-partition owns 16 384 registers and every fourth warp).
// Roles therefore sit on warpgroup boundaries -- warps 0-3 epilogue, 4-11 producers, 12 MMA issue, 13 TMA loader,
// 14-15 idle (setmaxnreg.sync.aligned is a warpgroup-wide instruction: a partial last group hangs the inc) -- and
// are rebalanced with setmaxnreg: the CTA launches with 512 x 128 registers, the epilogue drops to 80, the last
// group to 40, the producers grow to 184.
//
// (A 16-producer-warp variant -- one filter per warp, 96 registers -- was measured too: it raises the issue rate from
// 43 % to 68 % but pays the per-row waits / fences / reductions sixteen times instead of eight and came out equal,
// 0.68 ms against 0.66 ms; the eight-warp form is kept.)
//
// Inputs: the loader warp streams each row's x and the four y1 rows (10 KB) into a 3-deep shared-memory ring with
// cp.async.bulk (TMA) two rows ahead; the producers only wait on an mbarrier and read 128-bit values from shared
// memory.  (Register prefetch with plain loads did not work: the six hardware load scoreboards are shared with the
// per-sample dz2 reload, ncu showed 31 % of the producers' time in long-scoreboard stalls on the first FFMA of a row,
// and the per-thread 64-bit address arithmetic was ~15 % of the row's instructions.)
__global__ void __launch_bounds__(TCX_THREADS, 1)
tconv_bwd_fused_tc_kernel(const float *__restrict__ x, const int32_t *__restrict__ x_index, const FusedBwdArgs a,
                          int M, int B, int C, int T, int K1, int padl, int S) {
    extern __shared__ __align__(1024) float smem[];
    float *ring = smem;                                            // [STAGES][STAGE_FLOATS]
    float *diag = smem + TCX_STAGES * TCW_STAGE_FLOATS;            // [384][33]
    float *vtab = diag + 384 * TCW_SP;                             // [CMAX][32]  k2[g] * W2[g][c], g = 32 fg + j
    float *dw2s = vtab + TCX_CMAX * 32;                            // [4][CMAX][32]
    float *bnred = dw2s + 4 * TCX_CMAX * 32;                       // [8][4]
    float *inbuf = bnred + TCX_PROD_WARPS * 4;                     // [NB][5][INROW]: x, y1[f0..f3] of one row
    __shared__ uint64_t bar_full[TCX_STAGES], bar_empty[TCX_STAGES], bar_accfull, bar_accempty, bar_epi;
    __shared__ uint64_t bar_in_full[TCX_NB], bar_in_empty[TCX_NB];
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int F1 = 8, G = 64;
    const int PADL = (padl + 3) & ~3, DELTA = PADL - padl;
    const int mtiles = (K1 + 31 + DELTA + 127) / 128;
    const int ksteps = (T + 255) / 256;
    const int rows_m = B * C;
    const int n_units = 2 * M * S;

    if (tid == 0) {
        for (int s = 0; s < TCX_STAGES; ++s) { tc::mbar_init(&bar_full[s], TCX_PROD_WARPS); tc::mbar_init(&bar_empty[s], 1); }
        for (int s = 0; s < TCX_NB; ++s) { tc::mbar_init(&bar_in_full[s], 1); tc::mbar_init(&bar_in_empty[s], TCX_PROD_WARPS); }
        tc::mbar_init(&bar_accfull, 1);
        tc::mbar_init(&bar_accempty, 4);
        tc::mbar_init(&bar_epi, 128);
        tc::mbar_init_fence();
    }
    if (warp == 12) tc::tmem_alloc(&tmem_slot, 512);
    for (int i = tid; i < TCX_STAGES * TCW_STAGE_FLOATS; i += TCX_THREADS) ring[i] = 0.f;   // halos / tails stay zero
    for (int i = tid; i < 4 * TCX_CMAX * 32; i += TCX_THREADS) dw2s[i] = 0.f;
    for (int i = tid; i < TCX_NB * 5 * TCX_INROW; i += TCX_THREADS) inbuf[i] = 0.f;         // [T, INROW) is never copied to
    tc::fence_proxy_async_smem();
    tc::tc_fence_before_sync();
    __syncthreads();
    tc::tc_fence_after_sync();
    const uint32_t tmem = tmem_slot;

    auto unit_rows = [&](int u, int &m, int &fg, int &r_lo, int &r_hi) {
        fg = u & 1;
        const int ms = u >> 1;
        m = ms / S;
        const int sp = ms - m * S;
        r_lo = (int)((int64_t)rows_m * sp / S);
        r_hi = (int)((int64_t)rows_m * (sp + 1) / S);
    };

    if (warp >= 4 && warp < 12) {
        // ---------------- producers ----------------
        asm volatile("setmaxnreg.inc.sync.aligned.u32 184;" ::: "memory");
        const int pw = warp - 4, q = pw & 3, fp = pw >> 2, ptid = tid - 4 * 32;
        const int t = 4 * (lane + 32 * q);
        const bool tact = t < T;
        float4 dzr[2][8];
        int64_t cur_n = -1;
        int s = 0, sfull = 0;          // ring stage and whether it has been used before / parity of its last use
        uint32_t sphase = 0;
        int ib = 0;
        uint32_t iphase = 0;
        for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
            int m, fg, r_lo, r_hi;
            unit_rows(u, m, fg, r_lo, r_hi);
            const float *prm = a.params + (int64_t)m * a.pstride;
            producers_sync();                                    // everybody is done with the previous unit's table
            for (int i = ptid; i < TCX_CMAX * 32; i += TCX_PROD_WARPS * 32) {
                const int c = i >> 5, j = i & 31, gch = fg * 32 + j;
                float v = 0.f;
                if (c < C) v = prm[a.og2 + gch] * a.bnf2[(int64_t)m * G + gch].y * prm[a.oW2 + (int64_t)gch * C + c];
                vtab[i] = v;
            }
            producers_sync();
            float sc1[2], sh1[2], is1[2], nmi[2];                // BN1 scale, shift, invstd, -mean*invstd
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const float4 st = a.bnf1[(int64_t)m * F1 + fg * 4 + fp * 2 + h];
                sc1[h] = st.z; sh1[h] = st.w; is1[h] = st.y; nmi[h] = -st.x * st.y;
            }
            // BatchNorm-1 backward sums as packed pairs (even / odd time steps), folded at the end of the unit
            float2 p1p[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)}, p2p[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
            int b = r_lo / C, c = r_lo - b * C;
            // The cross-lane sum of a row's 16 dW2 accumulators is a chain of 5 dependent shuffles; it is issued one
            // row late, in the same basic block as the next row's FFMA work, so its latency hides under that work.
            float accp[16];
            int cprev = -1;
            auto flush_dw2 = [&]() {
                const float tot = transpose_reduce16(accp, lane);
                if ((lane & 1) == 0) {
                    const int idx = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
                    float *slot = dw2s + (q * TCX_CMAX + cprev) * 32 + fp * 16 + idx;
                    *slot += tot;
                }
            };
            for (int r = r_lo; r < r_hi; ++r) {
                const int64_t n = (int64_t)m * B + b;
                if (n != cur_n) {                                // a new sample: its dz2 stays in registers for C rows
                    cur_n = n;
#pragma unroll
                    for (int h = 0; h < 2; ++h)
#pragma unroll
                        for (int dd = 0; dd < 8; ++dd) {
                            const int gch = (fg * 4 + fp * 2 + h) * 8 + dd;
                            dzr[h][dd] = tact ? *reinterpret_cast<const float4 *>(a.dz2 + (n * G + gch) * (int64_t)T + t)
                                              : make_float4(0.f, 0.f, 0.f, 0.f);
                        }
                }
                // this row's inputs, staged by the loader warp
                tc::mbar_wait(&bar_in_full[ib], iphase);
                const float *in = inbuf + (size_t)ib * 5 * TCX_INROW;
                float4 xv = make_float4(0.f, 0.f, 0.f, 0.f), yv[2];
                if (fp == 0) xv = *reinterpret_cast<const float4 *>(in + t);
#pragma unroll
                for (int h = 0; h < 2; ++h) yv[h] = *reinterpret_cast<const float4 *>(in + (1 + fp * 2 + h) * TCX_INROW + t);
                float acc[16];
                float4 outv[2];
                if (cprev >= 0) flush_dw2();                                // previous row's dW2 sums (see above)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const float4 v0 = *reinterpret_cast<const float4 *>(vtab + c * 32 + fp * 16 + h * 8);
                    const float4 v1 = *reinterpret_cast<const float4 *>(vtab + c * 32 + fp * 16 + h * 8 + 4);
                    const float vv[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
                    // Packed fp32 (Blackwell FFMA2 / FMUL2 / FADD2 on 64-bit register pairs): the four time steps of the
                    // thread are two (even, odd) pairs, which is how dz2 already sits in its float4 registers -- the 16 FMAs
                    // per element of the two small contractions become 8 instructions.
                    const float2 y01 = make_float2(yv[h].x, yv[h].y), y23 = make_float2(yv[h].z, yv[h].w);
                    const float2 sc2 = make_float2(sc1[h], sc1[h]), sh2 = make_float2(sh1[h], sh1[h]);
                    const float2 pre01 = __ffma2_rn(y01, sc2, sh2), pre23 = __ffma2_rn(y23, sc2, sh2);
                    const float pre[4] = {pre01.x, pre01.y, pre23.x, pre23.y};
                    float ep[4], act[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        ep[e] = 1.f;
                        act[e] = pre[e];
                        if (a.elu1) {
                            // exp(pre) as ONE multiply + MUFU.EX2: the .ftz form needs no denormal pre-scale / post-square
                            // (4 more instructions per element with __expf); results below 2^-126 flush to 0, harmless
                            // for exp(x) - 1 and for a derivative that multiplies a gradient.
                            float ex;
                            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex) : "f"(pre[e] * 1.4426950408889634f));
                            const bool pos = pre[e] > 0.f;
                            ep[e] = pos ? 1.f : ex;
                            act[e] = pos ? pre[e] : ex - 1.f;
                        }
                    }
                    const float2 act01 = make_float2(act[0], act[1]), act23 = make_float2(act[2], act[3]);
                    float2 uu01 = make_float2(0.f, 0.f), uu23 = make_float2(0.f, 0.f);
#pragma unroll
                    for (int dd = 0; dd < 8; ++dd) {
                        const float2 d01 = make_float2(dzr[h][dd].x, dzr[h][dd].y), d23 = make_float2(dzr[h][dd].z, dzr[h][dd].w);
                        const float2 w2 = make_float2(vv[dd], vv[dd]);
                        uu01 = __ffma2_rn(w2, d01, uu01);
                        uu23 = __ffma2_rn(w2, d23, uu23);
                        const float2 a2 = __ffma2_rn(d23, act23, __fmul2_rn(d01, act01));
                        acc[h * 8 + dd] = a2.x + a2.y;
                    }
                    const float2 dz01 = __fmul2_rn(uu01, make_float2(ep[0], ep[1])), dz23 = __fmul2_rn(uu23, make_float2(ep[2], ep[3]));
                    const float2 is2 = make_float2(is1[h], is1[h]), nm2 = make_float2(nmi[h], nmi[h]);
                    p1p[h] = __fadd2_rn(__fadd2_rn(p1p[h], dz01), dz23);
                    p2p[h] = __ffma2_rn(dz01, __ffma2_rn(y01, is2, nm2), p2p[h]);
                    p2p[h] = __ffma2_rn(dz23, __ffma2_rn(y23, is2, nm2), p2p[h]);
                    outv[h] = make_float4(dz01.x, dz01.y, dz23.x, dz23.y);
                }
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&bar_in_empty[ib]);          // the loader may refill this input buffer
                if (++ib == TCX_NB) { ib = 0; iphase ^= 1u; }
                if (sfull) tc::mbar_wait(&bar_empty[s], sphase ^ 1u);       // the MMAs of the stage's previous use retired
                const uint32_t sbase = tc::smem_u32(ring + (size_t)s * TCW_STAGE_FLOATS);
                if (tact) {
                    if (fp == 0) {
                        float4 h4, l4;
                        split4(xv, h4, l4);
                        const uint32_t ad = sbase + 4u * (uint32_t)(PADL + t);
                        sts_f4(swz128_32(ad), h4);
                        sts_f4(swz128_32(ad + 4u * TCW_XS), l4);
                    }
                    const uint32_t dybase = sbase + 4u * 2 * TCW_XS;
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        float4 h4, l4;
                        split4(outv[h], h4, l4);
                        const uint32_t ad = dybase + 4u * (uint32_t)((fp * 2 + h) * TCW_DYROW + t);
                        sts_f4(swz128_32(ad), h4);
                        sts_f4(swz128_32(ad + 4u * 4 * TCW_DYROW), l4);
                    }
                }
                tc::fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&bar_full[s]);
                if (++s == TCX_STAGES) { s = 0; sfull = 1; sphase ^= 1u; }
                // dW2 of this row: 16 (filter, depth) sums over the warp's 128 time steps, reduced during the next row
#pragma unroll
                for (int i = 0; i < 16; ++i) accp[i] = acc[i];
                cprev = c;
                if (++c == C) { c = 0; ++b; }
            }
            if (cprev >= 0) flush_dw2();
            // ---- end of unit: dW2 slots and BatchNorm-1 sums -> global partials (fixed order)
#pragma unroll
            float p1[2], p2[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) { p1[h] = warp_sum(p1p[h].x + p1p[h].y); p2[h] = warp_sum(p2p[h].x + p2p[h].y); }
            if (lane == 0) { bnred[pw * 4 + 0] = p1[0]; bnred[pw * 4 + 1] = p2[0]; bnred[pw * 4 + 2] = p1[1]; bnred[pw * 4 + 3] = p2[1]; }
            producers_sync();
            const int sp = (u >> 1) - m * S;
            float *pw2 = a.partw2 + (((int64_t)(m * S + sp) * 2 + fg) * 4) * TCX_CMAX * 32;
            for (int i = ptid; i < 4 * TCX_CMAX * 32; i += TCX_PROD_WARPS * 32) {
                pw2[i] = dw2s[i];
                dw2s[i] = 0.f;
            }
            if (ptid < 8) {                                      // (fp, h, which) -> filter fg*4 + fp*2 + h
                const int fp2 = ptid >> 2, hw = ptid & 3;
                float sum = 0.f;
#pragma unroll
                for (int qq = 0; qq < 4; ++qq) sum += bnred[(fp2 * 4 + qq) * 4 + hw];
                const int f = fg * 4 + fp2 * 2 + (hw >> 1);
                a.partbn[(((int64_t)m * S + sp) * F1 + f) * 2 + (hw & 1)] = sum;
            }
        }
    } else if (warp >= 12) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;" ::: "memory");
        if (warp == 12) {
            // ---------------- MMA issuer (as in tconv_bwd_dw_tc_kernel) ----------------
            const bool leader = tc::elect_one();
            const uint32_t idesc = tc::idesc_tf32(128, 128, 1, 1);
            const uint64_t xdesc0 = desc_mn_sw(tc::smem_u32(ring), 128, 512);                            // stage 0: x hi
            const uint64_t ddesc0 = desc_mn_sw(tc::smem_u32(ring) + 4u * 2 * TCW_XS, 4 * TCW_DYROW, 512);   // stage 0: dz1 hi
            int g = 0, nu = 0;
            for (int u = blockIdx.x; u < n_units; u += gridDim.x, ++nu) {
                int m, fg, r_lo, r_hi;
                unit_rows(u, m, fg, r_lo, r_hi);
                if (nu > 0) tc::mbar_wait_sleep(&bar_accempty, (nu - 1) & 1, 200);
                for (int r = r_lo; r < r_hi; ++r, ++g) {
                    const int s = g % TCX_STAGES, use = g / TCX_STAGES;
                    tc::mbar_wait_sleep(&bar_full[s], use & 1, 40);
                    tc::tc_fence_after_sync();
                    if (leader) {
                        // descriptors = hoisted stage-0 descriptors + address-field offsets (16-byte units): this warp shares a
                        // scheduler with two producer warps, every instruction it does not issue is a slot for them
                        const uint64_t so = (uint64_t)((uint32_t)s * ((TCW_STAGE_FLOATS * 4) >> 4));
                        const uint64_t xh = xdesc0 + so, xl = xh + ((4 * TCW_XS) >> 4);
                        const uint64_t dh = ddesc0 + so, dl = dh + ((4 * 4 * TCW_DYROW) >> 4);
                        const uint32_t first = (r == r_lo) ? 0u : 1u;
                        for (int mt = 0; mt < mtiles; ++mt) {
                            const uint32_t d = tmem + mt * 128;
                            for (int ks = 0; ks < ksteps; ++ks) {
                                const uint32_t ao = mt * 32 + ks * 64, bo = ks * 64;
                                tc::mma_tf32_ss(d, xh + ao, dh + bo, idesc, (ks > 0) ? 1u : first);
                                tc::mma_tf32_ss(d, xh + ao, dl + bo, idesc, 1u);
                                tc::mma_tf32_ss(d, xl + ao, dh + bo, idesc, 1u);
                            }
                        }
                        tc::mma_commit(&bar_empty[s]);
                        if (r == r_hi - 1) tc::mma_commit(&bar_accfull);
                    }
                    __syncwarp();
                }
            }
        } else if (warp == 13 && lane == 0) {
            // ---------------- loader: x row + the four y1 rows of (n, c) -> input ring, TMA bulk copies ----------------
            const uint32_t row_bytes = (uint32_t)T * 4u;
            int buf = 0, wrapped = 0;
            uint32_t bphase = 0;
            for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
                int m, fg, r_lo, r_hi;
                unit_rows(u, m, fg, r_lo, r_hi);
                int b = r_lo / C, c = r_lo - b * C;
                int64_t n = (int64_t)m * B + b;
                int64_t xrow = x_index ? (int64_t)x_index[n] : n;        // one dependent load per SAMPLE, not per row
                for (int r = r_lo; r < r_hi; ++r) {
                    if (wrapped) tc::mbar_wait(&bar_in_empty[buf], bphase ^ 1u);
                    float *dst = inbuf + (size_t)buf * 5 * TCX_INROW;
                    tc::mbar_expect_tx(&bar_in_full[buf], 5u * row_bytes);
                    tc::tma_load_1d(dst, x + (xrow * C + c) * (int64_t)T, row_bytes, &bar_in_full[buf]);
#pragma unroll
                    for (int fl = 0; fl < 4; ++fl)
                        tc::tma_load_1d(dst + (1 + fl) * TCX_INROW, a.y1 + ((n * F1 + fg * 4 + fl) * C + c) * (int64_t)T,
                                        row_bytes, &bar_in_full[buf]);
                    if (++buf == TCX_NB) { buf = 0; wrapped = 1; bphase ^= 1u; }
                    if (++c == C) {
                        c = 0; ++b; ++n;
                        if (r + 1 < r_hi) xrow = x_index ? (int64_t)x_index[n] : n;
                    }
                }
            }
        }
    } else {
        // ---------------- epilogue: D' -> diagonal sums -> k1[f] * partial dW1 ----------------
        asm volatile("setmaxnreg.dec.sync.aligned.u32 80;" ::: "memory");
        int nu = 0;
        uint32_t epi_phase = 0;
        for (int u = blockIdx.x; u < n_units; u += gridDim.x, ++nu) {
            int m, fg, r_lo, r_hi;
            unit_rows(u, m, fg, r_lo, r_hi);
            const int sp = (u >> 1) - m * S;
            tc::mbar_wait_sleep(&bar_accfull, nu & 1, 1000);
            tc::tc_fence_after_sync();
            for (int fl = 0; fl < 4; ++fl) {
                for (int mt = 0; mt < mtiles; ++mt) {
                    float v[32];
                    tc::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + mt * 128 + fl * 32, v);
                    tc::tmem_ld_wait();
                    float *drow = diag + (size_t)(mt * 128 + warp * 32 + lane) * TCW_SP;
#pragma unroll
                    for (int j = 0; j < 32; ++j) drow[j] = v[j];
                }
                if (fl == 3) {
                    tc::tc_fence_before_sync();
                    __syncwarp();
                    if (lane == 0) tc::mbar_arrive(&bar_accempty);
                }
                epi_sync(&bar_epi, epi_phase);
                const int f = fg * 4 + fl;
                const float k1 = a.params[(int64_t)m * a.pstride + a.og1 + f] * a.bnf1[(int64_t)m * F1 + f].y;
                float *dst = a.part + (((int64_t)m * S + sp) * F1 + f) * K1;
                for (int k = tid; k < K1; k += 128) {
                    float acc = 0.f;
                    const float *p = diag + (size_t)(k + DELTA) * TCW_SP;
#pragma unroll
                    for (int j = 0; j < 32; ++j) acc += p[j * (TCW_SP + 1)];
                    dst[k] = acc * k1;
                }
                epi_sync(&bar_epi, epi_phase);
            }
        }
    }
    tc::tc_fence_before_sync();
    __syncthreads();
    if (warp == 12) tc::tmem_dealloc(tmem, 512);
}

// One launch for the tail of the fused backward (three small kernels before: at 5-6 models per GPU a launch costs as much
// as the work): dW2[m][g][c] = k2[g] * sum over the model's units (in order) and the four time quarters of the per-unit
// slots; dW1[m][f][k] = sum of the S partial slabs; d(gamma1) / d(beta1) from the S partial BatchNorm sums.
__global__ void __launch_bounds__(256)
block1_bwd_finalize_kernel(const float *__restrict__ partw2, const float *__restrict__ partw1, const float *__restrict__ partbn,
                           int S, int C, int K1, const float *__restrict__ params, int64_t pstride, int64_t og1, int64_t ob1,
                           int64_t og2, int64_t oW1, int64_t oW2, const float4 *__restrict__ bnf1,
                           const float4 *__restrict__ bnf2, float4 *__restrict__ bnb1, float *__restrict__ grads) {
    const int m = blockIdx.y, G = 64, F1 = 8;
    const int n2 = G * C, n1 = F1 * K1;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n2 + n1 + F1; i += gridDim.x * blockDim.x) {
        if (i < n2) {
            const int gch = i / C, c = i - gch * C;
            const int fg = gch >> 5, j = gch & 31;
            float s = 0.f;
            for (int sp = 0; sp < S; ++sp) {
                const float *p = partw2 + (((int64_t)(m * S + sp) * 2 + fg) * 4) * TCX_CMAX * 32 + c * 32 + j;
#pragma unroll
                for (int qq = 0; qq < 4; ++qq) s += p[qq * TCX_CMAX * 32];
            }
            const float k2 = params[(int64_t)m * pstride + og2 + gch] * bnf2[(int64_t)m * G + gch].y;
            grads[(int64_t)m * pstride + oW2 + i] = s * k2;
        } else if (i < n2 + n1) {
            const int e = i - n2;
            const float *p = partw1 + (int64_t)m * S * n1 + e;
            float s = 0.f;
            for (int sp = 0; sp < S; ++sp) s += p[(int64_t)sp * n1];
            grads[(int64_t)m * pstride + oW1 + e] = s;
        } else {
            const int f = i - n2 - n1;
            double s1 = 0.0, s2 = 0.0;
            for (int sp = 0; sp < S; ++sp) {
                s1 += (double)partbn[(((int64_t)m * S + sp) * F1 + f) * 2];
                s2 += (double)partbn[(((int64_t)m * S + sp) * F1 + f) * 2 + 1];
            }
            grads[(int64_t)m * pstride + og1 + f] = (float)s2;
            grads[(int64_t)m * pstride + ob1 + f] = (float)s1;
            bnb1[(int64_t)m * F1 + f] = make_float4(params[(int64_t)m * pstride + og1 + f] * bnf1[(int64_t)m * F1 + f].y, 0.f, 0.f, 0.f);
        }
    }
}

}  // namespace

// shape / environment part of the decision (the workspace layout may depend on this, never on the BN mode)
bool tconv_bwd_fused_shape_ok(const NetDims &d) {
    if (!tconv_bwd_dw_use_tc(d)) return false;
    if (!tc_path_enabled("EAV_FUSE_BWD")) return false;
    return d.F1 == 8 && d.D == 8 && d.C <= TCX_CMAX;
}
bool tconv_bwd_fused_ok(const NetDims &d) { return !d.bn_train && tconv_bwd_fused_shape_ok(d); }   // any dropout mode: dz2 comes from pool1_bwd
size_t tconv_bwd_fused_partw2_floats(const NetDims &d) {
    return (size_t)d.M * tconv_bwd_dw_tc_splits(d) * 2 * 4 * TCX_CMAX * 32;
}

int launch_tconv_bwd_fused_tc(const NetDims &d, const float *x, const int32_t *x_index, const float *dz2,
                              const float *y1, const float4 *bnf1, const float4 *bnf2, const float *params,
                              float *part, float *partw2, float *partbn, float *grads, int S, cudaStream_t st) {
    const size_t smem = TCX_SMEM_FLOATS * 4;
    static PerDevice<bool> attr_set_pd(false);
    bool &attr_set = attr_set_pd.here();
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(tconv_bwd_fused_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        EAV_REQUIRE(e == cudaSuccess, (int)e, "tconv_bwd_fused_tc: %s", cudaGetErrorString(e));
        attr_set = true;
    }
    const int sms = device_sm_count();
    const int units = 2 * d.M * S;
    const int grid = units < sms ? units : sms;
    FusedBwdArgs a{dz2, y1, bnf1, bnf2, params, d.pstride, d.oW2, d.og1, d.og2, part, partw2, partbn,
                   d.variant == EAV_VARIANT_TOR ? 1 : 0};
    tconv_bwd_fused_tc_kernel<<<grid, TCX_THREADS, smem, st>>>(x, x_index, a, d.M, d.B, d.C, d.T, d.K1, d.pad1l, S);
    EAV_CUDA_LAUNCH_CHECK("tconv_bwd_fused_tc");
    return 0;
}

int launch_block1_bwd_finalize(const NetDims &d, const float *partw2, const float *partw1, const float *partbn, int S,
                               const float *params, const float4 *bnf1, const float4 *bnf2, float4 *bnb1, float *grads,
                               cudaStream_t st) {
    const int n = 64 * d.C + d.F1 * d.K1 + d.F1;
    block1_bwd_finalize_kernel<<<dim3(cdiv(n, 256), d.M), 256, 0, st>>>(partw2, partw1, partbn, S, d.C, d.K1, params, d.pstride,
                                                                        d.og1, d.ob1, d.og2, d.oW1, d.oW2, bnf1, bnf2, bnb1, grads);
    EAV_CUDA_LAUNCH_CHECK("block1_bwd_finalize");
    return 0;
}

bool tconv_bwd_dw_use_tc(const NetDims &d) {
    if (!tc_env_enabled()) return false;
    const int delta = ((d.pad1l + 3) & ~3) - d.pad1l;
    return d.F1 == 8 && (d.T & 3) == 0 && d.T >= 4 && d.T <= 512 && d.K1 + 31 + delta <= 384 && d.B * d.C >= 8;
}

// row ranges per model: minimises  waves * (rows per unit * MMA time + epilogue)  over the split count
int tconv_bwd_dw_tc_splits(const NetDims &d) {
    const int rows = d.B * d.C;
    int best = 1;
    double best_cost = 1e30;
    for (int S = 1; S <= 128 && S * 8 <= rows; ++S) {
        const int64_t units = 2ll * d.M * S;
        const double waves = (double)((units + 147) / 148);
        const double cost = waves * ((double)((rows + S - 1) / S) * 1152.0 + 12000.0);
        if (cost < best_cost) { best_cost = cost; best = S; }
    }
    return best;
}

int launch_tconv_bwd_dw_tc(const NetDims &d, const float *x, const int32_t *x_index, const float *dz1,
                           const float *y1, const float4 *bnf1, const float4 *bnb1, float *part, int S,
                           cudaStream_t st) {
    const size_t smem = ((size_t)TCW_STAGES * TCW_STAGE_FLOATS + 384 * TCW_SP) * 4;
    static PerDevice<bool> attr_set_pd(false);
    bool &attr_set = attr_set_pd.here();
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(tconv_bwd_dw_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        EAV_REQUIRE(e == cudaSuccess, (int)e, "tconv_bwd_dw_tc: %s", cudaGetErrorString(e));
        attr_set = true;
    }
    const int sms = device_sm_count();
    const int units = 2 * d.M * S;
    const int grid = units < sms ? units : sms;
    tconv_bwd_dw_tc_kernel<<<grid, TCW_THREADS, smem, st>>>(x, x_index, dz1, y1, bnf1, bnb1, d.bn_train, d.M, d.B, d.C,
                                                            d.T, d.K1, d.pad1l, S, part);
    EAV_CUDA_LAUNCH_CHECK("tconv_bwd_dw_tc");
    return 0;
}

}  // namespace eav

extern "C" int eav_tc_probe(const float *image_dev, int image_floats, int M, int N, int ksteps, int reps, int n_acc,
                            int a_off, int a_lbo, int a_sbo, int a_major, int a_step, int b_off, int b_lbo,
                            int b_sbo, int b_major, int b_step, int a_bits, int b_bits, float *d_out_dev, long long *cycles_dev,
                            void *stream) {
    using namespace eav;
    EAV_REQUIRE(image_dev && d_out_dev && cycles_dev, EAV_ERR_BAD_ARG, "eav_tc_probe: null pointer");
    EAV_REQUIRE((M == 64 || M == 128) && N >= 32 && N <= 256 && N % 32 == 0 && ksteps >= 1 && reps >= 1 && n_acc >= 1 && n_acc * N <= 512,
                EAV_ERR_BAD_ARG, "eav_tc_probe: M in {64,128}, N a multiple of 32 in [32,256]");
    EAV_REQUIRE(image_floats >= 0 && (size_t)image_floats * 4 <= 200 * 1024, EAV_ERR_BAD_ARG,
                "eav_tc_probe: image larger than 200 KB");
    ProbeArgs a{image_floats, M, N, ksteps, reps, n_acc, a_off, a_lbo, a_sbo, a_major, a_step,
                b_off, b_lbo, b_sbo, b_major, b_step, a_bits, b_bits};
    size_t smem = (size_t)image_floats * 4 + 1024;
    cudaError_t e = cudaFuncSetAttribute(tc_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    EAV_REQUIRE(e == cudaSuccess, (int)e, "eav_tc_probe: %s", cudaGetErrorString(e));
    // EAV_TC_PROBE_GRID > 1 replicates the CTA (same inputs, same outputs) to time co-resident CTAs sharing a tensor core
    const char *g = getenv("EAV_TC_PROBE_GRID");
    const int grid = g ? atoi(g) : 1;
    tc_probe_kernel<<<grid > 0 ? grid : 1, 128, smem, (cudaStream_t)stream>>>(image_dev, a, d_out_dev, cycles_dev);
    EAV_CUDA_LAUNCH_CHECK("tc_probe");
    return 0;
}

// EEG preprocessing for sm_100a: Dataload_eeg.py:85-152 (downsampling -> bandpass_filter
// -> segment_and_select_classes) for a batch of subjects.
//
// Per (subject, channel) the reference filters ONE continuous sequence made of all trials
// (order='F' reshapes at Dataload_eeg.py:94,110), so the independent units are
// S*C sequences of n_trials*trial_len samples.  Pipeline (HBM layout in brackets):
//   K_fir        raw [S][trial][ch][time] f32  ->  dec [S][ch][n_dec] f32
//                zero-phase decimating FIR, out[j] = sum_t h[t] * x[down*j + H - t].
//   K_sos<false> per (sequence, trial chunk): 5-biquad DF2T cascade (fp64 state) from ZERO
//                state over the chunk -> 10-double end state z
//   K_carry      per sequence: s_{j+1} = A^L s_j + z_j  (A^L precomputed on the host, fp64)
//   K_sos<true>  per KEPT trial only: cascade from the carried state -> epochs
//                [S][epoch][ch][ep_len] f32 (the model's input layout), epoch = n_sub*slot+q.
// The time-parallel scan is exact up to fp64 rounding (linear superposition of the
// zero-state and zero-input responses); dropped trials are never filtered in pass 3.
// Algorithmic bytes: raw read once + kept epochs written once (264 MB per subject).
#include <string.h>
#include <vector>

#include "eav_common.cuh"

namespace eav {

constexpr int FIR_R = 8;          // outputs per thread
constexpr int FIR_THREADS = 256;
constexpr int FIR_OT = FIR_R * FIR_THREADS;   // 2048 outputs per CTA tile
constexpr int MAX_TAPS = 128;

// hp[t] = (h[t], h[t + down]): the tap pair two neighbouring outputs of a thread apply to the same input sample (packed FFMA2)
struct FirTaps { float h[MAX_TAPS]; float2 hp[MAX_TAPS]; double sum; };

// Tile loader: xs[s] = x_seq[i0 + s] for s in [0, n_in), zero outside the record.  The
// continuous index is mapped to (trial, offset) ONCE per thread and then advanced
// incrementally (no per-element 64-bit division).  When every quantity is even the tile is
// fetched as 8-byte pairs (a pair never straddles a trial boundary): LDG.64 -> STS.64.
template <typename TIn>
__device__ __forceinline__ void fir_load_tile(const TIn *__restrict__ raw, int64_t seq_base, int64_t ch_off,
                                              int trial_len, int64_t trial_stride, int64_t n_seq, int64_t i0,
                                              int n_in, float *xs) {
    const bool pairs = sizeof(TIn) == 4 && ((trial_len | n_in) & 1) == 0 && (i0 & 1) == 0 &&
                       (((seq_base + ch_off) | trial_stride) & 1) == 0 && (reinterpret_cast<uintptr_t>(raw) & 7) == 0;
    const TIn *base = raw + seq_base + ch_off;
    if (pairs) {
        const int step = 2 * blockDim.x;
        int64_t i = i0 + 2 * threadIdx.x;
        // floor division that also works for the (small) negative head of the first tile
        int64_t tr = (i >= 0) ? i / trial_len : -1;
        int off = (int)(i - tr * trial_len);
        for (int s = 2 * threadIdx.x; s < n_in; s += step) {
            float2 v = make_float2(0.f, 0.f);
            if (i >= 0 && i < n_seq) v = *reinterpret_cast<const float2 *>(reinterpret_cast<const float *>(base) + tr * trial_stride + off);
            *reinterpret_cast<float2 *>(xs + s) = v;
            i += step;
            off += step;
            while (off >= trial_len) { off -= trial_len; ++tr; }
        }
    } else {
        const int step = blockDim.x;
        int64_t i = i0 + threadIdx.x;
        int64_t tr = (i >= 0) ? i / trial_len : -((-i + trial_len - 1) / trial_len);
        int off = (int)(i - tr * trial_len);
        for (int s = threadIdx.x; s < n_in; s += step) {
            float v = 0.f;
            if (i >= 0 && i < n_seq) v = (float)base[tr * trial_stride + off];
            xs[s] = v;
            i += step;
            off += step;
            while (off >= trial_len) { off -= trial_len; ++tr; }
        }
    }
}

// ---- TMA (cp.async.bulk) + mbarrier helpers -------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// 1-D bulk copy global -> shared through the TMA engine; completion is signalled on `bar`.
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// Compile-time expansion of the FIR inner product (template recursion instead of
// `#pragma unroll`, whose size heuristics silently refused the 1120-iteration nest):
// II = offset of one staged input sample relative to x[DOWN*j - H]; for each of the FIR_R
// outputs of the thread the tap index t = DOWN*r + 2H - II is a constant, so every FFMA
// takes its tap from a uniform register.
template <int DOWN, int NTAPS, bool SKIP, int T>
__device__ __forceinline__ constexpr bool fir_tap_live() {
    constexpr int H = (NTAPS - 1) / 2;
    return T >= 0 && T < NTAPS && !(SKIP && T != H && ((T - H) % DOWN) == 0);
}
// Outputs r and r + 1 of a thread see the same input sample through taps t and t + DOWN: one packed FFMA2 (Blackwell
// fma.rn.f32x2; the tap pair comes from the constant bank through a uniform register pair, the input is a broadcast
// scalar operand) instead of two FFMA.  Same products, same accumulation order per output as the scalar form.
template <int DOWN, int NTAPS, bool SKIP, int II, int R>
__device__ __forceinline__ void fir_pair(const FirTaps &taps, float v, float2 (&acc)[FIR_R / 2]) {
    constexpr int H = (NTAPS - 1) / 2;
    constexpr int t = DOWN * R + 2 * H - II, t2 = t + DOWN;
    constexpr bool l0 = fir_tap_live<DOWN, NTAPS, SKIP, t>(), l1 = fir_tap_live<DOWN, NTAPS, SKIP, t2>();
    if constexpr (l0 && l1) acc[R / 2] = __ffma2_rn(taps.hp[t], make_float2(v, v), acc[R / 2]);
    else if constexpr (l0) acc[R / 2].x = fmaf(taps.h[t], v, acc[R / 2].x);
    else if constexpr (l1) acc[R / 2].y = fmaf(taps.h[t2], v, acc[R / 2].y);
}
template <int DOWN, int NTAPS, bool SKIP, int II>
__device__ __forceinline__ void fir_one_input(const FirTaps &taps, float v, float2 (&acc0)[FIR_R / 2], float2 (&acc1)[FIR_R / 2]) {
    static_assert(FIR_R == 8, "fir_one_input is written for 8 outputs per thread");
    if constexpr (II >= 0) {
        if constexpr (II & 1) {
            fir_pair<DOWN, NTAPS, SKIP, II, 0>(taps, v, acc1); fir_pair<DOWN, NTAPS, SKIP, II, 2>(taps, v, acc1);
            fir_pair<DOWN, NTAPS, SKIP, II, 4>(taps, v, acc1); fir_pair<DOWN, NTAPS, SKIP, II, 6>(taps, v, acc1);
        } else {
            fir_pair<DOWN, NTAPS, SKIP, II, 0>(taps, v, acc0); fir_pair<DOWN, NTAPS, SKIP, II, 2>(taps, v, acc0);
            fir_pair<DOWN, NTAPS, SKIP, II, 4>(taps, v, acc0); fir_pair<DOWN, NTAPS, SKIP, II, 6>(taps, v, acc0);
        }
    }
}
template <int DOWN, int NTAPS, bool SKIP, int LEAD, int Q, int NQ>
struct FirSteps {
    static __device__ __forceinline__ void run(const FirTaps &taps, const float *xw, float m, float2 (&acc0)[FIR_R / 2],
                                               float2 (&acc1)[FIR_R / 2]) {
        const float4 v4 = *reinterpret_cast<const float4 *>(xw + 4 * Q);
        const float2 nm = make_float2(-m, -m);
        const float2 a = __fadd2_rn(make_float2(v4.x, v4.y), nm), b = __fadd2_rn(make_float2(v4.z, v4.w), nm);
        fir_one_input<DOWN, NTAPS, SKIP, 4 * Q + 0 - LEAD>(taps, a.x, acc0, acc1);
        fir_one_input<DOWN, NTAPS, SKIP, 4 * Q + 1 - LEAD>(taps, a.y, acc0, acc1);
        fir_one_input<DOWN, NTAPS, SKIP, 4 * Q + 2 - LEAD>(taps, b.x, acc0, acc1);
        fir_one_input<DOWN, NTAPS, SKIP, 4 * Q + 3 - LEAD>(taps, b.y, acc0, acc1);
        FirSteps<DOWN, NTAPS, SKIP, LEAD, Q + 1, NQ>::run(taps, xw, m, acc0, acc1);
    }
};
template <int DOWN, int NTAPS, bool SKIP, int LEAD, int NQ>
struct FirSteps<DOWN, NTAPS, SKIP, LEAD, NQ, NQ> {
    static __device__ __forceinline__ void run(const FirTaps &, const float *, float, float2 (&)[FIR_R / 2], float2 (&)[FIR_R / 2]) {}
};

// Fully unrolled decimating FIR: DOWN and NTAPS are compile-time so every tap index is
// static and the taps sit in uniform registers fed from the kernel-parameter constant bank.
// Staging: xs[s] = x[DOWN*j0 - H - 2 + s].  The 2-sample lead makes both the trial body
// (a multiple of 16 B into the tile) and every thread window (DOWN*FIR_R*4 B apart) 16-byte
// aligned, so in the trial-aligned case one elected thread stages the whole tile with three
// TMA bulk copies (tail of the previous trial, the trial, head of the next trial) on an
// mbarrier, and the compute threads read it back with LDS.128.
// Each thread removes a local offset m (its first window sample) before the fp32
// accumulation and adds m*sum(h) back in fp64: the low-frequency content (drift / DC)
// that dominates raw EEG magnitude then costs no fp32 mantissa (error ~2e-6 of channel
// RMS after the band-pass instead of ~1e-5).
// SKIP: taps at distance k*DOWN from the centre are the sinc zeros of the decimation
// filter (|h| ~ 7.6e-18); when the host verified that they are < 1e-15 |h_centre| they are
// compiled out (20 % fewer FFMA), a perturbation far below one fp64 ulp of the result.
template <typename TIn, int DOWN, int NTAPS, bool SKIP>
__global__ void __launch_bounds__(FIR_THREADS)
fir_decimate_kernel(const TIn *__restrict__ raw, const __grid_constant__ FirTaps taps, int n_trials, int n_chans,
                    int trial_len, int64_t n_dec, int tile_out, int use_tma, float *__restrict__ dec,
                    const uint8_t *__restrict__ needed) {
    extern __shared__ __align__(128) float xs[];
    __shared__ __align__(8) uint64_t bar;
    // tile == trial (TMA path): trials whose decimated samples nobody reads -- neither kept by
    // segment_and_select_classes nor the warm-up predecessor of a kept trial -- are not computed at all
    if (needed != nullptr && !needed[(int64_t)blockIdx.z * n_trials + blockIdx.x]) return;
    constexpr int H = (NTAPS - 1) / 2;
    constexpr int LEAD = 2;
    constexpr int WIN = DOWN * (FIR_R - 1) + NTAPS + LEAD;     // window per thread (138)
    constexpr int WIN4 = (WIN + 3) / 4 * 4;
    const int s = blockIdx.z, c = blockIdx.y;
    const int64_t j0 = (int64_t)blockIdx.x * tile_out;
    const int64_t n_seq = (int64_t)n_trials * trial_len;
    const int n_in = (DOWN * (tile_out - FIR_R) + WIN4 + 3) & ~3;   // smem extent every active window stays inside
    const int64_t i0 = (int64_t)DOWN * j0 - H - LEAD;
    const int64_t seq_base = (int64_t)s * n_trials * n_chans * trial_len;
    const int64_t ch_off = (int64_t)c * trial_len, tstride = (int64_t)n_chans * trial_len;
    if (use_tma) {
        // tile == trial blockIdx.x: head = last (H+LEAD) samples of the previous trial, body = the
        // trial, tail = first samples of the next trial.  Out-of-record parts are zero-filled.
        const int head = H + LEAD, tail = n_in - head - trial_len;
        const int64_t tr = blockIdx.x;
        const float *body = reinterpret_cast<const float *>(raw) + seq_base + tr * tstride + ch_off;
        if (threadIdx.x == 0) {
            mbar_init(&bar, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        if (tr == 0) for (int i = threadIdx.x; i < head; i += blockDim.x) xs[i] = 0.f;
        if (tr == n_trials - 1) for (int i = threadIdx.x; i < tail; i += blockDim.x) xs[head + trial_len + i] = 0.f;
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t bytes = (uint32_t)trial_len * 4u;
            if (tr > 0) bytes += (uint32_t)head * 4u;
            if (tr < n_trials - 1) bytes += (uint32_t)tail * 4u;
            mbar_expect_tx(&bar, bytes);
            if (tr > 0) tma_load_1d(xs, body - tstride + (trial_len - head), (uint32_t)head * 4u, &bar);
            tma_load_1d(xs + head, body, (uint32_t)trial_len * 4u, &bar);
            if (tr < n_trials - 1) tma_load_1d(xs + head + trial_len, body + tstride, (uint32_t)tail * 4u, &bar);
            mbar_wait(&bar, 0);     // only the issuing thread polls the mbarrier ...
        }
        __syncthreads();            // ... everyone else sleeps on the hardware barrier (no issue slots burnt)
    } else {
        fir_load_tile<TIn>(raw, seq_base, ch_off, trial_len, tstride, n_seq, i0, n_in, xs);
        __syncthreads();
    }

    const int jl = threadIdx.x * FIR_R;
    if (jl >= tile_out) return;
    const float *xw = xs + DOWN * jl;                          // 16B aligned: DOWN*FIR_R % 4 == 0
    const float m = xw[LEAD];
    float2 acc0[FIR_R / 2], acc1[FIR_R / 2];          // even / odd input samples; element r/2 = outputs (r, r + 1)
#pragma unroll
    for (int r = 0; r < FIR_R / 2; ++r) { acc0[r] = make_float2(0.f, 0.f); acc1[r] = make_float2(0.f, 0.f); }
    FirSteps<DOWN, NTAPS, SKIP, LEAD, 0, WIN4 / 4>::run(taps, xw, m, acc0, acc1);
    float *dst = dec + ((int64_t)s * n_chans + c) * n_dec + j0 + jl;
    const double md = (double)m * taps.sum;
    float o[FIR_R];
#pragma unroll
    for (int r = 0; r < FIR_R; ++r) {
        const float a0 = (r & 1) ? acc0[r >> 1].y : acc0[r >> 1].x, a1 = (r & 1) ? acc1[r >> 1].y : acc1[r >> 1].x;
        o[r] = (float)((double)(a0 + a1) + md);
    }
    if (jl + FIR_R <= tile_out && j0 + jl + FIR_R <= n_dec && ((n_dec | tile_out) & 3) == 0) {
        reinterpret_cast<float4 *>(dst)[0] = make_float4(o[0], o[1], o[2], o[3]);
        reinterpret_cast<float4 *>(dst)[1] = make_float4(o[4], o[5], o[6], o[7]);
    } else {
#pragma unroll
        for (int r = 0; r < FIR_R; ++r)
            if (jl + r < tile_out && j0 + jl + r < n_dec) dst[r] = o[r];
    }
}

// Any (down, n_taps): one output per thread, taps from the constant bank, fp32 with the same
// offset trick.  Correct for every shape; only the 500->100 Hz case above is tuned.
template <typename TIn>
__global__ void __launch_bounds__(FIR_THREADS)
fir_decimate_generic_kernel(const TIn *__restrict__ raw, const __grid_constant__ FirTaps taps, int n_taps,
                            int down, int n_trials, int n_chans, int trial_len, int64_t n_dec,
                            float *__restrict__ dec) {
    const int s = blockIdx.z, c = blockIdx.y;
    const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (j >= n_dec) return;
    const int H = (n_taps - 1) / 2;
    const int64_t n_seq = (int64_t)n_trials * trial_len;
    const TIn *base = raw + (int64_t)s * n_trials * n_chans * trial_len + (int64_t)c * trial_len;
    const int64_t tstride = (int64_t)n_chans * trial_len;
    auto at = [&](int64_t i) -> float {
        if (i < 0 || i >= n_seq) return 0.f;
        int64_t tr = i / trial_len;
        return (float)base[tr * tstride + (i - tr * trial_len)];
    };
    const float m = at((int64_t)down * j);
    float acc = 0.f;
    double hs = 0.0;
    for (int t = 0; t < n_taps; ++t) {
        int64_t i = (int64_t)down * j + H - t;
        if (i >= 0 && i < n_seq) { acc = fmaf(taps.h[t], at(i) - m, acc); hs += (double)taps.h[t]; }
    }
    dec[((int64_t)s * n_chans + c) * n_dec + j] = (float)((double)acc + (double)m * hs);
}

// ---------------------------------------------------------------------------------
// SOS cascade.  One thread per chunk (= one trial of the decimated sequence); the 32
// chunks of a warp are staged through smem in 32-sample tiles so global traffic stays
// 128-byte coalesced while each thread walks its own chunk serially in fp64.
// ---------------------------------------------------------------------------------
constexpr int MAX_SEC = 6;
struct SosCoef { double b0[MAX_SEC], b1[MAX_SEC], b2[MAX_SEC], a1[MAX_SEC], a2[MAX_SEC]; };
struct CarryMat { double a[4 * MAX_SEC * MAX_SEC]; };   // (2*NSEC)^2 row-major

constexpr int SOS_THREADS = 128;
constexpr int SOS_TILE = 32;

// numerator patterns of the 5-section Butterworth band-pass (b = b0 * (1, c1, c2)); PAT 0 = general coefficients
__host__ __device__ constexpr int sos_pat_c1(int pat, int k) {
    return (k == 2 || k > 4) ? 0 : ((k < 2) == (pat == 1) ? 2 : -2);
}
__host__ __device__ constexpr int sos_pat_c2(int, int k) { return k == 2 ? -1 : 1; }

template <int NSEC, bool APPLY, int PAT = 0>
__global__ void __launch_bounds__(SOS_THREADS)
sos_kernel(const float *__restrict__ dec, const __grid_constant__ SosCoef co, int n_chans, int n_trials,
           int chunk_len, int64_t n_dec, const int32_t *__restrict__ kept_trial, int n_kept,
           double *__restrict__ zstate, const double *__restrict__ start, int n_sub, int ep_len,
           int n_epochs_out, float *__restrict__ epochs, int64_t n_work, int raw_layout, int warm = 0) {
    // warm = 1 (APPLY, start == nullptr): instead of a carried start state the thread first runs the cascade from
    // rest over the PREVIOUS trial's chunk (state only, nothing written) and then filters its own chunk.  The
    // cascade forgets: the zero-input response over one chunk, ||A^L||, is checked on the host to be < 1e-6, so the
    // state after the warm-up equals the true carried state to that relative accuracy (measured 1.2e-8 of the
    // channel RMS for band [0.5, 45]; the parity gate is 1e-5).  One kernel replaces the zero-state pass over ALL
    // trials, the carry scan and the apply pass, and trials that are neither kept nor a predecessor are never read.
    // raw_layout = 1 (legacy order, band-pass BEFORE decimation): `dec` is the raw recording
    // [S][trial][ch][chunk_len], every trial is filtered (kept_trial == nullptr: identity) and the output row
    // goes to the same position of `epochs` (= the filtered recording, n_sub = 1, ep_len = chunk_len).
    __shared__ float tile[SOS_THREADS][SOS_TILE + 1];
    __shared__ int64_t row_src[SOS_THREADS];   // offset of the chunk in dec, or -1
    __shared__ int64_t row_dst[SOS_THREADS];   // offset of the trial's first epoch row in epochs
    __shared__ int64_t row_prev[SOS_THREADS];  // warm mode: offset of the previous trial's chunk, or -1
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t w = (int64_t)blockIdx.x * SOS_THREADS + tid;
    // work item -> (sequence, trial)
    int64_t my_seq = 0;
    int my_trial = -1, my_slot = 0;
    if (w < n_work) {
        if (APPLY) {
            my_seq = w / n_kept;                       // (subject, channel)
            my_slot = (int)(w - my_seq * n_kept);
            my_trial = kept_trial ? kept_trial[(my_seq / n_chans) * n_kept + my_slot] : my_slot;
        } else {
            my_seq = w / n_trials;
            my_trial = (int)(w - my_seq * n_trials);
        }
    }
    const bool my_ok = my_trial >= 0;
    if (raw_layout) {
        const int64_t subj = my_seq / n_chans, ch = my_seq - subj * n_chans;
        row_src[tid] = my_ok ? ((subj * n_trials + my_trial) * n_chans + ch) * (int64_t)chunk_len : -1;
        row_prev[tid] = -1;
        if (APPLY) row_dst[tid] = row_src[tid];
    } else {
        row_src[tid] = my_ok ? my_seq * n_dec + (int64_t)my_trial * chunk_len : -1;
        row_prev[tid] = (my_ok && my_trial > 0) ? my_seq * n_dec + (int64_t)(my_trial - 1) * chunk_len : -1;
        if (APPLY) {
            int64_t subj = my_seq / n_chans, ch = my_seq - subj * n_chans;
            row_dst[tid] = ((subj * n_epochs_out + (int64_t)my_slot * n_sub) * n_chans + ch) * ep_len;
        }
    }

    double s0[NSEC], s1[NSEC];
#pragma unroll
    for (int k = 0; k < NSEC; ++k) { s0[k] = 0.0; s1[k] = 0.0; }
    if (APPLY && my_ok && start != nullptr) {
        const double *st = start + (my_seq * n_trials + my_trial) * (2 * NSEC);
#pragma unroll
        for (int k = 0; k < NSEC; ++k) { s0[k] = st[2 * k]; s1[k] = st[2 * k + 1]; }
    }
    const int64_t ep_stride = (int64_t)n_chans * ep_len;
    __syncthreads();

    // Software pipeline: the 32 rows of the NEXT 32-sample tile are fetched into registers
    // (32 independent 128-byte row reads per warp in flight) while the current tile is being
    // filtered out of shared memory.
    float pre[32];
    for (int pass = (APPLY && warm) ? 0 : 1; pass < 2; ++pass) {
    const bool emit = APPLY && pass == 1;
    const int64_t *rows = pass == 0 ? row_prev : row_src;
    auto fetch = [&](int i0) {
        const int nt = min(SOS_TILE, chunk_len - i0);
#pragma unroll
        for (int r = 0; r < 32; ++r) {
            const int64_t src = rows[warp * 32 + r];
            pre[r] = (src >= 0 && lane < nt) ? dec[src + i0 + lane] : 0.f;
        }
    };
    fetch(0);
    for (int i0 = 0; i0 < chunk_len; i0 += SOS_TILE) {
        const int nt = min(SOS_TILE, chunk_len - i0);
        __syncwarp();
#pragma unroll
        for (int r = 0; r < 32; ++r) tile[warp * 32 + r][lane] = pre[r];
        __syncwarp();
        if (i0 + SOS_TILE < chunk_len) fetch(i0 + SOS_TILE);
        if (my_ok) {
#pragma unroll 4
            for (int i = 0; i < nt; ++i) {
                double x = (double)tile[tid][i];
#pragma unroll
                for (int k = 0; k < NSEC; ++k) {
                    if (PAT != 0) {
                        // Butterworth band-pass numerators (PAT 1: (z+1)^2, (z+1)^2, z^2-1, (z-1)^2, (z-1)^2; PAT 2: mirrored),
                        // gain in section 0, checked on the host: 20 fp64 operations per sample instead of 25 -- the kernel
                        // is bound by the fp64 pipe.  b1 * x = +-2u is one DFMA with a literal, b2 * x = +-u a negation.
                        const int c1 = sos_pat_c1(PAT, k), c2 = sos_pat_c2(PAT, k);
                        const double u = k == 0 ? co.b0[0] * x : x;
                        const double y = u + s0[k];
                        const double t = c1 == 0 ? s1[k] : fma((double)c1, u, s1[k]);
                        s0[k] = fma(-co.a1[k], y, t);
                        s1[k] = fma(-co.a2[k], y, c2 < 0 ? -u : u);
                        x = y;
                    } else {
                        double y = fma(co.b0[k], x, s0[k]);
                        s0[k] = fma(co.b1[k], x, fma(-co.a1[k], y, s1[k]));
                        s1[k] = fma(co.b2[k], x, -co.a2[k] * y);
                        x = y;
                    }
                }
                if (emit) tile[tid][i] = (float)x;
            }
        }
        if (emit) {
            __syncwarp();
            const int i = i0 + lane;
            const int q = i / ep_len, off = i - q * ep_len;
#pragma unroll 8
            for (int r = 0; r < 32; ++r) {
                const int row = warp * 32 + r;
                if (row_src[row] >= 0 && lane < nt && q < n_sub)
                    epochs[row_dst[row] + q * ep_stride + off] = tile[row][lane];
            }
        }
    }
    }
    if (!APPLY && my_ok) {
        double *z = zstate + (my_seq * n_trials + my_trial) * (2 * NSEC);
#pragma unroll
        for (int k = 0; k < NSEC; ++k) { z[2 * k] = s0[k]; z[2 * k + 1] = s1[k]; }
    }
}

// Zero-state end state of every chunk as a dot product with a precomputed table instead of a
// recurrence:  z = sum_i G[L-1-i] * x_i,  G[j] = A^j b  (b = state after a unit input from rest).
// 2*NSEC independent DFMAs per sample (10 instead of the 25 dependent ones of the cascade, and no
// serial chain); all lanes of a warp use the same G[.] at the same time, so the table tile sits in
// shared memory and is read as a broadcast.
template <int NSEC>
__global__ void __launch_bounds__(SOS_THREADS, 6)
sos_state_dot_kernel(const float *__restrict__ dec, const double *__restrict__ gtab, int n_trials, int chunk_len,
                     int64_t n_dec, double *__restrict__ zstate, int64_t n_work) {
    constexpr int NS = 2 * NSEC;
    __shared__ float tile[SOS_THREADS][SOS_TILE + 1];
    __shared__ __align__(16) double gs[SOS_TILE][NS];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t w0 = (int64_t)blockIdx.x * SOS_THREADS;
    const int64_t w = w0 + tid;
    const bool my_ok = w < n_work;
    // chunk w of the flattened (sequence, trial) list starts at w * chunk_len in dec (n_dec = n_trials*chunk_len)
    double acc[NS];
#pragma unroll
    for (int q = 0; q < NS; ++q) acc[q] = 0.0;
    float pre[32];
    constexpr int GPT = (SOS_TILE * NS + SOS_THREADS - 1) / SOS_THREADS;   // table entries per thread per tile
    double gpre[GPT];
    auto fetch = [&](int i0) {
        const int nt = min(SOS_TILE, chunk_len - i0);
#pragma unroll
        for (int r = 0; r < 32; ++r) {
            const int64_t row = w0 + warp * 32 + r;
            pre[r] = (row < n_work && lane < nt) ? dec[row * chunk_len + i0 + lane] : 0.f;
        }
#pragma unroll
        for (int g = 0; g < GPT; ++g) {
            const int e = tid + g * SOS_THREADS;
            const int i = e / NS, q = e - i * NS;
            gpre[g] = (i < nt) ? gtab[(int64_t)(chunk_len - 1 - (i0 + i)) * NS + q] : 0.0;
        }
    };
    fetch(0);
    for (int i0 = 0; i0 < chunk_len; i0 += SOS_TILE) {
        const int nt = min(SOS_TILE, chunk_len - i0);
        __syncthreads();                               // previous tile (x and G) fully consumed
#pragma unroll
        for (int r = 0; r < 32; ++r) tile[warp * 32 + r][lane] = pre[r];
#pragma unroll
        for (int g = 0; g < GPT; ++g) {
            const int e = tid + g * SOS_THREADS;
            if (e < SOS_TILE * NS) (&gs[0][0])[e] = gpre[g];
        }
        __syncthreads();
        if (i0 + SOS_TILE < chunk_len) fetch(i0 + SOS_TILE);
        if (my_ok) {
#pragma unroll 4
            for (int i = 0; i < nt; ++i) {
                const double x = (double)tile[tid][i];
#pragma unroll
                for (int q = 0; q < NS; ++q) acc[q] = fma(gs[i][q], x, acc[q]);
            }
        }
    }
    if (my_ok) {
        double *z = zstate + w * NS;
#pragma unroll
        for (int q = 0; q < NS; ++q) z[q] = acc[q];
    }
}

// start[seq][0] = 0;  start[seq][j+1] = A^L start[seq][j] + z[seq][j].  The z of the next step is
// fetched while the current 10x10 mat-vec runs, so the serial chain sees no memory latency.
template <int NSEC>
__global__ void sos_carry_kernel(const double *__restrict__ zstate, const __grid_constant__ CarryMat A,
                                 int n_trials, int64_t n_seq_total, double *__restrict__ start) {
    constexpr int NS = 2 * NSEC;
    const int64_t seq = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (seq >= n_seq_total) return;
    double s[NS], zc[NS], zn[NS];
#pragma unroll
    for (int i = 0; i < NS; ++i) { s[i] = 0.0; zc[i] = zstate[seq * n_trials * NS + i]; }
    for (int j = 0; j < n_trials; ++j) {
        double *dst = start + (seq * n_trials + j) * NS;
        if (j + 8 < n_trials) {                        // pull the z of 8 steps ahead into L1 (the chain is latency-bound)
            const char *pf = reinterpret_cast<const char *>(zstate + (seq * n_trials + j + 8) * NS);
#pragma unroll
            for (int o = 0; o < NS * 8; o += 32) asm volatile("prefetch.global.L1 [%0];" ::"l"(pf + o));
        }
        if (j + 1 < n_trials) {
            const double *z = zstate + (seq * n_trials + j + 1) * NS;
#pragma unroll
            for (int i = 0; i < NS; ++i) zn[i] = z[i];
        }
        double nx[NS];
#pragma unroll
        for (int i = 0; i < NS; ++i) {
            dst[i] = s[i];
            double a0 = zc[i], a1 = 0.0;               // two partial chains halve the dependent depth
#pragma unroll
            for (int k = 0; k < NS; k += 2) {
                a0 = fma(A.a[i * NS + k], s[k], a0);
                if (k + 1 < NS) a1 = fma(A.a[i * NS + k + 1], s[k + 1], a1);
            }
            nx[i] = a0 + a1;
        }
#pragma unroll
        for (int i = 0; i < NS; ++i) { s[i] = nx[i]; zc[i] = zn[i]; }
    }
}

__global__ void invert_slots_kernel(const int32_t *__restrict__ epoch_slot, int n_trials, int n_kept, int total,
                                    int32_t *__restrict__ kept_trial) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    int subj = i / n_trials, trial = i - subj * n_trials;
    int slot = epoch_slot[i];
    if (slot >= 0 && slot < n_kept) kept_trial[(int64_t)subj * n_kept + slot] = trial;
}

// needed[s][trial] = 1 for kept trials and for the trial before a kept one (its chunk warms the filter state up)
__global__ void needed_trials_kernel(const int32_t *__restrict__ epoch_slot, int n_trials, int n_kept, int total,
                                     uint8_t *__restrict__ needed) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int trial = i % n_trials;
    const int slot = epoch_slot[i];
    const bool kept = slot >= 0 && slot < n_kept;
    bool next_kept = false;
    if (trial + 1 < n_trials) { const int sn = epoch_slot[i + 1]; next_kept = sn >= 0 && sn < n_kept; }
    needed[i] = (kept || next_kept) ? 1 : 0;
}

// helper stream for overlapping the fp64 SOS passes with the next group's FIR (one per device)
struct PreSide { cudaStream_t stream; cudaEvent_t fork[4], join; };
static PreSide *pre_side_stream() {
    static PreSide pool[64];
    static bool made[64] = {false};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    if (!made[dev]) {
        if (cudaStreamCreateWithFlags(&pool[dev].stream, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
        for (int i = 0; i < 4; ++i) cudaEventCreateWithFlags(&pool[dev].fork[i], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&pool[dev].join, cudaEventDisableTiming);
        made[dev] = true;
    }
    return &pool[dev];
}

struct PreLayout { size_t dec, z, start, kept, gtab, filt, needed, total; };
static PreLayout pre_layout(const eav_preproc_cfg *c, bool own_dec) {
    PreLayout l;
    size_t o = 0;
    const size_t n_dec = (size_t)c->n_trials * c->trial_len / c->down;
    const size_t seqs = (size_t)c->n_subjects * c->n_chans;
    auto take = [&](size_t bytes) { size_t at = o; o = align_up(o + bytes, 256); return at; };
    l.dec = take(own_dec ? seqs * n_dec * sizeof(float) : 0);
    l.z = take(seqs * c->n_trials * 2 * c->n_sections * sizeof(double));
    l.start = take(seqs * c->n_trials * 2 * c->n_sections * sizeof(double));
    l.kept = take((size_t)c->n_subjects * c->n_trials * sizeof(int32_t));
    l.gtab = take((size_t)(c->trial_len / c->down) * 2 * c->n_sections * sizeof(double));
    // legacy order: the band-passed recording at fs_orig (fp32, raw layout)
    l.filt = take(c->order == EAV_PREPROC_ORDER_BANDPASS_FIRST
                      ? (size_t)c->n_subjects * c->n_trials * c->n_chans * c->trial_len * sizeof(float) : 0);
    l.needed = take((size_t)c->n_subjects * c->n_trials);
    l.total = o;
    return l;
}

static int check_cfg(const eav_preproc_cfg *c) {
    EAV_REQUIRE(c != nullptr, EAV_ERR_BAD_ARG, "preproc cfg is NULL");
    EAV_REQUIRE(c->n_subjects > 0 && c->n_trials > 0 && c->n_chans > 0 && c->trial_len > 0 && c->down > 0 &&
                    c->n_sub > 0, EAV_ERR_BAD_ARG, "preproc: sizes must be positive");
    EAV_REQUIRE(c->n_taps > 0 && c->n_taps <= MAX_TAPS && (c->n_taps & 1), EAV_ERR_UNSUPPORTED,
                "preproc: n_taps=%d must be odd and <= %d", c->n_taps, MAX_TAPS);
    EAV_REQUIRE(c->n_sections >= 1 && c->n_sections <= MAX_SEC, EAV_ERR_UNSUPPORTED,
                "preproc: n_sections=%d unsupported (1..%d)", c->n_sections, MAX_SEC);
    EAV_REQUIRE(c->trial_len % c->down == 0, EAV_ERR_UNSUPPORTED, "preproc: trial_len %% down != 0");
    EAV_REQUIRE((c->trial_len / c->down) % c->n_sub == 0, EAV_ERR_UNSUPPORTED, "preproc: trial not divisible into n_sub epochs");
    EAV_REQUIRE(c->n_chans <= 65535 && c->n_subjects <= 65535, EAV_ERR_UNSUPPORTED, "preproc: too many channels/subjects");
    EAV_REQUIRE(c->order == EAV_PREPROC_ORDER_DECIMATE_FIRST || c->order == EAV_PREPROC_ORDER_BANDPASS_FIRST,
                EAV_ERR_BAD_ARG, "preproc: unknown order %d", c->order);
    EAV_REQUIRE(c->order == EAV_PREPROC_ORDER_DECIMATE_FIRST || !c->raw_is_f64, EAV_ERR_UNSUPPORTED,
                "preproc: the band-pass-first order takes float32 recordings");
    return 0;
}

// EAV_SOS_STATE=dot: pass 1 of the exact scan as a table dot product (needs the G table, built on the host per call)
static bool sos_state_use_dot() {
    static int use_dot = -1;
    if (use_dot < 0) { const char *e = getenv("EAV_SOS_STATE"); use_dot = (e && strcmp(e, "dot") == 0) ? 1 : 0; }
    return use_dot != 0;
}

template <int NSEC>
static int run_sos(const eav_preproc_cfg *c, const float *dec, const SosCoef &co, const CarryMat &A,
                   const double *gtab, const int32_t *kept, int n_kept, double *z, double *start,
                   int32_t n_epochs_out, float *epochs, cudaStream_t st) {
    const int chunk = c->trial_len / c->down;
    const int64_t n_dec = (int64_t)c->n_trials * chunk;
    const int64_t seqs = (int64_t)c->n_subjects * c->n_chans;
    const int64_t work1 = seqs * c->n_trials;
    // Pass 1 has two implementations.  Measured on B200 (42 subjects): cascade recurrence 0.96 ms,
    // table dot product 1.05 ms -- the dot product needs 5 broadcast LDS.128 per sample and a broadcast
    // LDS still costs 512 B of register write-back per warp (128 B/clk/SM), which outweighs the 2.5x
    // fewer DFMAs.  The recurrence stays the default; EAV_SOS_STATE=dot selects the other one.
    if (sos_state_use_dot())
        sos_state_dot_kernel<NSEC><<<(unsigned)cdiv64(work1, SOS_THREADS), SOS_THREADS, 0, st>>>(dec, gtab, c->n_trials, chunk,
                                                                                                n_dec, z, work1);
    else
        sos_kernel<NSEC, false><<<(unsigned)cdiv64(work1, SOS_THREADS), SOS_THREADS, 0, st>>>(
            dec, co, c->n_chans, c->n_trials, chunk, n_dec, nullptr, 0, z, nullptr, c->n_sub, chunk / c->n_sub, 0, nullptr, work1, 0);
    EAV_CUDA_LAUNCH_CHECK("sos_state");
    sos_carry_kernel<NSEC><<<(unsigned)cdiv64(seqs, 64), 64, 0, st>>>(z, A, c->n_trials, seqs, start);
    EAV_CUDA_LAUNCH_CHECK("sos_carry");
    const int64_t work3 = seqs * n_kept;
    if (work3 > 0) {
        sos_kernel<NSEC, true><<<(unsigned)cdiv64(work3, SOS_THREADS), SOS_THREADS, 0, st>>>(
            dec, co, c->n_chans, c->n_trials, chunk, n_dec, kept, n_kept, nullptr, start, c->n_sub, chunk / c->n_sub,
            n_epochs_out, epochs, work3, 0);
        EAV_CUDA_LAUNCH_CHECK("sos_apply");
    }
    return 0;
}

// Legacy order (CNN_tensorflow/CNN_EEG_tf.py:64-75,182-189): band-pass the RAW recording at fs_orig, continuous over
// all trials of a (subject, channel) sequence, into `filt` (same layout as raw); the FIR then decimates `filt`.
template <int NSEC>
static int run_sos_raw(const eav_preproc_cfg *c, const float *raw, const SosCoef &co, const CarryMat &A, double *z,
                       double *start, float *filt, cudaStream_t st) {
    const int chunk = c->trial_len;
    const int64_t seqs = (int64_t)c->n_subjects * c->n_chans;
    const int64_t work = seqs * c->n_trials;
    sos_kernel<NSEC, false><<<(unsigned)cdiv64(work, SOS_THREADS), SOS_THREADS, 0, st>>>(
        raw, co, c->n_chans, c->n_trials, chunk, 0, nullptr, 0, z, nullptr, 1, chunk, 0, nullptr, work, 1);
    EAV_CUDA_LAUNCH_CHECK("sos_state(raw)");
    sos_carry_kernel<NSEC><<<(unsigned)cdiv64(seqs, 64), 64, 0, st>>>(z, A, c->n_trials, seqs, start);
    EAV_CUDA_LAUNCH_CHECK("sos_carry");
    sos_kernel<NSEC, true><<<(unsigned)cdiv64(work, SOS_THREADS), SOS_THREADS, 0, st>>>(
        raw, co, c->n_chans, c->n_trials, chunk, 0, nullptr, c->n_trials, nullptr, start, 1, chunk, 0, filt, work, 1);
    EAV_CUDA_LAUNCH_CHECK("sos_apply(raw)");
    return 0;
}

// epochs[s][slot*n_sub + q][ch][off] = dec[s][ch][trial*chunk + q*ep_len + off] for the kept trials
// (segment_and_select_classes / mysplit: CNN_EEG_tf.py:84-101,191-199)
__global__ void epoch_gather_kernel(const float *__restrict__ dec, const int32_t *__restrict__ kept_trial, int n_kept,
                                    int n_chans, int64_t n_dec, int chunk, int n_sub, int n_epochs_out,
                                    float *__restrict__ epochs, int64_t total) {
    const int ep_len = chunk / n_sub;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int t = (int)(i % chunk);
        int64_t r = i / chunk;
        const int ch = (int)(r % n_chans); r /= n_chans;
        const int slot = (int)(r % n_kept);
        const int64_t subj = r / n_kept;
        const int trial = kept_trial[subj * n_kept + slot];
        if (trial < 0) continue;
        const int q = t / ep_len, off = t - q * ep_len;
        epochs[((subj * n_epochs_out + (int64_t)slot * n_sub + q) * n_chans + ch) * ep_len + off] =
            dec[(subj * n_chans + ch) * n_dec + (int64_t)trial * chunk + t];
    }
}

template <typename TIn, bool SKIP>
static int launch_fir_5_101(const eav_preproc_cfg *c, const TIn *raw, const FirTaps &taps, float *dec,
                            const uint8_t *needed, cudaStream_t st) {
    const int chunk = c->trial_len / c->down;
    const int64_t n_dec = (int64_t)c->n_trials * chunk;
    const int tile_out = chunk <= FIR_OT ? chunk : FIR_OT;
    const int tiles = (int)cdiv64(n_dec, tile_out);
    const size_t smem = (size_t)(5 * FIR_OT + 160) * sizeof(float);
    // TMA staging needs: float input, one trial per tile, 16-byte aligned rows and tile geometry
    int use_tma = 0;
    if (sizeof(TIn) == 4 && tile_out == chunk && (tile_out % FIR_R) == 0 && (c->trial_len & 3) == 0 &&
        (reinterpret_cast<uintptr_t>(raw) & 15) == 0 && c->trial_len >= 64)
        use_tma = 1;
    static PerDevice<bool> attr_set_pd(false);
    bool &attr_set = attr_set_pd.here();
    if (!attr_set) {
        cudaFuncSetAttribute(fir_decimate_kernel<TIn, 5, 101, SKIP>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
        attr_set = true;
    }
    fir_decimate_kernel<TIn, 5, 101, SKIP><<<dim3(tiles, c->n_chans, c->n_subjects), FIR_THREADS, smem, st>>>(
        raw, taps, c->n_trials, c->n_chans, c->trial_len, n_dec, tile_out, use_tma, dec, use_tma ? needed : nullptr);
    return 0;
}

template <typename TIn>
static int run_fir(const eav_preproc_cfg *c, const TIn *raw, const FirTaps &taps, bool skip_zero_taps, float *dec,
                   cudaStream_t st, const uint8_t *needed = nullptr) {
    const int chunk = c->trial_len / c->down;
    const int64_t n_dec = (int64_t)c->n_trials * chunk;
    if (c->down == 5 && c->n_taps == 101) {
        if (skip_zero_taps) launch_fir_5_101<TIn, true>(c, raw, taps, dec, needed, st);
        else launch_fir_5_101<TIn, false>(c, raw, taps, dec, needed, st);
    } else {
        fir_decimate_generic_kernel<TIn><<<dim3((unsigned)cdiv64(n_dec, FIR_THREADS), c->n_chans, c->n_subjects),
                                          FIR_THREADS, 0, st>>>(raw, taps, c->n_taps, c->down, c->n_trials,
                                                                c->n_chans, c->trial_len, n_dec, dec);
    }
    EAV_CUDA_LAUNCH_CHECK("fir_decimate");
    return 0;
}

}  // namespace eav

using namespace eav;

extern "C" size_t eav_preproc_workspace_bytes(const eav_preproc_cfg *cfg) {
    if (check_cfg(cfg)) return 0;
    return pre_layout(cfg, true).total;
}

extern "C" int eav_preproc_run(const eav_preproc_cfg *cfg, const void *raw, const double *taps, const double *sos,
                               const int32_t *epoch_slot, int32_t n_epochs_out, float *epochs, float *dec_out,
                               void *workspace, size_t workspace_bytes, void *stream) {
    int rc = check_cfg(cfg);
    if (rc) return rc;
    EAV_REQUIRE(raw && taps && sos && epoch_slot && workspace && (epochs || n_epochs_out == 0), EAV_ERR_BAD_ARG,
                "preproc_run: null pointer");
    EAV_REQUIRE(n_epochs_out >= 0 && n_epochs_out % cfg->n_sub == 0, EAV_ERR_BAD_ARG,
                "preproc_run: n_epochs_out=%d must be a multiple of n_sub", n_epochs_out);
    const PreLayout l = pre_layout(cfg, dec_out == nullptr);
    const PreLayout lfull = pre_layout(cfg, true);
    EAV_REQUIRE(workspace_bytes >= l.total, EAV_ERR_WORKSPACE, "preproc_run: workspace %zu < required %zu",
                workspace_bytes, dec_out ? l.total : lfull.total);
    cudaStream_t st = (cudaStream_t)stream;
    char *ws = reinterpret_cast<char *>(workspace);
    float *dec = dec_out ? dec_out : reinterpret_cast<float *>(ws + l.dec);
    double *z = reinterpret_cast<double *>(ws + l.z);
    double *start = reinterpret_cast<double *>(ws + l.start);
    int32_t *kept = reinterpret_cast<int32_t *>(ws + l.kept);
    const int n_kept = n_epochs_out / cfg->n_sub;
    EAV_REQUIRE(n_kept <= cfg->n_trials, EAV_ERR_BAD_ARG, "preproc_run: more kept trials than trials");

    // taps: fp32 for the FIR (scipy casts h to x.dtype for float32 input), their exact sum in fp64
    FirTaps ft;
    memset(&ft, 0, sizeof(ft));
    double hs = 0.0;
    for (int t = 0; t < cfg->n_taps; ++t) { ft.h[t] = (float)taps[t]; hs += (double)ft.h[t]; }
    for (int t = 0; t + cfg->down < cfg->n_taps; ++t) ft.hp[t] = make_float2(ft.h[t], ft.h[t + cfg->down]);
    ft.sum = hs;

    // SOS coefficients and the zero-input transition matrix A^L of the cascade (host, fp64)
    const int NSEC = cfg->n_sections, NS = 2 * NSEC;
    SosCoef co;
    memset(&co, 0, sizeof(co));
    for (int k = 0; k < NSEC; ++k) {
        const double *r = sos + 6 * k;
        EAV_REQUIRE(r[3] == 1.0, EAV_ERR_UNSUPPORTED, "preproc_run: sos a0 must be 1 (scipy normalises it)");
        co.b0[k] = r[0]; co.b1[k] = r[1]; co.b2[k] = r[2]; co.a1[k] = r[4]; co.a2[k] = r[5];
    }
    // numerator pattern of the 5-section Butterworth band-pass (exact comparisons): b = b0 * (1, c1, c2), b0 = 1 for k > 0
    int sos_pat = 0;
    if (NSEC == 5) {
        for (int pat = 1; pat <= 2 && sos_pat == 0; ++pat) {
            bool ok = true;
            for (int k = 0; k < NSEC; ++k) {
                const double b0 = co.b0[k];
                ok = ok && co.b1[k] == (double)sos_pat_c1(pat, k) * b0 && co.b2[k] == (double)sos_pat_c2(pat, k) * b0 &&
                     (k == 0 || b0 == 1.0);
            }
            if (ok) sos_pat = pat;
        }
    }
    { const char *e = getenv("EAV_SOS_PATTERN"); if (e && e[0] == '0') sos_pat = 0; }
    const bool bp_first = cfg->order == EAV_PREPROC_ORDER_BANDPASS_FIRST;
    const int chunk = cfg->trial_len / cfg->down;               // decimated samples per trial
    const int sos_chunk = bp_first ? cfg->trial_len : chunk;    // samples per trial at the rate the SOS runs at
    CarryMat A;
    memset(&A, 0, sizeof(A));
    {
        // one-step zero-input matrix: column e = state after feeding x = 0 from unit state e
        long double Am[4 * MAX_SEC * MAX_SEC] = {0}, P[4 * MAX_SEC * MAX_SEC] = {0}, Tm[4 * MAX_SEC * MAX_SEC];
        for (int e = 0; e < NS; ++e) {
            long double s0[MAX_SEC] = {0}, s1[MAX_SEC] = {0};
            if (e & 1) s1[e >> 1] = 1.0L; else s0[e >> 1] = 1.0L;
            long double x = 0.0L;
            for (int k = 0; k < NSEC; ++k) {
                long double y = (long double)co.b0[k] * x + s0[k];
                s0[k] = (long double)co.b1[k] * x - (long double)co.a1[k] * y + s1[k];
                s1[k] = (long double)co.b2[k] * x - (long double)co.a2[k] * y;
                x = y;
            }
            for (int k = 0; k < NSEC; ++k) { Am[(2 * k) * NS + e] = s0[k]; Am[(2 * k + 1) * NS + e] = s1[k]; }
        }
        for (int i = 0; i < NS; ++i) P[i * NS + i] = 1.0L;
        int pw = sos_chunk;   // P = Am^chunk by binary exponentiation
        while (pw > 0) {
            if (pw & 1) {
                for (int i = 0; i < NS; ++i) for (int j = 0; j < NS; ++j) {
                    long double a = 0; for (int k = 0; k < NS; ++k) a += P[i * NS + k] * Am[k * NS + j];
                    Tm[i * NS + j] = a; }
                memcpy(P, Tm, sizeof(long double) * NS * NS);
            }
            for (int i = 0; i < NS; ++i) for (int j = 0; j < NS; ++j) {
                long double a = 0; for (int k = 0; k < NS; ++k) a += Am[i * NS + k] * Am[k * NS + j];
                Tm[i * NS + j] = a; }
            memcpy(Am, Tm, sizeof(long double) * NS * NS);
            pw >>= 1;
        }
        for (int i = 0; i < NS * NS; ++i) A.a[i] = (double)P[i];
    }
    // G[j] = A1^j b for j in [0, chunk): b = state after feeding a unit sample into the cascade at rest,
    // A1 = the one-step zero-input matrix (recomputed here: the squaring above overwrote it).
    // Only the (non-default) dot-product form of pass 1 reads it: building it costs ~1 ms of host time per call and its
    // pageable upload synchronises the stream -- with it built unconditionally the GPU idled 0.6 ms of every 3.2 ms run.
    double *gtab = reinterpret_cast<double *>(ws + l.gtab);
    if (sos_state_use_dot() && !bp_first) {
        std::vector<long double> A1((size_t)NS * NS), g(NS), gn(NS);
        for (int e = 0; e <= NS; ++e) {            // e == NS: unit input from rest -> b
            long double s0[MAX_SEC] = {0}, s1[MAX_SEC] = {0};
            if (e < NS) { if (e & 1) s1[e >> 1] = 1.0L; else s0[e >> 1] = 1.0L; }
            long double x = (e == NS) ? 1.0L : 0.0L;
            for (int k = 0; k < NSEC; ++k) {
                long double y = (long double)co.b0[k] * x + s0[k];
                s0[k] = (long double)co.b1[k] * x - (long double)co.a1[k] * y + s1[k];
                s1[k] = (long double)co.b2[k] * x - (long double)co.a2[k] * y;
                x = y;
            }
            for (int k = 0; k < NSEC; ++k) {
                if (e < NS) { A1[(size_t)(2 * k) * NS + e] = s0[k]; A1[(size_t)(2 * k + 1) * NS + e] = s1[k]; }
                else { g[2 * k] = s0[k]; g[2 * k + 1] = s1[k]; }
            }
        }
        std::vector<double> host((size_t)chunk * NS);
        for (int j = 0; j < chunk; ++j) {
            for (int i = 0; i < NS; ++i) host[(size_t)j * NS + i] = (double)g[i];
            for (int i = 0; i < NS; ++i) {
                long double a = 0;
                for (int k = 0; k < NS; ++k) a += A1[(size_t)i * NS + k] * g[k];
                gn[i] = a;
            }
            g = gn;
        }
        // pageable source: the runtime stages it before returning, so `host` may die at scope exit
        cudaMemcpyAsync(gtab, host.data(), host.size() * sizeof(double), cudaMemcpyHostToDevice, st);
    }

    // kept_trial[s][slot] = trial (or -1)
    const int total = cfg->n_subjects * cfg->n_trials;
    cudaMemsetAsync(kept, 0xff, (size_t)cfg->n_subjects * cfg->n_trials * sizeof(int32_t), st);
    if (n_kept > 0) {
        invert_slots_kernel<<<cdiv(total, 256), 256, 0, st>>>(epoch_slot, cfg->n_trials, n_kept, total, kept);
        EAV_CUDA_LAUNCH_CHECK("invert_slots");
    }

    // sinc zeros of the decimation filter (taps k*down away from the centre): compile them out when negligible
    bool skip_zero = true;
    {
        const int Hh = (cfg->n_taps - 1) / 2;
        for (int t = 0; t < cfg->n_taps; ++t)
            if (t != Hh && ((t - Hh) % cfg->down) == 0 && fabs(taps[t]) > 1e-15 * fabs(taps[Hh])) skip_zero = false;
    }
    if (bp_first) {
        // band-pass at fs_orig over the raw recording -> filt, decimate filt -> dec, gather the kept epochs
        float *filt = reinterpret_cast<float *>(ws + l.filt);
        const float *rawf = reinterpret_cast<const float *>(raw);
        switch (NSEC) {
            case 1: rc = run_sos_raw<1>(cfg, rawf, co, A, z, start, filt, st); break;
            case 2: rc = run_sos_raw<2>(cfg, rawf, co, A, z, start, filt, st); break;
            case 3: rc = run_sos_raw<3>(cfg, rawf, co, A, z, start, filt, st); break;
            case 4: rc = run_sos_raw<4>(cfg, rawf, co, A, z, start, filt, st); break;
            case 5: rc = run_sos_raw<5>(cfg, rawf, co, A, z, start, filt, st); break;
            default: rc = run_sos_raw<6>(cfg, rawf, co, A, z, start, filt, st); break;
        }
        if (rc) return rc;
        rc = run_fir<float>(cfg, filt, ft, skip_zero, dec, st);
        if (rc) return rc;
        if (n_kept > 0) {
            const int64_t n_dec = (int64_t)cfg->n_trials * chunk;
            const int64_t tot = (int64_t)cfg->n_subjects * n_kept * cfg->n_chans * chunk;
            epoch_gather_kernel<<<(unsigned)std::min<int64_t>(cdiv64(tot, 256), 148 * 32), 256, 0, st>>>(
                dec, kept, n_kept, cfg->n_chans, n_dec, chunk, cfg->n_sub, n_epochs_out, epochs, tot);
            EAV_CUDA_LAUNCH_CHECK("epoch_gather");
        }
        return 0;
    }
    // Forgetting-filter fast path (the default whenever it is valid): if the cascade's zero-input response over one
    // chunk is negligible (||A^L||_inf < 1e-6: 2.9e-7 for the band [0.5, 45] of the callers, 1e-15 for [1, 40], but
    // 1e-3 for [0.3, 49], which therefore takes the exact path below), a trial's start state is the end state of a
    // from-rest run over the previous trial alone.  Then only kept trials and their predecessors are decimated, and ONE
    // SOS kernel (warm-up over the previous chunk + apply) replaces state pass, carry scan and apply pass.
    // EAV_SOS_EXACT=1 forces the exact chunked scan.
    {
        double ninf = 0.0;
        for (int i = 0; i < NS; ++i) {
            double rsum = 0.0;
            for (int j = 0; j < NS; ++j) rsum += fabs(A.a[i * NS + j]);
            if (rsum > ninf) ninf = rsum;
        }
        const char *ex = getenv("EAV_SOS_EXACT");
        const bool exact = ex && ex[0] == '1';
        if (!exact && ninf < 1e-6 && dec_out == nullptr && n_kept > 0) {
            uint8_t *needed = reinterpret_cast<uint8_t *>(ws + l.needed);
            needed_trials_kernel<<<cdiv(total, 256), 256, 0, st>>>(epoch_slot, cfg->n_trials, n_kept, total, needed);
            EAV_CUDA_LAUNCH_CHECK("needed_trials");
            // Subject groups: the fp64-bound SOS kernel of group i runs on a forked stream while the HBM-bound FIR of
            // group i+1 runs on the caller's stream (EAV_PREPROC_GROUPS, default 1).
            int ng = 1;
            { const char *e = getenv("EAV_PREPROC_GROUPS"); if (e && atoi(e) > 0) ng = atoi(e); }
            if (ng > cfg->n_subjects) ng = cfg->n_subjects;
            PreSide *sd = ng > 1 ? pre_side_stream() : nullptr;
            if (sd == nullptr) ng = 1;
            const size_t raw_subj_e = (size_t)cfg->n_trials * cfg->n_chans * cfg->trial_len;
            const size_t dec_subj_e = (size_t)cfg->n_chans * cfg->n_trials * chunk;
            const size_t ep_subj_e = (size_t)n_epochs_out * cfg->n_chans * (chunk / cfg->n_sub);
            const int64_t n_dec = (int64_t)cfg->n_trials * chunk;
            for (int g = 0; g < ng; ++g) {
                const int s0 = (int)((int64_t)cfg->n_subjects * g / ng), s1 = (int)((int64_t)cfg->n_subjects * (g + 1) / ng);
                if (s1 <= s0) continue;
                eav_preproc_cfg sub = *cfg;
                sub.n_subjects = s1 - s0;
                float *dec_g = dec + s0 * dec_subj_e;
                const uint8_t *need_g = needed + (size_t)s0 * cfg->n_trials;
                if (cfg->raw_is_f64) rc = run_fir<double>(&sub, reinterpret_cast<const double *>(raw) + s0 * raw_subj_e, ft, skip_zero, dec_g, st, need_g);
                else rc = run_fir<float>(&sub, reinterpret_cast<const float *>(raw) + s0 * raw_subj_e, ft, skip_zero, dec_g, st, need_g);
                if (rc) return rc;
                cudaStream_t sst = st;
                if (sd != nullptr) {
                    cudaEventRecord(sd->fork[g % 4], st);
                    cudaStreamWaitEvent(sd->stream, sd->fork[g % 4], 0);
                    sst = sd->stream;
                }
                const int64_t work = (int64_t)sub.n_subjects * cfg->n_chans * n_kept;
                const int32_t *kept_g = kept + (size_t)s0 * n_kept;
                float *ep_g = epochs + s0 * ep_subj_e;
#define EAV_WARM_P(NS_, PAT_)                                                                                       \
    sos_kernel<NS_, true, PAT_><<<(unsigned)cdiv64(work, SOS_THREADS), SOS_THREADS, 0, sst>>>(                      \
        dec_g, co, cfg->n_chans, cfg->n_trials, chunk, n_dec, kept_g, n_kept, nullptr, nullptr, cfg->n_sub,         \
        chunk / cfg->n_sub, n_epochs_out, ep_g, work, 0, 1)
#define EAV_WARM(NS_) EAV_WARM_P(NS_, 0)
                switch (NSEC) {
                    case 1: EAV_WARM(1); break;
                    case 2: EAV_WARM(2); break;
                    case 3: EAV_WARM(3); break;
                    case 4: EAV_WARM(4); break;
                    case 5:
                        if (sos_pat == 1) EAV_WARM_P(5, 1);
                        else if (sos_pat == 2) EAV_WARM_P(5, 2);
                        else EAV_WARM(5);
                        break;
                    default: EAV_WARM(6); break;
                }
#undef EAV_WARM
#undef EAV_WARM_P
                EAV_CUDA_LAUNCH_CHECK("sos_warm_apply");
            }
            if (sd != nullptr) {
                cudaEventRecord(sd->join, sd->stream);
                cudaStreamWaitEvent(st, sd->join, 0);
            }
            return 0;
        }
    }
    // Optional grouping (EAV_PREPROC_GROUPS=n): the FIR of group i+1 on the caller's stream, the SOS
    // passes of group i on a forked stream.  The idea was to overlap the fp32/HBM-bound FIR with the
    // fp64-bound SOS passes; measured on B200 it does NOT pay (42 subjects: 1 group 4.42 ms, 2: 4.78,
    // 3: 5.15, 6: 6.76 -- the FIR grid saturates every SM's CTA slots, so the forked kernels only
    // run in its tail), hence one group by default.
    const int S = cfg->n_subjects;
    int n_groups = 1;
    {
        static int forced = -1;
        if (forced < 0) { const char *e = getenv("EAV_PREPROC_GROUPS"); forced = e ? atoi(e) : 0; }
        if (forced > 0) n_groups = forced < S ? forced : S;
    }
    PreSide *side = n_groups > 1 ? pre_side_stream() : nullptr;
    if (side == nullptr) n_groups = 1;
    const size_t raw_subj = (size_t)cfg->n_trials * cfg->n_chans * cfg->trial_len;          // elements
    const size_t dec_subj = (size_t)cfg->n_chans * cfg->n_trials * chunk;
    const size_t st_subj = (size_t)cfg->n_chans * cfg->n_trials * NS;
    const size_t ep_subj = (size_t)n_epochs_out * cfg->n_chans * (chunk / cfg->n_sub);
    for (int g = 0; g < n_groups; ++g) {
        const int s0 = (int)((int64_t)S * g / n_groups), s1 = (int)((int64_t)S * (g + 1) / n_groups);
        if (s1 <= s0) continue;
        eav_preproc_cfg sub = *cfg;
        sub.n_subjects = s1 - s0;
        float *dec_g = dec + s0 * dec_subj;
        if (cfg->raw_is_f64) rc = run_fir<double>(&sub, reinterpret_cast<const double *>(raw) + s0 * raw_subj, ft, skip_zero, dec_g, st);
        else rc = run_fir<float>(&sub, reinterpret_cast<const float *>(raw) + s0 * raw_subj, ft, skip_zero, dec_g, st);
        if (rc) return rc;
        cudaStream_t sst = st;
        if (side != nullptr) {
            cudaEventRecord(side->fork[g % 4], st);
            cudaStreamWaitEvent(side->stream, side->fork[g % 4], 0);
            sst = side->stream;
        }
        double *z_g = z + s0 * st_subj, *start_g = start + s0 * st_subj;
        const int32_t *kept_g = kept + (size_t)s0 * n_kept;
        float *ep_g = epochs ? epochs + s0 * ep_subj : nullptr;
        switch (NSEC) {
            case 1: rc = run_sos<1>(&sub, dec_g, co, A, gtab, kept_g, n_kept, z_g, start_g, n_epochs_out, ep_g, sst); break;
            case 2: rc = run_sos<2>(&sub, dec_g, co, A, gtab, kept_g, n_kept, z_g, start_g, n_epochs_out, ep_g, sst); break;
            case 3: rc = run_sos<3>(&sub, dec_g, co, A, gtab, kept_g, n_kept, z_g, start_g, n_epochs_out, ep_g, sst); break;
            case 4: rc = run_sos<4>(&sub, dec_g, co, A, gtab, kept_g, n_kept, z_g, start_g, n_epochs_out, ep_g, sst); break;
            case 5: rc = run_sos<5>(&sub, dec_g, co, A, gtab, kept_g, n_kept, z_g, start_g, n_epochs_out, ep_g, sst); break;
            default: rc = run_sos<6>(&sub, dec_g, co, A, gtab, kept_g, n_kept, z_g, start_g, n_epochs_out, ep_g, sst); break;
        }
        if (rc) return rc;
    }
    if (side != nullptr) {
        cudaEventRecord(side->join, side->stream);
        cudaStreamWaitEvent(st, side->join, 0);
    }
    return 0;
}

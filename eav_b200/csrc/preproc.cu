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
#include "eav_common.cuh"

namespace eav {

constexpr int FIR_R = 8;          // outputs per thread
constexpr int FIR_THREADS = 256;
constexpr int FIR_OT = FIR_R * FIR_THREADS;   // 2048 outputs per CTA tile
constexpr int MAX_TAPS = 128;

struct FirTaps { float h[MAX_TAPS]; double sum; };

// Generic-shape tile loader: xs[s] = x_seq[i0 + s] for s in [0, n_in), zero outside the record.
template <typename TIn>
__device__ __forceinline__ void fir_load_tile(const TIn *__restrict__ raw, int64_t seq_base, int64_t ch_off,
                                              int trial_len, int64_t trial_stride, int64_t n_seq, int64_t i0,
                                              int n_in, float *xs) {
    for (int s = threadIdx.x; s < n_in; s += blockDim.x) {
        int64_t i = i0 + s;
        float v = 0.f;
        if (i >= 0 && i < n_seq) {
            int64_t tr = i / trial_len;
            int off = (int)(i - tr * trial_len);
            v = (float)raw[seq_base + tr * trial_stride + ch_off + off];
        }
        xs[s] = v;
    }
}

// Fully unrolled decimating FIR: DOWN and NTAPS are compile-time so every tap index is
// static and the taps are read straight from the kernel-parameter constant bank.
// Each thread removes a local offset m (its first window sample) before the fp32
// accumulation and adds m*sum(h) back in fp64: the low-frequency content (drift / DC)
// that dominates raw EEG magnitude then costs no fp32 mantissa (error ~2e-6 of channel
// RMS after the band-pass instead of ~1e-5).
template <typename TIn, int DOWN, int NTAPS>
__global__ void __launch_bounds__(FIR_THREADS)
fir_decimate_kernel(const TIn *__restrict__ raw, const __grid_constant__ FirTaps taps, int n_trials, int n_chans,
                    int trial_len, int64_t n_dec, int tile_out, float *__restrict__ dec) {
    extern __shared__ __align__(16) float xs[];
    constexpr int H = (NTAPS - 1) / 2;
    constexpr int WIN = DOWN * (FIR_R - 1) + NTAPS;            // window per thread (136)
    constexpr int WIN4 = (WIN + 3) / 4 * 4;
    const int s = blockIdx.z, c = blockIdx.y;
    const int64_t j0 = (int64_t)blockIdx.x * tile_out;
    const int64_t n_seq = (int64_t)n_trials * trial_len;
    const int n_in = DOWN * tile_out + NTAPS + 8;             // smem extent (covers every active thread's WIN4)
    fir_load_tile<TIn>(raw, (int64_t)s * n_trials * n_chans * trial_len, (int64_t)c * trial_len, trial_len,
                       (int64_t)n_chans * trial_len, n_seq, (int64_t)DOWN * j0 - H, n_in, xs);
    __syncthreads();

    const int jl = threadIdx.x * FIR_R;
    if (jl >= tile_out) return;
    const float *xw = xs + DOWN * jl;                          // 16B aligned when DOWN*FIR_R % 4 == 0
    const float m = xw[0];
    float acc0[FIR_R], acc1[FIR_R];
#pragma unroll
    for (int r = 0; r < FIR_R; ++r) { acc0[r] = 0.f; acc1[r] = 0.f; }
#pragma unroll
    for (int q = 0; q < WIN4 / 4; ++q) {
        float4 v4 = *reinterpret_cast<const float4 *>(xw + 4 * q);
        const float v[4] = {v4.x - m, v4.y - m, v4.z - m, v4.w - m};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int ii = 4 * q + e;
#pragma unroll
            for (int r = 0; r < FIR_R; ++r) {
                const int t = DOWN * r + 2 * H - ii;           // tap index for output r, input ii
                if (t >= 0 && t < NTAPS) {
                    if (ii & 1) acc1[r] = fmaf(taps.h[t], v[e], acc1[r]);
                    else acc0[r] = fmaf(taps.h[t], v[e], acc0[r]);
                }
            }
        }
    }
    float *dst = dec + ((int64_t)s * n_chans + c) * n_dec + j0 + jl;
    const double md = (double)m * taps.sum;
#pragma unroll
    for (int r = 0; r < FIR_R; ++r)
        if (jl + r < tile_out && j0 + jl + r < n_dec) dst[r] = (float)((double)(acc0[r] + acc1[r]) + md);
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

template <int NSEC, bool APPLY>
__global__ void __launch_bounds__(SOS_THREADS)
sos_kernel(const float *__restrict__ dec, const __grid_constant__ SosCoef co, int n_chans, int n_trials,
           int chunk_len, int64_t n_dec, const int32_t *__restrict__ kept_trial, int n_kept,
           double *__restrict__ zstate, const double *__restrict__ start, int n_sub, int ep_len,
           int n_epochs_out, float *__restrict__ epochs, int64_t n_work) {
    __shared__ float tile[SOS_THREADS][SOS_TILE + 1];
    __shared__ int64_t row_src[SOS_THREADS];   // offset of the chunk in dec, or -1
    __shared__ int64_t row_dst[SOS_THREADS];   // offset of the trial's first epoch row in epochs
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t w = (int64_t)blockIdx.x * SOS_THREADS + tid;
    // work item -> (sequence, trial)
    int64_t my_seq = 0;
    int my_trial = -1, my_slot = 0;
    if (w < n_work) {
        if (APPLY) {
            my_seq = w / n_kept;                       // (subject, channel)
            my_slot = (int)(w - my_seq * n_kept);
            my_trial = kept_trial[(my_seq / n_chans) * n_kept + my_slot];
        } else {
            my_seq = w / n_trials;
            my_trial = (int)(w - my_seq * n_trials);
        }
    }
    const bool my_ok = my_trial >= 0;
    row_src[tid] = my_ok ? my_seq * n_dec + (int64_t)my_trial * chunk_len : -1;
    if (APPLY) {
        int64_t subj = my_seq / n_chans, ch = my_seq - subj * n_chans;
        row_dst[tid] = ((subj * n_epochs_out + (int64_t)my_slot * n_sub) * n_chans + ch) * ep_len;
    }

    double s0[NSEC], s1[NSEC];
#pragma unroll
    for (int k = 0; k < NSEC; ++k) { s0[k] = 0.0; s1[k] = 0.0; }
    if (APPLY && my_ok) {
        const double *st = start + (my_seq * n_trials + my_trial) * (2 * NSEC);
#pragma unroll
        for (int k = 0; k < NSEC; ++k) { s0[k] = st[2 * k]; s1[k] = st[2 * k + 1]; }
    }
    const int64_t ep_stride = (int64_t)n_chans * ep_len;
    __syncthreads();

    for (int i0 = 0; i0 < chunk_len; i0 += SOS_TILE) {
        const int nt = min(SOS_TILE, chunk_len - i0);
        __syncwarp();
        // each warp stages the 32 chunks its own lanes walk: lane = sample -> 128 B per row
        for (int r = 0; r < 32; ++r) {
            const int row = warp * 32 + r;
            const int64_t src = row_src[row];
            float v = 0.f;
            if (src >= 0 && lane < nt) v = dec[src + i0 + lane];
            tile[row][lane] = v;
        }
        __syncwarp();
        if (my_ok) {
#pragma unroll 4
            for (int i = 0; i < nt; ++i) {
                double x = (double)tile[tid][i];
#pragma unroll
                for (int k = 0; k < NSEC; ++k) {
                    double y = fma(co.b0[k], x, s0[k]);
                    s0[k] = fma(co.b1[k], x, fma(-co.a1[k], y, s1[k]));
                    s1[k] = fma(co.b2[k], x, -co.a2[k] * y);
                    x = y;
                }
                if (APPLY) tile[tid][i] = (float)x;
            }
        }
        if (APPLY) {
            __syncwarp();
            const int i = i0 + lane;
            const int q = i / ep_len, off = i - q * ep_len;
            for (int r = 0; r < 32; ++r) {
                const int row = warp * 32 + r;
                if (row_src[row] >= 0 && lane < nt && q < n_sub)
                    epochs[row_dst[row] + q * ep_stride + off] = tile[row][lane];
            }
        }
    }
    if (!APPLY && my_ok) {
        double *z = zstate + (my_seq * n_trials + my_trial) * (2 * NSEC);
#pragma unroll
        for (int k = 0; k < NSEC; ++k) { z[2 * k] = s0[k]; z[2 * k + 1] = s1[k]; }
    }
}

// start[seq][0] = 0;  start[seq][j+1] = A^L start[seq][j] + z[seq][j]
template <int NSEC>
__global__ void sos_carry_kernel(const double *__restrict__ zstate, const __grid_constant__ CarryMat A,
                                 int n_trials, int64_t n_seq_total, double *__restrict__ start) {
    constexpr int NS = 2 * NSEC;
    const int64_t seq = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (seq >= n_seq_total) return;
    double s[NS];
#pragma unroll
    for (int i = 0; i < NS; ++i) s[i] = 0.0;
    for (int j = 0; j < n_trials; ++j) {
        double *dst = start + (seq * n_trials + j) * NS;
        const double *z = zstate + (seq * n_trials + j) * NS;
        double nx[NS];
#pragma unroll
        for (int i = 0; i < NS; ++i) {
            dst[i] = s[i];
            double a = z[i];
#pragma unroll
            for (int k = 0; k < NS; ++k) a = fma(A.a[i * NS + k], s[k], a);
            nx[i] = a;
        }
#pragma unroll
        for (int i = 0; i < NS; ++i) s[i] = nx[i];
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

struct PreLayout { size_t dec, z, start, kept, total; };
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
    return 0;
}

template <int NSEC>
static int run_sos(const eav_preproc_cfg *c, const float *dec, const SosCoef &co, const CarryMat &A,
                   const int32_t *kept, int n_kept, double *z, double *start, int32_t n_epochs_out, float *epochs,
                   cudaStream_t st) {
    const int chunk = c->trial_len / c->down;
    const int64_t n_dec = (int64_t)c->n_trials * chunk;
    const int64_t seqs = (int64_t)c->n_subjects * c->n_chans;
    const int64_t work1 = seqs * c->n_trials;
    sos_kernel<NSEC, false><<<(unsigned)cdiv64(work1, SOS_THREADS), SOS_THREADS, 0, st>>>(
        dec, co, c->n_chans, c->n_trials, chunk, n_dec, nullptr, 0, z, nullptr, c->n_sub, chunk / c->n_sub, 0, nullptr, work1);
    EAV_CUDA_LAUNCH_CHECK("sos_state");
    sos_carry_kernel<NSEC><<<(unsigned)cdiv64(seqs, 64), 64, 0, st>>>(z, A, c->n_trials, seqs, start);
    EAV_CUDA_LAUNCH_CHECK("sos_carry");
    const int64_t work3 = seqs * n_kept;
    if (work3 > 0) {
        sos_kernel<NSEC, true><<<(unsigned)cdiv64(work3, SOS_THREADS), SOS_THREADS, 0, st>>>(
            dec, co, c->n_chans, c->n_trials, chunk, n_dec, kept, n_kept, nullptr, start, c->n_sub, chunk / c->n_sub,
            n_epochs_out, epochs, work3);
        EAV_CUDA_LAUNCH_CHECK("sos_apply");
    }
    return 0;
}

template <typename TIn>
static int run_fir(const eav_preproc_cfg *c, const TIn *raw, const FirTaps &taps, float *dec, cudaStream_t st) {
    const int chunk = c->trial_len / c->down;
    const int64_t n_dec = (int64_t)c->n_trials * chunk;
    if (c->down == 5 && c->n_taps == 101) {
        const int tile_out = chunk <= FIR_OT ? chunk : FIR_OT;
        const int tiles = (int)cdiv64(n_dec, tile_out);
        size_t smem = (size_t)(5 * FIR_OT + 101 + 8) * sizeof(float);
        static bool attr_set = false;
        if (!attr_set) {
            cudaFuncSetAttribute(fir_decimate_kernel<float, 5, 101>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
            cudaFuncSetAttribute(fir_decimate_kernel<double, 5, 101>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
            attr_set = true;
        }
        fir_decimate_kernel<TIn, 5, 101><<<dim3(tiles, c->n_chans, c->n_subjects), FIR_THREADS, smem, st>>>(
            raw, taps, c->n_trials, c->n_chans, c->trial_len, n_dec, tile_out, dec);
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
    const int chunk = cfg->trial_len / cfg->down;
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
        int pw = chunk;   // P = Am^chunk by binary exponentiation
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

    // kept_trial[s][slot] = trial (or -1)
    const int total = cfg->n_subjects * cfg->n_trials;
    cudaMemsetAsync(kept, 0xff, (size_t)cfg->n_subjects * cfg->n_trials * sizeof(int32_t), st);
    if (n_kept > 0) {
        invert_slots_kernel<<<cdiv(total, 256), 256, 0, st>>>(epoch_slot, cfg->n_trials, n_kept, total, kept);
        EAV_CUDA_LAUNCH_CHECK("invert_slots");
    }

    if (cfg->raw_is_f64) rc = run_fir<double>(cfg, reinterpret_cast<const double *>(raw), ft, dec, st);
    else rc = run_fir<float>(cfg, reinterpret_cast<const float *>(raw), ft, dec, st);
    if (rc) return rc;

    switch (NSEC) {
        case 1: return run_sos<1>(cfg, dec, co, A, kept, n_kept, z, start, n_epochs_out, epochs, st);
        case 2: return run_sos<2>(cfg, dec, co, A, kept, n_kept, z, start, n_epochs_out, epochs, st);
        case 3: return run_sos<3>(cfg, dec, co, A, kept, n_kept, z, start, n_epochs_out, epochs, st);
        case 4: return run_sos<4>(cfg, dec, co, A, kept, n_kept, z, start, n_epochs_out, epochs, st);
        case 5: return run_sos<5>(cfg, dec, co, A, kept, n_kept, z, start, n_epochs_out, epochs, st);
        default: return run_sos<6>(cfg, dec, co, A, kept, n_kept, z, start, n_epochs_out, epochs, st);
    }
}

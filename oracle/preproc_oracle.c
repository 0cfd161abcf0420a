/*
 * oracle/preproc_oracle.c -- CPU restatement (plain C, fp64) of the EEG
 * preprocessing arithmetic on the EAV hot path.
 *
 * TEST INFRASTRUCTURE ONLY: only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library, and only as the
 * checker or the timed CPU baseline -- never on the product path.
 *
 * What it restates (reference = nubcico/EAV, third-party arithmetic = scipy 1.18.1):
 *   oracle_fir_decimate  <- Dataload_eeg.py:85-102 `downsampling`, i.e.
 *       scipy.signal.resample_poly(tm, up=1, down=q, axis=1) on the per-channel
 *       CONTINUOUS sequence (all trials concatenated in time, order='F' reshape at
 *       Dataload_eeg.py:94).  scipy's published algorithm: h = firwin(2*10*q+1,
 *       1/q, window=('kaiser',5.0)); zero-phase polyphase FIR with zero padding
 *       outside the record; n_pre_pad/n_pre_remove centre the filter so that
 *           out[j] = sum_{d=-H}^{H} h[H+d] * x[q*j - d],  x = 0 outside [0, n)
 *       with H = 10*q.
 *   oracle_sosfilt       <- Dataload_eeg.py:104-121 `bandpass_filter`, i.e.
 *       scipy.signal.sosfilt (Cython _sosfilt): cascaded biquads, transposed
 *       direct form II, zero initial state, state carried over the whole
 *       continuous sequence:
 *           y  = b0*x + s0;  s0 = b1*x - a1*y + s1;  s1 = b2*x - a2*y;  x = y
 *   oracle_epoch_gather  <- Dataload_eeg.py:123-152 `segment_and_select_classes`
 *       epoch e = 4*k + q  <-  trial k, samples [500q, 500q+500) (row P4).
 *
 * Build: gcc -O2 -shared -fPIC (no -ffast-math: the sosfilt restatement must keep
 * scipy's operation order to stay bit-exact with it).
 */
#include <stddef.h>
#include <stdint.h>

/* seq is the continuous per-channel sequence of n samples, already padded by the
 * caller with H zeros in front and H + q zeros behind (x[i] == seq[H + i]), so
 *     out[j] = sum_{t=0}^{2H} h[t] * seq[q*j + 2H - t]          (t = H + d). */
void oracle_fir_decimate(const double *seq, const double *h, int H, int q, double *out, int64_t n_out) {
    const int ntaps = 2 * H + 1;
    for (int64_t j = 0; j < n_out; ++j) {
        const double *w = seq + (int64_t)q * j + 2 * H;
        double acc = 0.0;
        for (int t = 0; t < ntaps; ++t) acc += h[t] * w[-t];
        out[j] = acc;
    }
}

/* In-place cascaded-biquad filter over one sequence; sos is [n_sections][6]
 * = b0 b1 b2 a0 a1 a2 (a0 == 1 after scipy's normalisation). */
void oracle_sosfilt(const double *sos, int n_sections, double *x, int64_t n) {
    double zi[64][2];
    for (int s = 0; s < n_sections && s < 64; ++s) zi[s][0] = zi[s][1] = 0.0;
    for (int64_t i = 0; i < n; ++i) {
        double x_cur = x[i];
        for (int s = 0; s < n_sections; ++s) {
            const double *c = sos + 6 * s;
            double x_new = c[0] * x_cur + zi[s][0];
            zi[s][0] = c[1] * x_cur - c[4] * x_new + zi[s][1];
            zi[s][1] = c[2] * x_cur - c[5] * x_new;
            x_cur = x_new;
        }
        x[i] = x_cur;
    }
}

/* seq: [n_ch][n_trials*trial_len] filtered continuous sequences (trial-major in
 * time).  keep[k] != 0 selects trial k.  Writes epochs [rank][ch][ep_len] with
 * rank counting kept epochs in ascending e = k*n_sub + q order. Returns #epochs. */
int64_t oracle_epoch_gather(const double *seq, int n_ch, int64_t n_trials, int64_t trial_len,
                            int n_sub, const uint8_t *keep, double *out) {
    const int64_t ep_len = trial_len / n_sub;
    int64_t rank = 0;
    for (int64_t k = 0; k < n_trials; ++k) {
        if (!keep[k]) continue;
        for (int q = 0; q < n_sub; ++q, ++rank) {
            for (int c = 0; c < n_ch; ++c) {
                const double *src = seq + (int64_t)c * n_trials * trial_len + k * trial_len + q * ep_len;
                double *dst = out + (rank * n_ch + c) * ep_len;
                for (int64_t s = 0; s < ep_len; ++s) dst[s] = src[s];
            }
        }
    }
    return rank;
}

// Host-side construction of the read-only tables (double precision, rounded once to float32).
// Formulae follow the published librosa 0.11 / scipy definitions the reference reaches through
// librosa.feature.mfcc (/root/reference/easywakeword/wakeword.py:561-563); the same definitions are
// restated in numpy in oracle/librosa_restated.py and the two are compared in tests/test_abi.py.
#pragma once
#include <cmath>
#include <cstring>
#include <vector>

#include "ewk_frame.cuh"

namespace ewk {

inline double hz_to_mel_slaney(double f) {
    const double f_sp = 200.0 / 3, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp;
    const double logstep = std::log(6.4) / 27.0;
    return f >= min_log_hz ? min_log_mel + std::log(f / min_log_hz) / logstep : f / f_sp;
}

inline double mel_to_hz_slaney(double m) {
    const double f_sp = 200.0 / 3, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp;
    const double logstep = std::log(6.4) / 27.0;
    return m >= min_log_mel ? min_log_hz * std::exp(logstep * (m - min_log_mel)) : f_sp * m;
}

// dense [128][257] float32 Slaney filterbank, librosa.filters.mel(sr=16000, n_fft=512, n_mels=128)
inline void build_mel_dense(std::vector<float>& mel) {
    const int n_mels = N_MELS, n_bins = N_BINS;
    const double sr = 16000.0, fmin = 0.0, fmax = sr / 2;
    mel.assign((size_t)n_mels * n_bins, 0.f);
    std::vector<double> mel_f(n_mels + 2), fftfreqs(n_bins);
    const double val = 1.0 / (N_FFT * (1.0 / sr));
    for (int k = 0; k < n_bins; k++) fftfreqs[k] = k * val;
    const double lo = hz_to_mel_slaney(fmin), hi = hz_to_mel_slaney(fmax);
    const double step = (hi - lo) / (n_mels + 1);
    for (int i = 0; i < n_mels + 2; i++) mel_f[i] = mel_to_hz_slaney(i == n_mels + 1 ? hi : i * step + lo);
    for (int i = 0; i < n_mels; i++) {
        const double fd0 = mel_f[i + 1] - mel_f[i], fd1 = mel_f[i + 2] - mel_f[i + 1];
        const double enorm = 2.0 / (mel_f[i + 2] - mel_f[i]);
        for (int k = 0; k < n_bins; k++) {
            const double lower = -(mel_f[i] - fftfreqs[k]) / fd0;
            const double upper = (mel_f[i + 2] - fftfreqs[k]) / fd1;
            const double w = std::fmax(0.0, std::fmin(lower, upper));
            const float w32 = (float)w;                       // stored into a float32 array ...
            mel[(size_t)i * n_bins + k] = (float)((double)w32 * enorm);   // ... then scaled in place
        }
    }
}

inline void build_hann(float* w) {
    for (int n = 0; n < N_FFT; n++) w[n] = (float)(0.5 - 0.5 * std::cos(2.0 * M_PI * n / N_FFT));
}

// ortho DCT-II (scipy.fft.dct(type=2, norm='ortho')), first 20 rows, transposed to [b][k]
inline void build_dct_t(float* d) {
    for (int b = 0; b < N_MELS; b++)
        for (int k = 0; k < N_MFCC; k++) {
            const double s = k == 0 ? std::sqrt(1.0 / N_MELS) : std::sqrt(2.0 / N_MELS);
            d[b * N_MFCC + k] = (float)(s * std::cos(M_PI * k * (2 * b + 1) / (2.0 * N_MELS)));
        }
}

inline int build_tables(DeviceTables& T) {
    std::memset(&T, 0, sizeof(T));
    build_hann(T.hann);
    for (int m = 0; m < 256; m++) {
        T.w256[m] = make_float2((float)std::cos(2.0 * M_PI * m / 256), (float)-std::sin(2.0 * M_PI * m / 256));
        T.w512[m] = make_float2((float)std::cos(2.0 * M_PI * m / 512), (float)-std::sin(2.0 * M_PI * m / 512));
    }
    std::vector<float> mel;
    build_mel_dense(mel);
    int off = 0;
    for (int i = 0; i < N_MELS; i++) {
        int first = -1, last = -1;
        for (int k = 0; k < N_BINS; k++)
            if (mel[(size_t)i * N_BINS + k] != 0.f) { if (first < 0) first = k; last = k; }
        if (first < 0) { first = 0; last = -1; }
        const int len = last - first + 1;
        if (off + len > MEL_NNZ_CAP) return -1;
        T.mel_start[i] = first; T.mel_len[i] = len; T.mel_off[i] = off;
        for (int k = 0; k < len; k++) T.mel_w[off + k] = mel[(size_t)i * N_BINS + first + k];
        off += len;
    }
    // interval taps for the mel stage (ewk_frame.cuh: warp_log_mel).  Interval i = [f_i, f_(i+1)) of the mel grid
    // holds the rising edge of band i and the falling edge of band i - 1; lane l, group j owns interval l + 32 j + 1.
    // The weights are the entries of the dense bank above, regrouped: a non-zero tap of band b belongs to interval b
    // when its bin lies below the band's centre f_(b+1), to interval b + 1 otherwise.
    {
        const double sr = 16000.0, fmin = 0.0, fmax = sr / 2;
        std::vector<double> mel_f(N_MELS + 2);
        const double lo = hz_to_mel_slaney(fmin), hi = hz_to_mel_slaney(fmax), step = (hi - lo) / (N_MELS + 1);
        for (int i = 0; i < N_MELS + 2; i++) mel_f[i] = mel_to_hz_slaney(i == N_MELS + 1 ? hi : i * step + lo);
        const double binw = 1.0 / (N_FFT * (1.0 / sr));
        std::vector<int> ifirst(N_MELS + 2, -1), ilast(N_MELS + 2, -1);
        auto interval_of = [&](int b, int k) { return k * binw < mel_f[b + 1] ? b : b + 1; };
        for (int b = 0; b < N_MELS; b++)
            for (int k = 0; k < N_BINS; k++)
                if (mel[(size_t)b * N_BINS + k] != 0.f) {
                    const int i = interval_of(b, k);
                    if (ifirst[i] < 0 || k < ifirst[i]) ifirst[i] = k;
                    if (k > ilast[i]) ilast[i] = k;
                }
        if (ifirst[0] >= 0 || ifirst[N_MELS + 1] >= 0) return -1;       // band 0 has no rising taps, band 127's falling edge is I_128
        const int taps[4] = {MEL_TAPS0, MEL_TAPS1, MEL_TAPS2, MEL_TAPS3};
        int tap0 = 0;
        for (int j = 0; j < 4; j++) {
            for (int lane = 0; lane < 32; lane++) {
                const int i = lane + 32 * j + 1;                           // interval; falling edge of band i - 1, rising edge of band i
                const int first = ifirst[i] < 0 ? 0 : ifirst[i], len = ifirst[i] < 0 ? 0 : ilast[i] - ifirst[i] + 1;
                if (len > taps[j] || first + taps[j] > SCR_P) return -1;
                T.mel_first[j * 32 + lane] = first;
                for (int q = 0; q < taps[j]; q++) {
                    float wf = 0.f, wr = 0.f;
                    const int k = first + q;
                    if (q < len) {
                        const float a = mel[(size_t)(i - 1) * N_BINS + k];
                        if (a != 0.f && interval_of(i - 1, k) == i) wf = a;
                        if (i < N_MELS) {
                            const float r = mel[(size_t)i * N_BINS + k];
                            if (r != 0.f && interval_of(i, k) == i) wr = r;
                        }
                    }
                    T.mel_pad[(tap0 + q) * 32 + lane] = make_float2(wf, wr);
                }
            }
            tap0 += taps[j];
        }
    }
    build_dct_t(T.dct_t);
    return off;
}

}  // namespace ewk

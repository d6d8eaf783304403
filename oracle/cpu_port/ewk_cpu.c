/* ORACLE — TEST / MEASUREMENT INFRASTRUCTURE ONLY.  Not shipped, not on the product path.
 *
 * Best-effort multi-threaded C statement of the gated level-1 + level-2 path for the bench configuration
 * (CPU baseline (iii): "best-effort CPU, not the reference's code path, reported separately so the
 * speed-up is not measured only against Python overhead").  It follows the same reference semantics the numpy oracle
 * (oracle/ewk_oracle.py) restates, function by function:
 *   SoundBuffer._adjust_silence_threshold / is_silent        /root/reference/easywakeword/wakeword.py:472-496
 *   WakeWord._detect_word (timing state machine, segment cut) wakeword.py:1036-1159
 *   WordMatcher.extract_mfcc / calculate_similarity / matches wakeword.py:544-639 (librosa.feature.mfcc restated)
 * restricted to callback blocks of 1600 samples (one storage-order chunk per 100 ms tick) and int16 PCM, which is what
 * bench.py pushes.  Liberties a CPU implementation may take and this one does: no ring copy (the ring content at tick k
 * is pcm[1600 k - R : 1600 k] of the stream it is given), an incrementally maintained sorted chunk array instead of
 * np.percentile per callback, float32 FFT.  Pinned in tests/test_cpu_port.py against the numpy oracle: identical
 * events (tick, segment length, decision), scores within 0.01, features within 1e-4.
 *
 * Tables (Hann window, Slaney mel bank, ortho DCT-II) are passed in from oracle/librosa_restated.py so that they are
 * the oracle's, not a second derivation.   gcc -O3 -march=native -fopenmp -shared -fPIC (oracle/cpu_port/__init__.py). */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define N_FFT 512
#define HOP 160
#define N_BINS 257
#define N_MELS 128
#define N_MFCC 20
#define TICK 1600
#define MAX_FRAMES 301 /* 1 + 48000 / 160: segments longer than 3.0 s are dropped (wakeword.py:1114-1118) */

static float g_hann[N_FFT];
static float g_dct[N_MFCC][N_MELS];
static int g_mel_start[N_MELS], g_mel_len[N_MELS], g_mel_off[N_MELS];
static float g_mel_w[N_MELS * 16];
static float g_cos[N_FFT / 2], g_sin[N_FFT / 2]; /* exp(-2 pi i k / 512) */
static int g_rev[N_FFT / 2];
static int g_ready = 0;

/* mel: dense [128][257] float32; dct: [20][128] float32; hann: [512] float64 */
int ewk_cpu_init(const float* mel, const float* dct, const double* hann) {
    int off = 0;
    for (int b = 0; b < N_MELS; b++) {
        int s = -1, e = -1;
        for (int k = 0; k < N_BINS; k++)
            if (mel[b * N_BINS + k] != 0.f) { if (s < 0) s = k; e = k; }
        if (s < 0) { s = 0; e = -1; }
        if (e - s + 1 > 16) return -1;
        g_mel_start[b] = s; g_mel_len[b] = e - s + 1; g_mel_off[b] = off;
        for (int k = s; k <= e; k++) g_mel_w[off++] = mel[b * N_BINS + k];
    }
    memcpy(g_dct, dct, sizeof(g_dct));
    for (int i = 0; i < N_FFT; i++) g_hann[i] = (float)hann[i];
    for (int k = 0; k < N_FFT / 2; k++) {
        g_cos[k] = (float)cos(-2.0 * M_PI * k / N_FFT);
        g_sin[k] = (float)sin(-2.0 * M_PI * k / N_FFT);
        int r = 0;
        for (int b = 0; b < 8; b++) r |= ((k >> b) & 1) << (7 - b);
        g_rev[k] = r;
    }
    g_ready = 1;
    return 0;
}

/* 512 real samples -> power spectrum P[257]: 256-point complex FFT of z[n] = x[2n] + i x[2n+1] (radix 2, in place),
 * then the real-FFT untangle. */
static void power_spectrum(const float* x, float* P) {
    float re[N_FFT / 2], im[N_FFT / 2];
    for (int n = 0; n < N_FFT / 2; n++) { re[g_rev[n]] = x[2 * n]; im[g_rev[n]] = x[2 * n + 1]; }
    for (int len = 2; len <= N_FFT / 2; len <<= 1) {
        const int half = len >> 1, step = N_FFT / len; /* twiddle exp(-2 pi i j / len) = table[j * step] */
        for (int i = 0; i < N_FFT / 2; i += len)
            for (int j = 0; j < half; j++) {
                const float wr = g_cos[j * step], wi = g_sin[j * step];
                const float ur = re[i + j], ui = im[i + j];
                const float vr = re[i + j + half] * wr - im[i + j + half] * wi;
                const float vi = re[i + j + half] * wi + im[i + j + half] * wr;
                re[i + j] = ur + vr; im[i + j] = ui + vi;
                re[i + j + half] = ur - vr; im[i + j + half] = ui - vi;
            }
    }
    for (int k = 0; k <= N_FFT / 4; k++) {
        const int m = (N_FFT / 2 - k) & (N_FFT / 2 - 1);
        const float er = 0.5f * (re[k] + re[m]), ei = 0.5f * (im[k] - im[m]);
        const float or_ = 0.5f * (im[k] + im[m]), oi = -0.5f * (re[k] - re[m]);
        const float tr = g_cos[k] * or_ - g_sin[k] * oi, ti = g_cos[k] * oi + g_sin[k] * or_;
        const float ar = er + tr, ai = ei + ti, br = er - tr, bi = ei - ti;
        P[k] = ar * ar + ai * ai;
        P[N_FFT / 2 - k] = br * br + bi * bi;
    }
    P[N_FFT / 2] = (re[0] - im[0]) * (re[0] - im[0]);
    P[0] = (re[0] + im[0]) * (re[0] + im[0]);
}

/* WordMatcher.extract_mfcc on int16 PCM (x = q / 32768): mean[20], std[20] (ddof 0) over 1 + n / 160 frames. */
int ewk_cpu_features(const int16_t* q, int n, float* mean, float* stdv) {
    if (!g_ready || n < 1) return -1;
    const int F = 1 + n / HOP;
    if (F > MAX_FRAMES) return -2;
    static __thread float lm[MAX_FRAMES][N_MELS];
    float gmax = -INFINITY;
    for (int t = 0; t < F; t++) {
        float x[N_FFT], P[N_BINS];
        const int f0 = t * HOP - N_FFT / 2;
        for (int i = 0; i < N_FFT; i++) {
            const int p = f0 + i;
            x[i] = (p >= 0 && p < n) ? (float)q[p] * (1.0f / 32768.0f) * g_hann[i] : 0.f;
        }
        power_spectrum(x, P);
        for (int b = 0; b < N_MELS; b++) {
            float s = 0.f;
            const float* w = g_mel_w + g_mel_off[b];
            const float* p = P + g_mel_start[b];
            for (int k = 0; k < g_mel_len[b]; k++) s += w[k] * p[k];
            const float db = 10.0f * log10f(s > 1e-10f ? s : 1e-10f);
            lm[t][b] = db;
            if (db > gmax) gmax = db;
        }
    }
    const float floor_db = gmax - 80.0f; /* librosa.power_to_db(top_db=80): floor against the global maximum */
    double sum[N_MFCC] = {0}, sq[N_MFCC] = {0};
    static __thread float mf[MAX_FRAMES][N_MFCC];
    for (int t = 0; t < F; t++) {
        float v[N_MELS];
        for (int b = 0; b < N_MELS; b++) v[b] = lm[t][b] > floor_db ? lm[t][b] : floor_db;
        for (int k = 0; k < N_MFCC; k++) {
            float c = 0.f;
            for (int b = 0; b < N_MELS; b++) c += g_dct[k][b] * v[b];
            mf[t][k] = c;
            sum[k] += c;
        }
    }
    for (int k = 0; k < N_MFCC; k++) {
        const double mu = sum[k] / F;
        for (int t = 0; t < F; t++) { const double d = mf[t][k] - mu; sq[k] += d * d; }
        mean[k] = (float)mu;
        stdv[k] = (float)sqrt(sq[k] / F);
    }
    return F;
}

/* WordMatcher.calculate_similarity (wakeword.py:615-623) on feature vectors. */
static float one_minus_cosine(const float* u, const float* v) {
    double uv = 0, uu = 0, vv = 0;
    for (int k = 0; k < N_MFCC; k++) { uv += (double)u[k] * v[k]; uu += (double)u[k] * u[k]; vv += (double)v[k] * v[k]; }
    double dist = 1.0 - uv / sqrt(uu * vv);
    if (dist < 0) dist = 0; else if (dist > 2) dist = 2;
    return (float)(1.0 - dist);
}

float ewk_cpu_score(const float* ref_mean, const float* ref_std, const float* mean, const float* stdv) {
    const float p = 100.0f * (0.7f * one_minus_cosine(ref_mean, mean) + 0.3f * one_minus_cosine(ref_std, stdv));
    return p * sqrtf(p) / 10.0f;
}

typedef struct {
    double similarity_threshold, pre_speech_silence, speech_duration_min, speech_duration_max, post_speech_silence, timeout;
    int ring_samples;
} ewk_cpu_params;

typedef struct { int32_t tick, seg_len, matched; float score; } ewk_cpu_event;

/* np.percentile(rms, 25) with numpy's linear _lerp, from the ascending mean-square array */
static double percentile25(const double* S, int n) {
    const double vi = n * 0.25 - 0.25;
    const int lo = (int)floor(vi), hi = lo + 1 < n ? lo + 1 : n - 1;
    const double a = sqrt(S[lo]), b = sqrt(S[hi]), t = vi - lo, d = b - a;
    return t >= 0.5 ? b - d * (1.0 - t) : a + d * t;
}

/* One stream under the audio clock (tick k at time k * 0.1, 1600 new samples per tick, one callback per tick).
 * Returns the number of level-2 events written to ev (<= cap); *n_timeouts counts TimeoutError restarts. */
int ewk_cpu_detect_stream(const int16_t* pcm, int64_t n, const float* ref_mean, const float* ref_std,
                          const ewk_cpu_params* prm, ewk_cpu_event* ev, int cap, int* n_timeouts) {
    const int R = prm->ring_samples, NC = R / TICK;
    if (!g_ready || NC < 1 || NC > 4096) return -1;
    double* ms = (double*)malloc(sizeof(double) * 2 * NC); /* storage-order chunk mean squares, and the same sorted */
    double* so = ms + NC;
    const int64_t n_ticks = n / TICK;
    int n_ev = 0, n_to = 0;
    double thr = 0.01; /* SoundBuffer.silence_threshold until the ring is full (wakeword.py:431) */
    enum { WAITING, IN_SILENCE, IN_SOUND, AFTER_SOUND } state = WAITING;
    int started = 0, last_silent = 1;
    double start_time = 0, silence_start = 0, sound_start = 0, sound_end = 0;
    for (int64_t k = 1; k <= n_ticks; k++) {
        /* the callback of this tick: chunk (k - 1) % NC of the storage order gets the new block */
        const int16_t* blk = pcm + (k - 1) * TICK;
        int64_t acc = 0;
        for (int i = 0; i < TICK; i++) acc += (int32_t)blk[i] * (int32_t)blk[i];
        const double nv = ((double)acc * (1.0 / 1073741824.0)) / (double)TICK; /* np.mean(frame ** 2) */
        const int c = (int)((k - 1) % NC);
        const int full = k * (int64_t)TICK >= R;
        if (k <= NC) {
            ms[c] = nv;
            if (k == NC) { /* first full ring: sort once */
                memcpy(so, ms, sizeof(double) * NC);
                for (int i = 1; i < NC; i++) { const double v = so[i]; int j = i - 1; while (j >= 0 && so[j] > v) { so[j + 1] = so[j]; j--; } so[j + 1] = v; }
            }
        } else { /* replace the chunk's old value in the sorted array */
            const double ov = ms[c];
            ms[c] = nv;
            int i = 0;
            while (so[i] != ov) i++;
            if (nv > ov) { while (i + 1 < NC && so[i + 1] < nv) { so[i] = so[i + 1]; i++; } }
            else { while (i > 0 && so[i - 1] > nv) { so[i] = so[i - 1]; i--; } }
            so[i] = nv;
        }
        if (full) { const double t = percentile25(so, NC) * 1.5; thr = t > 0.005 ? t : 0.005; }
        const double rms = sqrt(nv); /* is_silent: the last 1600 samples are this block */
        const int silent = rms < thr;
        const double now = (double)k * 0.1;
        if (full && !started) { /* _wait_for_buffer done -> _detect_word entry (wakeword.py:1048-1057) */
            started = 1;
            state = silent ? IN_SILENCE : WAITING;
            start_time = now;
            if (silent) silence_start = now;
            last_silent = silent;
            continue;
        }
        if (!started) { last_silent = silent; continue; }
        const double prev = (double)(k - 1) * 0.1;
        if (prm->timeout > 0.0 && prev - start_time > prm->timeout) { /* TimeoutError, listen loop re-enters at prev */
            n_to++;
            state = last_silent ? IN_SILENCE : WAITING;
            start_time = prev;
            if (last_silent) silence_start = prev;
        }
        switch (state) {
            case WAITING: if (silent) { state = IN_SILENCE; silence_start = now; } break;
            case IN_SILENCE:
                if (!silent) {
                    if (now - silence_start >= prm->pre_speech_silence) { state = IN_SOUND; sound_start = now; }
                    else state = WAITING;
                }
                break;
            case IN_SOUND: {
                const double d = now - sound_start;
                if (!silent) { if (d > prm->speech_duration_max) state = WAITING; }
                else if (d >= prm->speech_duration_min && d <= prm->speech_duration_max) { state = AFTER_SOUND; sound_end = now; }
                else state = WAITING;
                break;
            }
            case AFTER_SOUND:
                if (!silent) { state = WAITING; break; }
                if (now - sound_end >= prm->post_speech_silence) {
                    int64_t n_back = (int64_t)(fabs(sound_start - now - 0.05) * 16000.0);
                    const int64_t n_drop = (int64_t)(fabs(sound_end - now + 0.05) * 16000.0);
                    if (n_back > R) n_back = R;
                    const int64_t len = n_back - n_drop;
                    state = WAITING;
                    if (len >= 1 && (double)len / 16000.0 <= 3.0) {
                        float mean[N_MFCC], stdv[N_MFCC];
                        ewk_cpu_features(pcm + k * (int64_t)TICK - n_back, (int)len, mean, stdv);
                        const float sc = ewk_cpu_score(ref_mean, ref_std, mean, stdv);
                        if (n_ev < cap) { ev[n_ev].tick = (int32_t)k; ev[n_ev].seg_len = (int32_t)len; ev[n_ev].score = sc;
                                          ev[n_ev].matched = sc >= (float)prm->similarity_threshold; }
                        n_ev++;
                    }
                }
                break;
        }
        last_silent = silent;
    }
    free(ms);
    if (n_timeouts) *n_timeouts = n_to;
    return n_ev;
}

/* n_streams independent streams, pcm[s * stride .. + n), over all OpenMP threads.  counts[s] = level-2 events. */
int ewk_cpu_detect_batch(const int16_t* pcm, int n_streams, int64_t n, int64_t stride, const float* ref_mean,
                         const float* ref_std, const ewk_cpu_params* prm, ewk_cpu_event* ev, int cap_per_stream,
                         int32_t* counts, int n_threads) {
    int bad = 0;
#pragma omp parallel for schedule(dynamic, 1) num_threads(n_threads) reduction(| : bad)
    for (int s = 0; s < n_streams; s++) {
        int to = 0;
        const int r = ewk_cpu_detect_stream(pcm + (int64_t)s * stride, n, ref_mean, ref_std, prm,
                                            ev + (int64_t)s * cap_per_stream, cap_per_stream, &to);
        counts[s] = r;
        if (r < 0) bad = 1;
    }
    return bad ? -1 : 0;
}

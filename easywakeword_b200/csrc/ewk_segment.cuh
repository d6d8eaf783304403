// K3 — segment_mfcc_match: one CTA turns one PCM segment into MFCC statistics and template scores
// without leaving the SM.  Replaces WordMatcher.extract_mfcc / calculate_similarity / matches
// (/root/reference/easywakeword/wakeword.py:544-567, 591-639) for a batch of segments.
//
//   phase A  warp per frame: PCM (linear or ring, f32 or i16, zero outside the segment = librosa's
//            center=True / pad_mode='constant') -> window -> FFT -> |X|^2 -> mel -> log  => smem
//   phase B  block max -> power_to_db floor  c = max - 80  (top_db couples every frame of a segment)
//   phase C  ortho DCT-II of max(logmel, c), first 20 coefficients                          => smem
//   phase D  mean / std over frames (two-pass, ddof 0), cosine vs each template, p^1.5/10, >= thr
//
// The log-mel matrix (<= 301 x 128 for the reference's 3.0 s cap, wakeword.py:1114-1118) stays in
// shared memory; longer inputs (set_reference on arbitrary audio) spill to a global workspace.
#pragma once
#include "ewk_frame.cuh"

namespace ewk {

constexpr int SEG_THREADS = 448;           // 14 warps: two CTAs leave 8 k registers per SM for the bulk push that runs beside K3;
                                           // the frame pipeline is shared-memory bound, so 28 warps per SM are as fast as 32
constexpr int SEG_WARPS = SEG_THREADS / 32;
constexpr int SEG_SMEM_FRAMES = 301;           // 1 + 48000/160
constexpr int FEAT = 2 * N_MFCC;               // mean[20] ++ std[20]

struct SegDesc {
    const void* base;      // device PCM
    long long start;       // first sample: linear index, or ring position when ring > 0
    int ring;              // physical ring length in samples (0: linear buffer)
    int len;               // samples in the segment (>= 1)
    int fmt;               // 0 f32, 1 i16
    int ws_frame_off;      // frame offset into the global workspace (used when frames > smem cap)
    long long frames_off;  // frame offset into frames_out (when requested)
    long long lm_off;      // frame offset into the log-mel workspace
};

constexpr int LM_ROW = N_MELS;                 // floats per frame in the log-mel workspace

struct TemplateFeat {
    float mean[N_MFCC];
    float std[N_MFCC];
    long long n_samples;
    int valid;
    int pad;
    float dmean[N_MFCC];   // the same features from the dense kernel's exact integer statistics (ewk_dense.cuh), so that a
    float dstd[N_MFCC];    // dense window that IS the template scores exactly 100.0
};

struct PcmReader {
    const float* f;
    const short* q;
    int ring;
    int len;
    long long start;
    float pre;             // pre-emphasis coefficient (0: none)
    __device__ __forceinline__ float at(int i) const {   // i relative to the segment
        if (i < 0 || i >= len) return 0.f;
        long long p = start + i;
        if (ring) { if (p >= ring) p -= ring; }
        return q ? (float)q[p] * (1.0f / 32768.0f) : f[p];
    }
    // Frame starting at segment-relative f0: can lanes read sample pairs (2 lane + 64 a, +1) with one
    // aligned 4- / 8-byte load each, without bounds or wrap checks?  Returns the buffer position of
    // the frame's first sample, or -1.
    __device__ __forceinline__ long long fast_base(int f0) const {
        if (f0 < 0 || f0 + N_FFT > len) return -1;
        long long p = start + f0;
        if (ring) { if (p >= ring) p -= ring; if (p + N_FFT > ring) return -1; }
        if (p & 1) return -1;
        if (q ? ((size_t)q & 3) : ((size_t)f & 7)) return -1;
        return p;
    }
    // Same, for frames that start on an ODD buffer position (segments are cut at arbitrary samples): int16 pairs
    // are assembled from two aligned words (the second reaches one sample past the frame, which must exist),
    // float pairs from two scalar loads.
    __device__ __forceinline__ long long odd_base(int f0) const {
        if (f0 < 0 || f0 + N_FFT > len) return -1;
        long long p = start + f0;
        if (ring) { if (p >= ring) p -= ring; if (p + N_FFT > ring) return -1; }
        if (!(p & 1)) return -1;
        if (q) {
            if ((size_t)q & 3) return -1;
            if (!ring && f0 + N_FFT + 1 > len) return -1;     // rings are even-sized: sample p + 512 is inside
        }
        return p;
    }
};

// scipy.spatial.distance.cosine + the reference's mixing and rescale (wakeword.py:615-623), in the
// float32 arithmetic numpy uses for float32 features.  NaN propagates exactly as in the reference
// (0/0 for a constant feature vector; NaN >= thr is False).
__device__ __forceinline__ float one_minus_cosine(const float* u, const float* v) {
    float uv = 0.f, uu = 0.f, vv = 0.f;
#pragma unroll
    for (int k = 0; k < N_MFCC; k++) {
        uv = fmaf(u[k], v[k], uv);
        uu = fmaf(u[k], u[k], uu);
        vv = fmaf(v[k], v[k], vv);
    }
    const float prod = __fmul_rn(uu, vv);
    const float q = (float)((double)uv / sqrt((double)prod));
    float dist = __fsub_rn(1.0f, q);
    dist = dist < 0.f ? 0.f : (dist > 2.f ? 2.f : dist);   // np.clip keeps NaN
    return __fsub_rn(1.0f, dist);
}

// The same score with one feature component per lane (lanes >= 20 pass zeros); every lane returns it.  The three
// dot products of each cosine are butterfly sums, so u == v gives uv == uu == vv bit for bit and a self-match
// scores exactly 100.
__device__ __forceinline__ float one_minus_cosine_warp(float u, float v) {
    float uv = u * v, uu = u * u, vv = v * v;
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        uv += __shfl_xor_sync(0xffffffffu, uv, o);
        uu += __shfl_xor_sync(0xffffffffu, uu, o);
        vv += __shfl_xor_sync(0xffffffffu, vv, o);
    }
    const float prod = __fmul_rn(uu, vv);
    const float q = (float)((double)uv / sqrt((double)prod));
    float dist = __fsub_rn(1.0f, q);
    dist = dist < 0.f ? 0.f : (dist > 2.f ? 2.f : dist);
    return __fsub_rn(1.0f, dist);
}

__device__ __forceinline__ float similarity_score_warp(float ref_mean, float ref_std, float mean, float std) {
    const float sim_mean = one_minus_cosine_warp(ref_mean, mean);
    const float sim_std = one_minus_cosine_warp(ref_std, std);
    const float combined = __fadd_rn(__fmul_rn(sim_mean, 0.7f), __fmul_rn(sim_std, 0.3f));
    const float p = __fmul_rn(combined, 100.0f);
    return __fdiv_rn(__fmul_rn(p, sqrtf(p)), 10.0f);
}

// the same from the six dot products (u = template, v = candidate; mean pair, then std pair)
__device__ __forceinline__ float score_from_dots(float uvm, float uum, float vvm, float uvs, float uus, float vvs) {
    const float qm = (float)((double)uvm / sqrt((double)__fmul_rn(uum, vvm)));
    float dm = __fsub_rn(1.0f, qm);
    dm = dm < 0.f ? 0.f : (dm > 2.f ? 2.f : dm);
    const float qs = (float)((double)uvs / sqrt((double)__fmul_rn(uus, vvs)));
    float ds = __fsub_rn(1.0f, qs);
    ds = ds < 0.f ? 0.f : (ds > 2.f ? 2.f : ds);
    const float combined = __fadd_rn(__fmul_rn(__fsub_rn(1.0f, dm), 0.7f), __fmul_rn(__fsub_rn(1.0f, ds), 0.3f));
    const float p = __fmul_rn(combined, 100.0f);
    return __fdiv_rn(__fmul_rn(p, sqrtf(p)), 10.0f);
}

__device__ __forceinline__ float similarity_score(const float* ref_mean, const float* ref_std,
                                                  const float* mean, const float* std) {
    const float sim_mean = one_minus_cosine(ref_mean, mean);
    const float sim_std = one_minus_cosine(ref_std, std);
    const float combined = __fadd_rn(__fmul_rn(sim_mean, 0.7f), __fmul_rn(sim_std, 0.3f));
    const float p = __fmul_rn(combined, 100.0f);
    return __fdiv_rn(__fmul_rn(p, sqrtf(p)), 10.0f);      // p**1.5 / 100**0.5
}

// Dynamic shared memory layout:
//   FrameTables | scratch[SEG_WARPS*SCR_WARP] | part[SEG_PART] | mfcc[cap*20] | fmin[cap] | fmax[cap]
constexpr int FR_STRIDE = N_MFCC + 2;          // per-frame record in the global spill: mfcc[20], min, max
constexpr int SEG_PART = SEG_WARPS * 2 * N_MFCC + 2 * N_MFCC + 32;   // per-warp (mean, M2)[20], feat[40], red[32]
__host__ __device__ inline size_t seg_smem_bytes(int cap_frames) {
    return sizeof(FrameTables) + sizeof(float) * ((size_t)SEG_WARPS * SCR_WARP + SEG_PART + (size_t)cap_frames * FR_STRIDE);
}

struct SegSmem {
    FrameTables* ft;
    float *scratch, *part, *feat, *red, *mf, *fmin, *fmax;
};

__device__ __forceinline__ SegSmem seg_carve(float* smem, int cap_frames) {
    SegSmem m;
    m.ft = reinterpret_cast<FrameTables*>(smem);
    m.scratch = smem + sizeof(FrameTables) / sizeof(float);
    m.part = m.scratch + SEG_WARPS * SCR_WARP;
    m.feat = m.part + SEG_WARPS * 2 * N_MFCC;
    m.red = m.feat + 2 * N_MFCC;
    m.mf = m.part + SEG_PART;
    m.fmin = m.mf + (size_t)cap_frames * N_MFCC;
    m.fmax = m.fmin + cap_frames;
    return m;
}

// once per CTA: tables into shared memory, the lane's mel band descriptors into registers
__device__ __forceinline__ void seg_prologue(const DeviceTables* __restrict__ T, const SegSmem& m) {
    copy_frame_tables(*m.ft, T, threadIdx.x, SEG_THREADS);
    __syncthreads();
}

// The lane's eight sample pairs (2 lane + 64 a, +1) of frame t, zero outside the segment
// (librosa center=True, pad_mode='constant').
__device__ __forceinline__ void load_frame_pairs_raw(const PcmReader& rd, int f0, int lane, float2 (&x)[8]) {
    const long long base = rd.fast_base(f0);
    if (base >= 0) {
        if (rd.q) {
            const unsigned* w = reinterpret_cast<const unsigned*>(rd.q + base);
#pragma unroll
            for (int a = 0; a < 8; a++) {
                const unsigned u = __ldg(w + lane + 32 * a);
                x[a] = make_float2((float)(short)(u & 0xffff) * (1.0f / 32768.0f), (float)((int)u >> 16) * (1.0f / 32768.0f));
            }
        } else {
            const float2* w = reinterpret_cast<const float2*>(rd.f + base);
#pragma unroll
            for (int a = 0; a < 8; a++) x[a] = __ldg(w + lane + 32 * a);
        }
        return;
    }
    const long long ob = rd.odd_base(f0);
    if (ob >= 0) {
        if (rd.q) {
            const unsigned* w = reinterpret_cast<const unsigned*>(rd.q + (ob - 1));
#pragma unroll
            for (int a = 0; a < 8; a++) {
                const unsigned lo = __ldg(w + lane + 32 * a), hi = __ldg(w + lane + 32 * a + 1);
                x[a] = make_float2((float)((int)lo >> 16) * (1.0f / 32768.0f), (float)(short)(hi & 0xffff) * (1.0f / 32768.0f));
            }
        } else {
            const float* w = rd.f + ob + 2 * lane;
#pragma unroll
            for (int a = 0; a < 8; a++) x[a] = make_float2(__ldg(w + 64 * a), __ldg(w + 64 * a + 1));
        }
        return;
    }
    // masked frame (window edges, ring wrap).  When the frame starts on an even buffer position every pair
    // is one aligned 4- / 8-byte load (a pair never straddles the ring end: ring length is even); the mask
    // is applied per sample.
    const long long p00 = rd.start + f0;                        // may be negative by up to N_FFT/2
    const bool even = ((p00 & 1) == 0) && (rd.ring == 0 || (rd.ring & 1) == 0) &&
                      !(rd.q ? ((size_t)rd.q & 3) : ((size_t)rd.f & 7));
    if (even && rd.ring) {
        // ring reader: every ring position is valid memory, so the pairs are loaded unconditionally (one wrap) and the
        // window mask is applied afterwards — frame-relative sample m is inside the segment iff mlo <= m < mhi
        int p0 = (int)p00;
        if (p0 < 0) p0 += rd.ring; else if (p0 >= rd.ring) p0 -= rd.ring;
        const unsigned mlo = (unsigned)max(0, -f0), span = (unsigned)max(0, min(N_FFT, rd.len - f0) - (int)mlo);
#pragma unroll
        for (int a = 0; a < 8; a++) {
            const int m = 2 * lane + 64 * a;
            int p = p0 + m;
            if (p >= rd.ring) p -= rd.ring;
            float2 v;
            if (rd.q) {
                const unsigned u = __ldg(reinterpret_cast<const unsigned*>(rd.q + p));
                v = make_float2((float)(short)(u & 0xffff) * (1.0f / 32768.0f), (float)((int)u >> 16) * (1.0f / 32768.0f));
            } else v = __ldg(reinterpret_cast<const float2*>(rd.f + p));
            x[a] = make_float2((unsigned)m - mlo < span ? v.x : 0.f, (unsigned)(m + 1) - mlo < span ? v.y : 0.f);
        }
        return;
    }
    if (even) {
        int p0 = (int)p00;                                       // |p00| < 2^31 for rings; linear buffers < 2^30 samples
        if (rd.ring) { if (p0 < 0) p0 += rd.ring; else if (p0 >= rd.ring) p0 -= rd.ring; }
#pragma unroll
        for (int a = 0; a < 8; a++) {
            const int off = 2 * lane + 64 * a;
            const int i = f0 + off;
            float2 v = make_float2(0.f, 0.f);
            if (i >= 0 && i < rd.len) {
                int p = p0 + off;
                if (rd.ring && p >= rd.ring) p -= rd.ring;
                if (i + 1 < rd.len) {
                    if (rd.q) {
                        const unsigned u = __ldg(reinterpret_cast<const unsigned*>(rd.q + p));
                        v = make_float2((float)(short)(u & 0xffff) * (1.0f / 32768.0f), (float)((int)u >> 16) * (1.0f / 32768.0f));
                    } else v = __ldg(reinterpret_cast<const float2*>(rd.f + p));
                } else {                                          // last sample of an odd-length window: never read past it
                    v.x = rd.q ? (float)__ldg(rd.q + p) * (1.0f / 32768.0f) : __ldg(rd.f + p);
                }
            }
            x[a] = v;
        }
        return;
    }
#pragma unroll
    for (int a = 0; a < 8; a++) {
        const int i = f0 + 2 * lane + 64 * a;
        x[a] = make_float2(rd.at(i), rd.at(i + 1));
    }
}

// Pre-emphasis of the segment, applied to the frame's pairs in registers: what librosa.effects.preemphasis(y, coef=a)
// returns for the segment y, sampled at the frame's positions (the reference calls librosa.feature.mfcc on the raw
// segment, wakeword.py:561-563: a = 0 and this function is never entered).  scipy.signal.lfilter([1, -a], [1], y,
// zi = 2 y[0] - y[1]) in float32, direct form II transposed: out[n] = z + 1 * y[n], z' = -a * y[n]  =>
//     out[0] = (2 y[0] - y[1]) + y[0],      out[n] = fl(-a y[n-1]) + y[n]   (n >= 1),      zero outside the segment
// (centring pads AFTER the filter).  f0 is even (frames start at 160 t - 256; the dense kernel's views keep that), so
// segment sample 0 is always the even member of its pair.
__device__ __forceinline__ void preemph_pairs(const PcmReader& rd, int f0, int lane, float2 (&x)[8]) {
    const float nb = -rd.pre;
    float carry = rd.at(f0 - 1);                                 // raw sample before the frame (0 outside the segment)
#pragma unroll
    for (int k = 0; k < 8; k++) {
        float prev = __shfl_up_sync(FULL, x[k].y, 1);
        const float last = __shfl_sync(FULL, x[k].y, 31);
        if (lane == 0) prev = carry;
        carry = last;
        const int i = f0 + 2 * lane + 64 * k;                    // segment index of the pair's even sample
        float ye = __fadd_rn(__fmul_rn(nb, prev), x[k].x);
        float yo = __fadd_rn(__fmul_rn(nb, x[k].x), x[k].y);
        if (i == 0) ye = __fadd_rn(__fsub_rn(__fmul_rn(2.0f, x[k].x), rd.len > 1 ? x[k].y : 0.f), x[k].x);
        if (i < 0 || i >= rd.len) ye = 0.f;
        if (i + 1 < 0 || i + 1 >= rd.len) yo = 0.f;
        x[k] = make_float2(ye, yo);
    }
}

// PRE selects the kernel instantiation with pre-emphasis (launched only when ewk_config.preemphasis != 0), so the
// reference-parity build of every kernel carries none of it.
template <bool PRE>
__device__ __forceinline__ void load_frame_pairs_at(const PcmReader& rd, int f0, int lane, float2 (&x)[8]) {
    load_frame_pairs_raw(rd, f0, lane, x);
    if (PRE) preemph_pairs(rd, f0, lane, x);
}

template <bool PRE>
__device__ __forceinline__ void load_frame_pairs(const PcmReader& rd, int t, int lane, float2 (&x)[8]) {
    load_frame_pairs_at<PRE>(rd, t * HOP - N_FFT / 2, lane, x);
}

// One segment by the whole CTA (all threads must call it); returns a shared-memory pointer to mean[20] ++ std[20]
// (valid until the next call).  Warp w owns frames w, w + 16, ...:
//   A  every frame -> MFCC without the power_to_db floor, plus the frame's log-mel min / max          | barrier
//   B  floor = (max over frames of the log-mel max) - 80 (librosa.power_to_db(top_db=80) couples all frames of a
//      segment); every warp reduces the F maxima itself (F is small; long inputs use a block reduction)
//   C  the warp's own frames whose min lies below the floor are redone with it (none in the common case): from the
//      stored log-mel row (lm, one DCT) or, without the workspace, from the PCM
//   D  mean / std over frames (ddof 0): two-pass mean and M2 over the warp's own frames                | barrier
//      then one warp pools the 16 partials (Chan et al.), in fixed order                               | barrier
// lm: this segment's rows of the log-mel workspace ([F][LM_ROW] floats, L2-resident scratch) or null.
template <bool PRE>
__device__ __forceinline__ float* segment_features(const SegDesc& sd, const SegSmem& m,
                                                   int cap_frames, float* __restrict__ ws,
                                                   float* __restrict__ frames_out, float* __restrict__ lm) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int F = 1 + sd.len / HOP;
    float* mf = m.mf;
    float* fmn = m.fmin;
    float* fmx = m.fmax;
    if (F > cap_frames) {
        mf = ws + (size_t)sd.ws_frame_off * FR_STRIDE;
        fmn = mf + (size_t)F * N_MFCC;
        fmx = fmn + F;
    }
    PcmReader rd;
    rd.f = sd.fmt == 0 ? (const float*)sd.base : nullptr;
    rd.q = sd.fmt == 1 ? (const short*)sd.base : nullptr;
    rd.ring = sd.ring; rd.len = sd.len; rd.start = sd.start; rd.pre = PRE ? m.ft->preemph : 0.f;
    float* scr = m.scratch + warp * SCR_WARP;
    // ---- A
    for (int t = warp; t < F; t += SEG_WARPS) {
        float2 x[8];
        load_frame_pairs<PRE>(rd, t, lane, x);
        float mn, mx;
        warp_frame_mfcc(x, *m.ft, scr, lane, -INFINITY, mf + (size_t)t * N_MFCC, mn, mx,
                        lm ? lm + (size_t)t * LM_ROW : nullptr);
        if (lane == 0) { fmn[t] = mn; fmx[t] = mx; }
    }
    __syncthreads();
    // ---- B
    float vmax = -INFINITY;
    if (F <= 64 * 32) {
        for (int t = lane; t < F; t += 32) vmax = fmaxf(vmax, fmx[t]);
#pragma unroll
        for (int o = 16; o; o >>= 1) vmax = fmaxf(vmax, __shfl_xor_sync(FULL, vmax, o));
    } else {
        for (int t = tid; t < F; t += SEG_THREADS) vmax = fmaxf(vmax, fmx[t]);
#pragma unroll
        for (int o = 16; o; o >>= 1) vmax = fmaxf(vmax, __shfl_xor_sync(FULL, vmax, o));
        if (lane == 0) m.red[warp] = vmax;
        __syncthreads();
        vmax = m.red[0];
#pragma unroll
        for (int w = 1; w < SEG_WARPS; w++) vmax = fmaxf(vmax, m.red[w]);
    }
    const float floor_db = vmax - 80.0f;                      // librosa.power_to_db(top_db=80)
    // ---- C + D over the warp's own frames
    float sum = 0.f;
    int nw = 0;
    for (int t = warp; t < F; t += SEG_WARPS, nw++) {
        if (fmn[t] < floor_db) {                              // warp-uniform
            if (lm) warp_refloor_mfcc(lm + (size_t)t * LM_ROW, m.ft, floor_db, mf + (size_t)t * N_MFCC);
            else {
                float2 x[8];
                load_frame_pairs<PRE>(rd, t, lane, x);
                float mn, mx;
                warp_frame_mfcc(x, *m.ft, scr, lane, floor_db, mf + (size_t)t * N_MFCC, mn, mx);
            }
            __syncwarp();
        }
        if (lane < N_MFCC) {
            const float v = mf[(size_t)t * N_MFCC + lane];
            sum += v;
            if (frames_out) frames_out[(sd.frames_off + t) * N_MFCC + lane] = v;
        }
    }
    if (lane < N_MFCC) {
        const float mu = nw ? sum / (float)nw : 0.f;
        float m2 = 0.f;
        for (int t = warp; t < F; t += SEG_WARPS) { const float d = mf[(size_t)t * N_MFCC + lane] - mu; m2 = fmaf(d, d, m2); }
        m.part[warp * 2 * N_MFCC + lane] = mu;
        m.part[warp * 2 * N_MFCC + N_MFCC + lane] = m2;
    }
    __syncthreads();
    if (tid < N_MFCC) {
        float n = 0.f, mean = 0.f, M2 = 0.f;
#pragma unroll 1
        for (int w = 0; w < SEG_WARPS; w++) {
            if (w >= F) break;
            const float cw = (float)((F - w + SEG_WARPS - 1) / SEG_WARPS);
            const float delta = m.part[w * 2 * N_MFCC + tid] - mean, nn = n + cw;
            mean = fmaf(delta, cw / nn, mean);
            M2 += m.part[w * 2 * N_MFCC + N_MFCC + tid] + delta * delta * (n * cw / nn);
            n = nn;
        }
        const bool kept = tid < m.ft->n_mfcc;                  // coefficients beyond n_mfcc (ewk_config) drop out of both cosines
        m.feat[tid] = kept ? mean : 0.f;
        m.feat[N_MFCC + tid] = kept ? sqrtf(M2 / (float)F) : 0.f;
    }
    __syncthreads();
    return m.feat;
}

// K3, batch form: one CTA per caller-described segment (extract_mfcc / calculate_similarity / matches).
template <bool PRE>
__global__ void __launch_bounds__(512, 2)
segment_mfcc_match_kernel(const DeviceTables* __restrict__ T, const SegDesc* __restrict__ segs,
                          int cap_frames, float* __restrict__ ws,           // global spill [frames][22]
                          float* __restrict__ lm_ws,                         // log-mel workspace [frames][128] or null
                          const TemplateFeat* __restrict__ tmpl, int n_tmpl, int tmpl_first,
                          float threshold,
                          float* __restrict__ feat_out,                      // [n_seg][40] or null
                          float* __restrict__ frames_out,                    // [frames][20] or null
                          float* __restrict__ scores,                        // [n_seg][n_tmpl] or null
                          unsigned char* __restrict__ matched)               // [n_seg][n_tmpl] or null
{
    extern __shared__ __align__(16) float smem[];
    const SegSmem m = seg_carve(smem, cap_frames);
    seg_prologue(T, m);
    const int tid = threadIdx.x;
    const SegDesc sd = segs[blockIdx.x];
    const float* feat = segment_features<PRE>(sd, m, cap_frames, ws, frames_out,
                                         lm_ws ? lm_ws + (size_t)sd.lm_off * LM_ROW : nullptr);
    if (feat_out && tid < FEAT) feat_out[(size_t)blockIdx.x * FEAT + tid] = feat[tid];
    if (scores && tid < n_tmpl) {
        const TemplateFeat& tf = tmpl[tmpl_first + tid];
        float sc = __int_as_float(0x7fc00000);
        if (tf.valid) sc = similarity_score(tf.mean, tf.std, feat, feat + N_MFCC);
        scores[(size_t)blockIdx.x * n_tmpl + tid] = sc;
        if (matched) matched[(size_t)blockIdx.x * n_tmpl + tid] = (sc >= threshold) ? 1 : 0;
    }
}

}  // namespace ewk

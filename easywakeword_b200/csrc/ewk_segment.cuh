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

constexpr int SEG_THREADS = 256;
constexpr int SEG_WARPS = SEG_THREADS / 32;
constexpr int LM_STRIDE = 129;                 // log-mel row stride (conflict-free column walks)
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
};

struct TemplateFeat {
    float mean[N_MFCC];
    float std[N_MFCC];
    long long n_samples;
    int valid;
    int pad;
};

struct PcmReader {
    const float* f;
    const short* q;
    int ring;
    int len;
    long long start;
    __device__ __forceinline__ float at(int i) const {   // i relative to the segment
        if (i < 0 || i >= len) return 0.f;
        long long p = start + i;
        if (ring) { if (p >= ring) p -= ring; }
        return q ? (float)q[p] * (1.0f / 32768.0f) : f[p];
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

__device__ __forceinline__ float similarity_score(const float* ref_mean, const float* ref_std,
                                                  const float* mean, const float* std) {
    const float sim_mean = one_minus_cosine(ref_mean, mean);
    const float sim_std = one_minus_cosine(ref_std, std);
    const float combined = __fadd_rn(__fmul_rn(sim_mean, 0.7f), __fmul_rn(sim_std, 0.3f));
    const float p = __fmul_rn(combined, 100.0f);
    return __fdiv_rn(__fmul_rn(p, sqrtf(p)), 10.0f);      // p**1.5 / 100**0.5
}

// Dynamic shared memory layout (floats):
//   melw[512] | dct_t[128*20] | scratch[SEG_WARPS*SCR_WARP] | red[64] | logmel[cap*129] | mfcc[cap*20]
__host__ __device__ inline size_t seg_smem_bytes(int cap_frames) {
    return sizeof(float) * ((size_t)MEL_NNZ_CAP + N_MELS * N_MFCC + SEG_WARPS * SCR_WARP + 64 +
                            (size_t)cap_frames * (LM_STRIDE + N_MFCC));
}

struct SegSmem {
    float *melw, *dct, *scratch, *red, *lm, *mf;
};

__device__ __forceinline__ SegSmem seg_carve(float* smem, int cap_frames) {
    SegSmem m;
    m.melw = smem;
    m.dct = m.melw + MEL_NNZ_CAP;
    m.scratch = m.dct + N_MELS * N_MFCC;
    m.red = m.scratch + SEG_WARPS * SCR_WARP;
    m.lm = m.red + 64;
    m.mf = m.lm + (size_t)cap_frames * LM_STRIDE;
    return m;
}

// once per CTA: tables into shared memory, per-lane constants into registers
__device__ __forceinline__ void seg_prologue(const DeviceTables* __restrict__ T, const SegSmem& m, LaneConsts& lc) {
    for (int i = threadIdx.x; i < MEL_NNZ_CAP; i += SEG_THREADS) m.melw[i] = T->mel_w[i];
    for (int i = threadIdx.x; i < N_MELS * N_MFCC; i += SEG_THREADS) m.dct[i] = T->dct_t[i];
    init_lane_consts(lc, T, threadIdx.x & 31);
    __syncthreads();
}

// Phases A-D for one segment by the whole CTA.  Returns a shared-memory pointer to mean[20] ++ std[20]
// (valid until the next call).  All threads must call it.
__device__ __forceinline__ float* segment_features(const SegDesc& sd, const SegSmem& m, const LaneConsts& lc,
                                                   int cap_frames, float* __restrict__ ws,
                                                   float* __restrict__ frames_out) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int F = 1 + sd.len / HOP;
    float* lm = m.lm;
    float* mf = m.mf;
    if (F > cap_frames) {
        lm = ws + (size_t)sd.ws_frame_off * (LM_STRIDE + N_MFCC);
        mf = lm + (size_t)F * LM_STRIDE;
    }
    PcmReader rd;
    rd.f = sd.fmt == 0 ? (const float*)sd.base : nullptr;
    rd.q = sd.fmt == 1 ? (const short*)sd.base : nullptr;
    rd.ring = sd.ring; rd.len = sd.len; rd.start = sd.start;

    // ---- phase A
    float* scr = m.scratch + warp * SCR_WARP;
    float vmax = -INFINITY;
    for (int t = warp; t < F; t += SEG_WARPS) {
        const int f0 = t * HOP - N_FFT / 2;
        warp_power_spectrum([&](int i) { return make_float2(rd.at(f0 + i), rd.at(f0 + i + 1)); }, lc, scr, lane);
        float v[4];
        warp_log_mel(scr + 2 * SCR_PLANE, m.melw, lc, v);
#pragma unroll
        for (int j = 0; j < 4; j++) {
            lm[(size_t)t * LM_STRIDE + lane + 32 * j] = v[j];
            vmax = fmaxf(vmax, v[j]);
        }
    }
    // ---- phase B
#pragma unroll
    for (int o = 16; o; o >>= 1) vmax = fmaxf(vmax, __shfl_xor_sync(FULL, vmax, o));
    if (lane == 0) m.red[warp] = vmax;
    __syncthreads();
    float gmax = m.red[0];
#pragma unroll
    for (int w = 1; w < SEG_WARPS; w++) gmax = fmaxf(gmax, m.red[w]);
    const float floor_db = gmax - 80.0f;       // librosa.power_to_db(top_db=80)
    __syncthreads();

    // ---- phase C: thread = (frame, half of the coefficients)
    for (int it = tid; it < 2 * F; it += SEG_THREADS) {
        const int t = it >> 1, g = it & 1;
        const float* row = lm + (size_t)t * LM_STRIDE;
        float acc[10];
#pragma unroll
        for (int k = 0; k < 10; k++) acc[k] = 0.f;
        for (int b = 0; b < N_MELS; b++) {
            const float x = fmaxf(row[b], floor_db);
            const float2* d = reinterpret_cast<const float2*>(m.dct + b * N_MFCC + 10 * g);
#pragma unroll
            for (int k = 0; k < 5; k++) {
                const float2 w = d[k];
                acc[2 * k] = fmaf(w.x, x, acc[2 * k]);
                acc[2 * k + 1] = fmaf(w.y, x, acc[2 * k + 1]);
            }
        }
#pragma unroll
        for (int k = 0; k < 10; k++) mf[(size_t)t * N_MFCC + 10 * g + k] = acc[k];
        if (frames_out) {
#pragma unroll
            for (int k = 0; k < 10; k++) frames_out[(sd.frames_off + t) * N_MFCC + 10 * g + k] = acc[k];
        }
    }
    __syncthreads();

    // ---- phase D: mean / std over frames; thread = (slice of frames, coefficient)
    constexpr int SL = 12;                       // 12 * 20 = 240 active threads
    float* part = m.scratch;                     // [SL][20], the FFT scratch is free now
    float* feat = m.scratch + SL * N_MFCC;       // mean[20] ++ std[20]
    const int k = tid % N_MFCC, sl = tid / N_MFCC;
    if (sl < SL) {
        float s = 0.f;
        for (int t = sl; t < F; t += SL) s += mf[(size_t)t * N_MFCC + k];
        part[sl * N_MFCC + k] = s;
    }
    __syncthreads();
    if (tid < N_MFCC) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < SL; i++) s += part[i * N_MFCC + tid];
        feat[tid] = s / (float)F;
    }
    __syncthreads();
    if (sl < SL) {
        const float mu = feat[k];
        float s = 0.f;
        for (int t = sl; t < F; t += SL) { const float d = mf[(size_t)t * N_MFCC + k] - mu; s = fmaf(d, d, s); }
        part[sl * N_MFCC + k] = s;
    }
    __syncthreads();
    if (tid < N_MFCC) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < SL; i++) s += part[i * N_MFCC + tid];
        feat[N_MFCC + tid] = sqrtf(s / (float)F);
    }
    __syncthreads();
    return feat;
}

// K3, batch form: one CTA per caller-described segment (extract_mfcc / calculate_similarity / matches).
__global__ void __launch_bounds__(SEG_THREADS, 1)
segment_mfcc_match_kernel(const DeviceTables* __restrict__ T, const SegDesc* __restrict__ segs,
                          int cap_frames, float* __restrict__ ws,           // global spill [frames][129+20]
                          const TemplateFeat* __restrict__ tmpl, int n_tmpl, int tmpl_first,
                          float threshold,
                          float* __restrict__ feat_out,                      // [n_seg][40] or null
                          float* __restrict__ frames_out,                    // [frames][20] or null
                          float* __restrict__ scores,                        // [n_seg][n_tmpl] or null
                          unsigned char* __restrict__ matched)               // [n_seg][n_tmpl] or null
{
    extern __shared__ float smem[];
    const SegSmem m = seg_carve(smem, cap_frames);
    LaneConsts lc;
    seg_prologue(T, m, lc);
    const int tid = threadIdx.x;
    const SegDesc sd = segs[blockIdx.x];
    const float* feat = segment_features(sd, m, lc, cap_frames, ws, frames_out);
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

// Warp-level MFCC frame pipeline for sm_100a: one warp turns one 512-sample frame into 20 MFCCs.
// Restates (not ports) what the reference reaches through librosa.feature.mfcc
// (/root/reference/easywakeword/wakeword.py:561-563): periodic-Hann window, 512-point real FFT,
// |X|^2, Slaney mel filterbank (128 bands, sparse: 504 non-zeros), 10*log10(max(1e-10, .)),
// power_to_db floor, ortho DCT-II (first 20 coefficients).
//
// FFT: the 512 real samples are packed as 256 complex points z[n] = x[2n] + i x[2n+1]; the warp runs
// a 256-point complex FFT as radix 8 x 8 x 4 with 8 points per lane in registers and two
// conflict-free shared-memory transposes, then untangles Z[k], Z[256-k] (one shuffle pair per bin)
// into the 257 real-FFT bins.  Mel: each lane owns four bands (lane + 32 j).  DCT: the rows of the
// DCT-II are (anti)symmetric about the middle of the 128 bands, so the bands are folded first
// (x[b] +- x[127-b], two shuffles); the lower half-warp forms the even coefficients, the upper the odd
// ones (40 FMAs per lane, weights as ten 16-byte shared loads) and a halving shuffle butterfly
// (11 exchanges) leaves coefficient k in one lane.
// Window taps, twiddles, mel weights and the DCT matrix live in shared memory (one copy per CTA) so
// the kernel stays under ~80 registers and several CTAs fit on an SM.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ewk {

constexpr int N_FFT = 512;
constexpr int HOP = 160;
constexpr int N_BINS = 257;
constexpr int N_MELS = 128;
constexpr int N_MFCC = 20;
constexpr int MEL_NNZ_CAP = 512;          // 504 used for sr=16000, n_fft=512, 128 bands
// Lane l owns mel bands l + 32 j (j = 0..3).  A triangular band is a rising edge over the interval
// I_b = [f_b, f_b+1) of the mel grid and a falling edge over I_b+1, and neighbouring bands share their intervals
// (I_b+1 is the falling edge of band b AND the rising edge of band b + 1).  So the lane sums over ONE interval,
// I_(b+1), with two weights per bin — the falling weight of band b and the rising weight of band b + 1 — and receives
// the rising part of its own band from lane l - 1: every power-spectrum bin is read once instead of twice.  The
// widest interval of each group of 32 has 1 / 2 / 3 / 6 bins; shorter ones are padded with zero weights, which gives
// branch-free, fully unrolled loops of 12 taps per lane (checked on the host).
constexpr int MEL_TAPS0 = 1, MEL_TAPS1 = 2, MEL_TAPS2 = 3, MEL_TAPS3 = 6;
constexpr int MEL_TAPS = MEL_TAPS0 + MEL_TAPS1 + MEL_TAPS2 + MEL_TAPS3;
constexpr int SCR_PLANE = 320;            // floats per re / im transpose plane (8 rows x 40)
constexpr int SCR_P = 264;                // power spectrum, 257 bins padded; aliases the re plane (dead by then)
constexpr int SCR_WARP = 2 * SCR_PLANE;   // floats of scratch per warp (2560 B)
constexpr int DCT_LANE4 = 11;              // float4 per lane row of the shared DCT table (40 weights + 4 pad: conflict-free)
constexpr unsigned FULL = 0xffffffffu;
constexpr float TEN_LOG10_2 = 3.01029995663981195f;   // 10 * log10(2)

// Read-only tables, built on the host in double precision (ewk_tables.hpp) and copied once.
struct DeviceTables {
    float hann[N_FFT];            // scipy.signal.get_window('hann', 512, fftbins=True)
    float2 w256[256];             // exp(-2 pi i m / 256)
    float2 w512[256];             // exp(-2 pi i m / 512), m < 256
    int mel_start[N_MELS];        // first FFT bin with non-zero weight
    int mel_len[N_MELS];          // number of non-zero weights
    int mel_off[N_MELS];          // offset into mel_w
    float mel_w[MEL_NNZ_CAP];     // librosa.filters.mel(htk=False, norm='slaney'), row-compressed
    float dct_t[N_MELS * N_MFCC]; // ortho DCT-II, transposed: dct_t[b*20 + k]
    float2 mel_pad[MEL_TAPS * 32];// interval taps: mel_pad[tap * 32 + lane] = (falling weight of band b, rising weight of band
                                  // b + 1) for the tap's bin of interval b + 1, b = lane + 32 j; zero-padded, taps grouped by j
    int mel_first[4 * 32];        // mel_first[j * 32 + lane] = first FFT bin of that interval
};

// Per-CTA shared copy, laid out for conflict-free lane-indexed access.
struct FrameTables {
    float2 hw[8 * 32];            // hw[a*32 + lane]   = hann[2 lane + 64 a], hann[2 lane + 64 a + 1]
    float2 tw1[8 * 32];           // tw1[k*32 + lane]  = W256^(lane k)
    float2 tw3[4 * 32];           // tw3[m*32 + lane]  = W512^(lane + 32 m), m < 4 (bin pairs k, 256-k share it)
    float2 tw2[8 * 4];            // tw2[k*4 + c]      = W32^(c k)
    float2 melp[MEL_TAPS * 32];   // zero-padded interval taps (falling, rising), [tap][lane]
    int mfirst[4 * 32];           // first FFT bin of interval lane + 32 j + 1, [j][lane]
    int coef_of_lane[32];         // which MFCC coefficient the lane holds after warp_dct20 (-1: none)
    float4 dct[32 * DCT_LANE4];        // per lane: [band slot s < 4][m < 10] = dct(k = 2 m + (lane >> 4), band_s), see warp_dct20
    // front-end parameters of the context (ewk_config, ABI 2), set on the host after load_frame_tables
    float preemph;                // pre-emphasis coefficient (0: none = the reference, wakeword.py:561-563)
    int n_mfcc;                   // coefficients kept (1..20; the reference: 20)
    int pad_[2];
};

// Lays the per-CTA tables out from the flat ones.  Run ONCE, on the host, at context creation (tid 0 of 1): the result is
// appended to the DeviceTables allocation and every CTA copies that image (copy_frame_tables) instead of rebuilding it
// with scattered loads and integer divisions at the start of every launch.
__host__ __device__ inline void load_frame_tables(FrameTables& ft, const DeviceTables* __restrict__ T, int tid, int nthr) {
    for (int i = tid; i < 256; i += nthr) {
        const int a = i >> 5, lane = i & 31;
        ft.hw[i] = make_float2(T->hann[2 * lane + 64 * a], T->hann[2 * lane + 64 * a + 1]);
        ft.tw1[i] = T->w256[(lane * a) & 255];
        if (a < 4) ft.tw3[i] = T->w512[lane + 32 * a];
    }
    for (int i = tid; i < 32; i += nthr) ft.tw2[i] = T->w256[(8 * (i & 3) * (i >> 2)) & 255];
    for (int i = tid; i < MEL_TAPS * 32; i += nthr) ft.melp[i] = T->mel_pad[i];      // float2 (falling, rising)
    for (int i = tid; i < 4 * 32; i += nthr) ft.mfirst[i] = T->mel_first[i];
    float* dd = reinterpret_cast<float*>(ft.dct);
    for (int i = tid; i < 32 * DCT_LANE4 * 4; i += nthr) {
        const int lane = i / (DCT_LANE4 * 4), r = i - lane * (DCT_LANE4 * 4);
        const int sl = r / 10, m = r - sl * 10;                                // band slot, coefficient pair index
        const int owner = sl < 2 ? lane : (lane ^ 16);                          // slots 2, 3: the partner lane's bands
        const int b = owner + 32 * (sl & 1);                                    // folded band (< 64)
        dd[i] = r < 40 ? T->dct_t[b * N_MFCC + 2 * m + (lane >> 4)] : 0.f;
    }
    for (int i = tid; i < 32; i += nthr) {
        const int w = i & 7;
        const int within = w == 0 ? 0 : w == 1 ? 1 : w == 2 ? 2 : w == 4 ? 3 : w == 5 ? 4 : -1;
        ft.coef_of_lane[i] = within < 0 ? -1 : 2 * (5 * ((i >> 3) & 1) + within) + ((i >> 4) & 1);
    }
}

static_assert(sizeof(FrameTables) % 16 == 0, "the image is copied as 16-byte words");
constexpr size_t FRAME_IMAGE_OFFSET = (sizeof(DeviceTables) + 15) & ~(size_t)15;   // FrameTables image behind the flat tables

__device__ __forceinline__ void copy_frame_tables(FrameTables& ft, const DeviceTables* __restrict__ T, int tid, int nthr) {
    const int4* src = reinterpret_cast<const int4*>(reinterpret_cast<const char*>(T) + FRAME_IMAGE_OFFSET);
    int4* dst = reinterpret_cast<int4*>(&ft);
    for (int i = tid; i < (int)(sizeof(FrameTables) / 16); i += nthr) dst[i] = __ldg(src + i);
}

// packed fp32x2 arithmetic (sm_100: FADD2 / FMUL2 / FFMA2, one issue slot for both halves of a complex value)
__device__ __forceinline__ unsigned long long f2_as_u64(float2 a) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a.x), "f"(a.y));
    return r;
}
__device__ __forceinline__ float2 u64_as_f2(unsigned long long a) {
    float2 r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(a));
    return r;
}
__device__ __forceinline__ float2 operator+(float2 a, float2 b) {
    unsigned long long r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(f2_as_u64(a)), "l"(f2_as_u64(b)));
    return u64_as_f2(r);
}
__device__ __forceinline__ float2 operator-(float2 a, float2 b) {
    unsigned long long r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(f2_as_u64(a)), "l"(f2_as_u64(b)));
    return u64_as_f2(r);
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {            // element-wise (a.x b.x, a.y b.y): one FMUL2
    unsigned long long r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(f2_as_u64(a)), "l"(f2_as_u64(b)));
    return u64_as_f2(r);
}
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// forward 4-point DFT
__device__ __forceinline__ void dft4(float2 y0, float2 y1, float2 y2, float2 y3,
                                     float2& Y0, float2& Y1, float2& Y2, float2& Y3) {
    const float2 s0 = y0 + y2, s1 = y0 - y2, s2 = y1 + y3, d = y1 - y3;
    const float2 s3 = make_float2(d.y, -d.x);   // (y1 - y3) * (-i)
    Y0 = s0 + s2; Y2 = s0 - s2; Y1 = s1 + s3; Y3 = s1 - s3;
}

// forward 8-point DFT, natural order in and out
__device__ __forceinline__ void fft8(float2 (&v)[8]) {
    const float r = 0.70710678118654752440f;
    const float2 a0 = v[0] + v[4], a1 = v[1] + v[5], a2 = v[2] + v[6], a3 = v[3] + v[7];
    const float2 b0 = v[0] - v[4];
    float2 b1 = v[1] - v[5], b2 = v[2] - v[6], b3 = v[3] - v[7];
    b1 = make_float2((b1.x + b1.y) * r, (b1.y - b1.x) * r);     // * W8^1
    b2 = make_float2(b2.y, -b2.x);                               // * W8^2 = -i
    b3 = make_float2((b3.y - b3.x) * r, (-b3.x - b3.y) * r);    // * W8^3
    dft4(a0, a1, a2, a3, v[0], v[2], v[4], v[6]);
    dft4(b0, b1, b2, b3, v[1], v[3], v[5], v[7]);
}

// One warp: 512 windowed real samples -> power spectrum P[0..256] in shared memory.
// x[a] holds the PCM samples (2 lane + 64 a, +1) of the frame, already zero outside the signal.
// scr: SCR_WARP floats private to the warp.
__device__ __forceinline__ void warp_power_spectrum(const float2 (&x)[8], const FrameTables& ft, float* scr, int lane) {
    float* sre = scr;
    float* sim = scr + SCR_PLANE;
    float* P = scr;                           // written only after the last read of the transpose planes
    float2 v[8];
#pragma unroll
    for (int a = 0; a < 8; a++) {
        const float2 h = ft.hw[a * 32 + lane];
        v[a] = mul2(x[a], h);
    }
    // radix-8 over a  (n = 32a + lane)
    fft8(v);
#pragma unroll
    for (int k = 1; k < 8; k++) v[k] = cmul(v[k], ft.tw1[k * 32 + lane]);
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 8; k++) { sre[k * 36 + lane] = v[k].x; sim[k * 36 + lane] = v[k].y; }
    __syncwarp();
    // radix-8 over b  (lane = 4b + c -> thread (k_a, c))
    const int ka = lane >> 2, c = lane & 3;
#pragma unroll
    for (int b = 0; b < 8; b++) v[b] = make_float2(sre[ka * 36 + 4 * b + c], sim[ka * 36 + 4 * b + c]);
    fft8(v);
#pragma unroll
    for (int k = 1; k < 8; k++) v[k] = cmul(v[k], ft.tw2[k * 4 + c]);
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 8; k++) { sre[k * 40 + c * 8 + ka] = v[k].x; sim[k * 40 + c * 8 + ka] = v[k].y; }
    __syncwarp();
    // radix-4 over c  (thread (k_a, j) owns k_b in {j, j+4}); afterwards o[m] = Z[lane + 32 m]
    const int ka2 = lane & 7, j = lane >> 3;
    float2 o[8];
#pragma unroll
    for (int e = 0; e < 2; e++) {
        const int base = (j + 4 * e) * 40 + ka2;
        const float2 w0 = make_float2(sre[base], sim[base]);
        const float2 w1 = make_float2(sre[base + 8], sim[base + 8]);
        const float2 w2 = make_float2(sre[base + 16], sim[base + 16]);
        const float2 w3 = make_float2(sre[base + 24], sim[base + 24]);
        dft4(w0, w1, w2, w3, o[e], o[e + 2], o[e + 4], o[e + 6]);
    }
    __syncwarp();                             // every lane holds its Z values: the planes may be overwritten by P
    // real-FFT untangle, two bins per step: with E = (Z[k] + conj Z[256-k])/2, O = (Z[k] - conj Z[256-k])/(2i) and
    // T = W512^k O,  X[k] = E + T  and  X[256-k] = conj(E - T).  The lane owning k = lane + 32 m (m < 4) fetches
    // Z[256-k] from lane 32 - lane (register 7 - m) and writes both powers; lane 0 pairs with itself.
    const int pl = (32 - lane) & 31;
#pragma unroll
    for (int m = 0; m < 4; m++) {
        float pr = __shfl_sync(FULL, o[7 - m].x, pl);
        float pi = __shfl_sync(FULL, o[7 - m].y, pl);
        if (lane == 0) { pr = o[(8 - m) & 7].x; pi = o[(8 - m) & 7].y; }
        const float er = o[m].x + pr, ei = o[m].y - pi;     // 2E
        const float qr = o[m].y + pi, qi = pr - o[m].x;     // 2O
        const float2 w = ft.tw3[m * 32 + lane];
        const float tr = w.x * qr - w.y * qi, ti = w.x * qi + w.y * qr;
        const float ar = er + tr, ai = ei + ti, br = er - tr, bi = ei - ti;
        P[lane + 32 * m] = 0.25f * (ar * ar + ai * ai);
        P[(32 - lane) + 32 * (7 - m)] = 0.25f * (br * br + bi * bi);
    }
    if (lane == 0) P[128] = o[4].x * o[4].x + o[4].y * o[4].y;
    if (lane >= 1 && lane < SCR_P - 256) P[256 + lane] = 0.f;      // zero padding read by the padded mel taps
    __syncwarp();
}

// One warp: P[0..263] (bins 257..263 hold zeros) -> the lane's four log-mel values (bands lane + 32 j),
// 10*log10(max(1e-10, S)).  Branch-free: zero-padded taps, compile-time trip counts.
// falling part of band b (returned) and rising part of band b + 1 (rise) over the lane's interval
template <int NT>
__device__ __forceinline__ float mel_interval(const float* __restrict__ p, const float2* __restrict__ w, float& rise) {
    float fall = 0.f;
    rise = 0.f;
#pragma unroll
    for (int i = 0; i < NT; i++) {
        const float2 wt = w[i * 32];
        const float v = p[i];
        fall = fmaf(wt.x, v, fall);
        rise = fmaf(wt.y, v, rise);
    }
    return fall;
}

__device__ __forceinline__ void warp_log_mel(const float* __restrict__ P, const FrameTables& ft, int lane, float (&out)[4]) {
    const float2* w = ft.melp + lane;
    float fall[4], rise[4];
    fall[0] = mel_interval<MEL_TAPS0>(P + ft.mfirst[lane], w, rise[0]);
    fall[1] = mel_interval<MEL_TAPS1>(P + ft.mfirst[32 + lane], w + 32 * MEL_TAPS0, rise[1]);
    fall[2] = mel_interval<MEL_TAPS2>(P + ft.mfirst[64 + lane], w + 32 * (MEL_TAPS0 + MEL_TAPS1), rise[2]);
    fall[3] = mel_interval<MEL_TAPS3>(P + ft.mfirst[96 + lane], w + 32 * (MEL_TAPS0 + MEL_TAPS1 + MEL_TAPS2), rise[3]);
    // band b = rising part over I_b (held by the lane that owns interval b: lane - 1, or lane 31 of the previous
    // group; band 0 has none: no bin lies inside I_0) + falling part over I_(b+1) (own)
#pragma unroll
    for (int j = 0; j < 4; j++) {
        float r = __shfl_up_sync(FULL, rise[j], 1);
        const float rw = j > 0 ? __shfl_sync(FULL, rise[j > 0 ? j - 1 : 0], 31) : 0.f;
        if (lane == 0) r = rw;
        out[j] = TEN_LOG10_2 * __log2f(fmaxf(r + fall[j], 1e-10f));
    }
}

// One warp: log-mel x[4] per lane (bands lane + 32 j, already floored) -> ortho DCT-II.  Lanes 0-15 form
// coefficients 0-9, lanes 16-31 coefficients 10-19, each over its own four bands and the four of lane ^ 16
// (80 FMAs, 24 16-byte shared loads); a halving shuffle butterfly over the 16 lanes of a half
// (10 -> 5 -> 3 -> 2 -> 1 values, 11 exchanges) leaves coefficient ft.coef_of_lane[lane] in the lane.
__device__ __forceinline__ float warp_dct20(const float (&x)[4], const FrameTables& ft, int lane) {
    // fold: c(k, 127 - b) = (-1)^k c(k, b), so C[k] = sum_{b<64} c(k, b) (x[b] + (-1)^k x[127 - b]).  Band 127 - b of
    // the lane's bands b = lane, lane + 32 sits in lane 31 - lane (registers 3, 2).
    const float y3 = __shfl_sync(FULL, x[3], 31 - lane), y2 = __shfl_sync(FULL, x[2], 31 - lane);
    const bool half = lane & 16;                              // lower half-warp: even k, upper: odd k
    const float s0 = x[0] + y3, d0 = x[0] - y3, s1 = x[1] + y2, d1 = x[1] - y2;
    float v[4];
    v[0] = half ? d0 : s0;
    v[1] = half ? d1 : s1;
    v[2] = __shfl_xor_sync(FULL, half ? s0 : d0, 16);         // the partner's folded values of MY parity
    v[3] = __shfl_xor_sync(FULL, half ? s1 : d1, 16);
    float acc[10];
#pragma unroll
    for (int m = 0; m < 10; m++) acc[m] = 0.f;
    const float4* row = ft.dct + lane * DCT_LANE4;
#pragma unroll
    for (int q = 0; q < 10; q++) {
        const float4 w = row[q];
        acc[(4 * q + 0) % 10] = fmaf(w.x, v[(4 * q + 0) / 10], acc[(4 * q + 0) % 10]);
        acc[(4 * q + 1) % 10] = fmaf(w.y, v[(4 * q + 1) / 10], acc[(4 * q + 1) % 10]);
        acc[(4 * q + 2) % 10] = fmaf(w.z, v[(4 * q + 2) / 10], acc[(4 * q + 2) % 10]);
        acc[(4 * q + 3) % 10] = fmaf(w.w, v[(4 * q + 3) / 10], acc[(4 * q + 3) % 10]);
    }
    // halving butterfly over lane bits 3..0: 10 -> 5 -> (6) 3 -> (4) 2 -> 1 values per lane
    float a5[6], a3[4], a2[2];
    const bool u8 = lane & 8, u4 = lane & 4, u2 = lane & 2, u1 = lane & 1;
#pragma unroll
    for (int i = 0; i < 5; i++) {
        const float send = u8 ? acc[i] : acc[i + 5], keep = u8 ? acc[i + 5] : acc[i];
        a5[i] = keep + __shfl_xor_sync(FULL, send, 8);
    }
    a5[5] = 0.f;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const float send = u4 ? a5[i] : a5[i + 3], keep = u4 ? a5[i + 3] : a5[i];
        a3[i] = keep + __shfl_xor_sync(FULL, send, 4);
    }
    a3[3] = 0.f;
#pragma unroll
    for (int i = 0; i < 2; i++) {
        const float send = u2 ? a3[i] : a3[i + 2], keep = u2 ? a3[i + 2] : a3[i];
        a2[i] = keep + __shfl_xor_sync(FULL, send, 2);
    }
    const float send = u1 ? a2[0] : a2[1], keep = u1 ? a2[1] : a2[0];
    return keep + __shfl_xor_sync(FULL, send, 1);
}

// One warp, one frame: PCM -> 20 MFCCs written to out[0..19]; the frame's log-mel min / max (before
// flooring) are returned in every lane.  floor_db = -INFINITY disables the power_to_db floor.  When lm_out is
// given, the frame's 128 un-floored log-mel values are stored there (band b at lm_out[b]) so that a later floor
// costs a DCT (warp_refloor_mfcc) instead of the whole pipeline.
// Deliberately NOT inlined: every kernel shares one copy of the ~1.5k-instruction pipeline, which keeps the
// kernels inside the instruction cache.
__device__ __noinline__ float2 warp_frame_mfcc(float2 x0, float2 x1, float2 x2, float2 x3, float2 x4, float2 x5, float2 x6,
                                               float2 x7, const FrameTables* __restrict__ ftp, float* scr,
                                               float floor_db, float* __restrict__ out, float* __restrict__ lm_out) {
    const int lane = threadIdx.x & 31;
    const FrameTables& ft = *ftp;
    const float2 x[8] = {x0, x1, x2, x3, x4, x5, x6, x7};
    warp_power_spectrum(x, ft, scr, lane);
    float v[4];
    warp_log_mel(scr, ft, lane, v);
    float mn = fminf(fminf(v[0], v[1]), fminf(v[2], v[3]));
    float mx = fmaxf(fmaxf(v[0], v[1]), fmaxf(v[2], v[3]));
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        mn = fminf(mn, __shfl_xor_sync(FULL, mn, o));
        mx = fmaxf(mx, __shfl_xor_sync(FULL, mx, o));
    }
    if (lm_out) {
#pragma unroll
        for (int j = 0; j < 4; j++) __stcg(lm_out + lane + 32 * j, v[j]);
    }
#pragma unroll
    for (int j = 0; j < 4; j++) v[j] = fmaxf(v[j], floor_db);
    const float cft = warp_dct20(v, ft, lane);
    const int k = ft.coef_of_lane[lane];
    if (k >= 0) out[k] = cft;
    return make_float2(mn, mx);
}

__device__ __forceinline__ void warp_frame_mfcc(const float2 (&x)[8], const FrameTables& ft, float* scr, int lane,
                                                float floor_db, float* __restrict__ out, float& fmin_o, float& fmax_o,
                                                float* __restrict__ lm_out = nullptr) {
    const float2 r = warp_frame_mfcc(x[0], x[1], x[2], x[3], x[4], x[5], x[6], x[7], &ft, scr, floor_db, out, lm_out);
    fmin_o = r.x;
    fmax_o = r.y;
}

// One warp: a frame's stored log-mel values (warp_frame_mfcc's lm_out) -> its 20 MFCCs under floor_db.  Same
// arithmetic as the tail of warp_frame_mfcc, so the result is bit-identical to recomputing the frame with the floor.
__device__ __noinline__ void warp_refloor_mfcc(const float* __restrict__ lm, const FrameTables* __restrict__ ftp,
                                               float floor_db, float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const FrameTables& ft = *ftp;
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; j++) v[j] = fmaxf(__ldcg(lm + lane + 32 * j), floor_db);
    const float cft = warp_dct20(v, ft, lane);
    const int k = ft.coef_of_lane[lane];
    if (k >= 0) out[k] = cft;
}

}  // namespace ewk

// Warp-level MFCC frame pipeline for sm_100a: one warp turns one 512-sample frame into 128 log-mel
// values.  Restates (not ports) what the reference reaches through librosa.feature.mfcc
// (/root/reference/easywakeword/wakeword.py:561-563): periodic-Hann window, 512-point real FFT,
// |X|^2, Slaney mel filterbank (128 bands, sparse: 504 non-zeros), 10*log10(max(1e-10, .)).
//
// FFT: the 512 real samples are packed as 256 complex points z[n] = x[2n] + i x[2n+1]; the warp runs
// a 256-point complex FFT as radix 8 x 8 x 4 with 8 points per lane in registers and two
// conflict-free shared-memory transposes, then untangles Z[k], Z[256-k] (one shuffle pair per bin)
// into the 257 real-FFT bins.  All twiddles and the lane's 16 window taps live in registers for the
// whole kernel (they depend only on the lane), so the per-frame loop loads nothing but PCM.
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
constexpr int SCR_PLANE = 320;            // floats per re / im transpose plane (8 rows x 40)
constexpr int SCR_P = 264;                // power spectrum, 257 bins padded
constexpr int SCR_WARP = 2 * SCR_PLANE + SCR_P;   // floats of scratch per warp (3616 B)
constexpr unsigned FULL = 0xffffffffu;

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
};

struct LaneConsts {
    float2 hw[8];    // window taps for samples (2*lane + 64a, +1)
    float2 tw1[8];   // W256^(lane * k)
    float2 tw2[8];   // W32^((lane & 3) * k)
    float2 tw3[8];   // W512^(lane + 32 m)
    int mstart[4], mlen[4], moff[4];   // the lane's four mel bands: lane + 32 j
};

__device__ __forceinline__ void init_lane_consts(LaneConsts& lc, const DeviceTables* __restrict__ T, int lane) {
#pragma unroll
    for (int a = 0; a < 8; a++) lc.hw[a] = make_float2(T->hann[2 * lane + 64 * a], T->hann[2 * lane + 64 * a + 1]);
#pragma unroll
    for (int k = 0; k < 8; k++) lc.tw1[k] = T->w256[(lane * k) & 255];
    const int c = lane & 3;
#pragma unroll
    for (int k = 0; k < 8; k++) lc.tw2[k] = T->w256[(8 * c * k) & 255];
#pragma unroll
    for (int m = 0; m < 8; m++) lc.tw3[m] = T->w512[lane + 32 * m];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        lc.mstart[j] = T->mel_start[lane + 32 * j];
        lc.mlen[j] = T->mel_len[lane + 32 * j];
        lc.moff[j] = T->mel_off[lane + 32 * j];
    }
}

__device__ __forceinline__ float2 operator+(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 operator-(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
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
// ld(i) returns the PCM samples (i, i+1) of the frame, i even in [0, 512), already zero outside
// the signal.  scr: SCR_WARP floats private to the warp.
template <class Ld>
__device__ __forceinline__ void warp_power_spectrum(Ld&& ld, const LaneConsts& lc, float* scr, int lane) {
    float* sre = scr;
    float* sim = scr + SCR_PLANE;
    float* P = scr + 2 * SCR_PLANE;
    float2 v[8];
#pragma unroll
    for (int a = 0; a < 8; a++) {
        const float2 s = ld(2 * lane + 64 * a);
        v[a] = make_float2(s.x * lc.hw[a].x, s.y * lc.hw[a].y);
    }
    // radix-8 over a  (n = 32a + lane)
    fft8(v);
#pragma unroll
    for (int k = 1; k < 8; k++) v[k] = cmul(v[k], lc.tw1[k]);
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
    for (int k = 1; k < 8; k++) v[k] = cmul(v[k], lc.tw2[k]);
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
    // real-FFT untangle: X[k] = E[k] + W512^k O[k],  E = (Z[k] + conj Z[256-k])/2,  O = (Z[k] - conj Z[256-k])/(2i)
    const int pl = (32 - lane) & 31;
#pragma unroll
    for (int m = 0; m < 8; m++) {
        float pr = __shfl_sync(FULL, o[7 - m].x, pl);
        float pi = __shfl_sync(FULL, o[7 - m].y, pl);
        if (lane == 0) { pr = o[(8 - m) & 7].x; pi = o[(8 - m) & 7].y; }
        const float er = o[m].x + pr, ei = o[m].y - pi;     // 2E
        const float qr = o[m].y + pi, qi = pr - o[m].x;     // 2O
        const float2 w = lc.tw3[m];
        const float xr = er + (w.x * qr - w.y * qi);
        const float xi = ei + (w.x * qi + w.y * qr);
        P[lane + 32 * m] = 0.25f * (xr * xr + xi * xi);
    }
    if (lane == 0) { const float d = o[0].x - o[0].y; P[256] = d * d; }
    __syncwarp();
}

// One warp: P[0..256] -> the lane's four log-mel values (bands lane + 32 j), 10*log10(max(1e-10, S)).
__device__ __forceinline__ void warp_log_mel(const float* __restrict__ P, const float* __restrict__ melw,
                                             const LaneConsts& lc, float (&out)[4]) {
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const float* p = P + lc.mstart[j];
        const float* w = melw + lc.moff[j];
        float acc = 0.f;
        const int n = lc.mlen[j];
        for (int i = 0; i < n; i++) acc = fmaf(w[i], p[i], acc);
        out[j] = 10.0f * log10f(fmaxf(acc, 1e-10f));
    }
}

}  // namespace ewk

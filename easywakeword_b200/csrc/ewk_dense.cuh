// K4 — dense_score: the level-2 score of WordMatcher.calculate_similarity
// (/root/reference/easywakeword/wakeword.py:591-625) for a template-length window at EVERY 10 ms hop
// of every stream (usage shape of examples/tune_threshold.py:86-116 at hop
// granularity).  One CTA per stream walks the requested hops in sub-chunks of DH hops and never writes
// an intermediate to global memory.
//
// Window at hop h for template k (L samples, n = ceil(L/160), F = 1 + L/160 frames):
//     x[160 (h - n) : 160 (h - n) + L]      — the latest template-length window that starts on the
// hop grid and is complete at hop h (oracle.ewk_oracle.dense_window).  Handing that window to librosa
// means window-local zero-pad centring and a window-local top_db floor, so per window:
//   * frames t = 2 .. t_hi lie fully inside the window: they are frames of the STREAM grid
//     (centre 160 g, g = h - n + t), shared by all windows and templates -> MFCC ring G[g];
//   * frames t = 0, 1 (left edge) and t_hi+1 .. F-1 (right edge) see zeros outside the window and are
//     computed per (hop, template) -> edge rows;
//   * floor = (max log-mel over the window's frames) - 80.  Frames are first computed un-floored
//     together with their log-mel min / max; a window whose min is below its floor recomputes just
//     the affected frames with the floor ("patches"), everything else is reused.
//   * mean / std over the F frames, cosine vs the template, p^1.5/10.  Consecutive windows share all but one of
//     their stream-grid frames, so the statistics are pooled from blocks: (mean, M2) of every aligned block of 8
//     stream-grid rows is computed once per sub-chunk, and a window combines, in a fixed order, its left edge
//     frames, the loose rows before the first aligned block, the blocks, the loose rows after the last one and
//     its right edge frames with the pairwise update of Chan et al. (n, mean, M2) — about 25 items instead of
//     100 rows, no cancellation, and bit-identical for every way of cutting the request into calls or sub-chunks.
#pragma once
#include <climits>

#include "ewk_segment.cuh"
#include "ewk_streams.cuh"

namespace ewk {

constexpr int DENSE_THREADS = 512;
constexpr int DENSE_WARPS = DENSE_THREADS / 32;
constexpr int DH = 32;                 // hops per sub-chunk
constexpr int DENSE_MAX_T = 4;         // templates per launch
constexpr int DENSE_MAX_F = 224;       // frames per window (templates up to ~2.2 s)
constexpr int DENSE_MIN_L = 640;       // shorter templates would make a frame both left- and right-masked
constexpr int ROW = N_MFCC + 2;        // mfcc[20], log-mel min, log-mel max
constexpr int DENSE_KEEP = 224;        // stream-grid rows carried from one call to the next (>= frames of the longest window)
constexpr int PATCH_CAP = 5;           // per-warp rows recomputed with a floor: 4 edge frames + 1 stream-grid frame
constexpr int BLK = 8;                 // stream-grid rows per statistics block (aligned at absolute multiples of BLK)

struct DenseTmplDev {
    int L, n, F, t_hi, r, slot;        // r = F - 1 - t_hi right-edge frames
};

struct DenseArgs {
    long long hop0;                    // first hop scored (hop h <-> 160 h samples of the stream)
    int n_hops;
    int T;
    int DG;                            // rows of the grid-frame ring
    DenseTmplDev t[DENSE_MAX_T];
    float* out;                        // [n_streams][n_hops][T]
    // carry-over between consecutive calls: the newest stream-grid rows of every stream (functions of the PCM only)
    float* keep_rows;                  // [n_streams][DENSE_KEEP][ROW], row of grid frame g at g % DENSE_KEEP
    long long* keep_end;               // [n_streams][2]: grid frames [keep_end[2s], keep_end[2s+1]) are stored (0, 0: nothing)
};

__host__ __device__ inline int dense_nblk(int DG) { return DG / BLK + 2; }     // ring of block statistics

__host__ __device__ inline size_t dense_smem_bytes(int DG, int T) {
    return sizeof(FrameTables) +
           sizeof(float) * ((size_t)DENSE_WARPS * SCR_WARP + (size_t)DG * ROW + (size_t)T * DH * 4 * ROW +
                            (size_t)DENSE_WARPS * PATCH_CAP * N_MFCC + (size_t)DG * (N_MFCC + 1) + 8 +
                            (size_t)2 * dense_nblk(DG) * 2 * N_MFCC);
}

// frame `t` of the window of template `tp` starting at grid index j: pointer to its ROW
// (jb = j mod DG, so that the ring index is one add and one conditional subtract)
__device__ __forceinline__ const float* dense_row(const float* G, const float* edge_kh, const DenseTmplDev& tp, int DG,
                                                  int jb, int t) {
    if (t < 2) return edge_kh + t * ROW;
    if (t <= tp.t_hi) { int r = jb + t; if (r >= DG) r -= DG; return G + r * ROW; }
    return edge_kh + (2 + t - tp.t_hi - 1) * ROW;
}

// (mean, M2 = sum of squared deviations) of BLK values, two passes, fixed order
__device__ __forceinline__ void block_mean_m2(const float (&v)[BLK], float& mu, float& m2) {
    float sum = 0.f;
#pragma unroll
    for (int q = 0; q < BLK; q++) sum += v[q];
    mu = sum * (1.0f / BLK);
    m2 = 0.f;
#pragma unroll
    for (int q = 0; q < BLK; q++) { const float d = v[q] - mu; m2 = fmaf(d, d, m2); }
}

// pooled update (Chan, Golub, LeVeque): fold an item (nb values, mean mb, M2 m2b) into the running (n, mean, M2)
__device__ __forceinline__ void pool_item(float& n, float& mean, float& M2, float nb, float mb, float m2b) {
    const float nn = n + nb, f = __fdividef(nb, nn), delta = mb - mean;
    mean = fmaf(delta, f, mean);
    M2 += fmaf(delta * delta, n * f, m2b);
    n = nn;
}

template <bool PRE>
__global__ void __launch_bounds__(DENSE_THREADS, 2)
dense_score_kernel(const DeviceTables* __restrict__ T, BankView B, const TemplateFeat* __restrict__ tmpl, DenseArgs A) {
    extern __shared__ __align__(16) float smem[];
    FrameTables* ft = reinterpret_cast<FrameTables*>(smem);
    float* scratch = smem + sizeof(FrameTables) / sizeof(float);
    float* G = scratch + DENSE_WARPS * SCR_WARP;                 // [DG][ROW]
    float* edge = G + (size_t)A.DG * ROW;                        // [T][DH][4][ROW]
    float* wbuf = edge + (size_t)A.T * DH * 4 * ROW;             // per warp: patch[PATCH_CAP][20]
    // alternate ring: rows of G recomputed with ONE floor value (the current one of the stream), tagged per row, so
    // that a floored stream-grid frame is recomputed once per floor value and not once per window that contains it
    float* G2 = wbuf + (size_t)DENSE_WARPS * PATCH_CAP * N_MFCC;                        // [DG][20]
    int* g2tag = reinterpret_cast<int*>(G2 + (size_t)A.DG * N_MFCC);                    // [DG] floor bits of the row
    int* fstar_s = g2tag + A.DG;                                                        // [1] floor bits served by G2
    // statistics of aligned blocks of BLK stream-grid rows, (mean[20], M2[20]) each, block g / BLK at (g / BLK) % NBLK:
    // BS[0] over the un-floored rows G, BS[1] over the rows as floored by the stream's current floor (G2 where it applies)
    const int NBLK = dense_nblk(A.DG);
    float* BS = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(fstar_s + 4) + 15) & ~(uintptr_t)15);   // [2][NBLK][40]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int s = blockIdx.x;
    copy_frame_tables(*ft, T, tid, DENSE_THREADS);
    float* scr = scratch + warp * SCR_WARP;
    for (int i = tid; i < A.DG; i += DENSE_THREADS) g2tag[i] = 0x7fc00001;     // matches no floor
    __syncthreads();

    float* patch = wbuf + (size_t)warp * PATCH_CAP * N_MFCC;

    const size_t esz = B.fmt == 1 ? 2 : 4;
    PcmReader rd;
    rd.f = B.fmt == 0 ? (const float*)((const char*)B.ring + (size_t)s * B.P * esz) : nullptr;
    rd.q = B.fmt == 1 ? (const short*)((const char*)B.ring + (size_t)s * B.P * esz) : nullptr;
    rd.ring = B.P;
    __syncthreads();
    rd.pre = PRE ? ft->preemph : 0.f;
    // stream-grid frames are read through a view that starts `back` samples early, so that pre-emphasis finds x[n-1]
    // of the frame's first sample inside the view (no window starts there); without pre-emphasis the view is the frame
    constexpr int back = PRE ? 2 : 0;

    int max_n = 0, min_n = INT_MAX;
    for (int k = 0; k < A.T; k++) { max_n = max(max_n, A.t[k].n); min_n = min(min_n, A.t[k].n); }

    long long g_done = LLONG_MIN, g_valid_lo = LLONG_MAX;       // ring G holds rows [max(g_valid_lo, g_done - DG), g_done)
    {
        // rows computed by the previous call are reused when this call continues where it stopped: the history of
        // the first window (up to t_hi frames) is loaded instead of recomputed
        long long need_lo = LLONG_MAX;
        for (int k = 0; k < A.T; k++) need_lo = min(need_lo, A.hop0 - A.t[k].n + 2);
        if (need_lo < 2) need_lo = 2;
        const long long klo = A.keep_rows ? A.keep_end[2 * s] : 0, kend = A.keep_rows ? A.keep_end[2 * s + 1] : 0;
        if (klo <= need_lo && need_lo < kend && kend - need_lo <= A.DG) {
            const float* kr = A.keep_rows + (size_t)s * DENSE_KEEP * ROW;
            const int cnt = (int)(kend - need_lo);
            for (int i = tid; i < cnt * ROW; i += DENSE_THREADS) {
                const long long g = need_lo + i / ROW;
                G[(size_t)(g % A.DG) * ROW + i % ROW] = kr[(size_t)(g % DENSE_KEEP) * ROW + i % ROW];
            }
            g_done = kend;
            g_valid_lo = need_lo;
        }
        __syncthreads();
    }
    for (long long hs = A.hop0; hs < A.hop0 + A.n_hops; hs += DH) {
        const int nh = (int)min((long long)DH, A.hop0 + A.n_hops - hs);
        // ---- frames of this sub-chunk: new stream-grid frames, then the edge frames of every (hop, template)
        long long g_lo = LLONG_MAX, g_hi = LLONG_MIN;
        for (int k = 0; k < A.T; k++) {
            g_lo = min(g_lo, hs - A.t[k].n + 2);
            g_hi = max(g_hi, hs + nh - 1 - A.t[k].n + A.t[k].t_hi);
        }
        if (g_lo < 2) g_lo = 2;                                   // grid frame g needs samples from 160 g - 256 >= 0
        const long long g_from = max(g_lo, g_done);
        if (g_valid_lo == LLONG_MAX) g_valid_lo = g_from;
        const int n_g = (int)max(0LL, g_hi - g_from + 1);
        int n_e = 0;
        for (int k = 0; k < A.T; k++) n_e += DH * (2 + A.t[k].r);        // job = (template, edge e, hop slot): slots >= nh are skipped
        // 32-bit bases for this sub-chunk: ring position of sample 160*hs and ring row of grid frame g_from
        const int hs_pos = (int)((160 * hs) % B.P);
        const int gfrom_row = n_g > 0 ? (int)(g_from % A.DG) : 0;
        const int gfrom_rel = (int)(g_from - hs);                  // grid index relative to hs
        for (int job = warp; job < n_g + n_e; job += DENSE_WARPS) {
            float* row;
            int f0;
            if (job < n_g) {
                // absolute first sample 160 g - 256 (unmasked frame of the stream grid)
                int pos = hs_pos + 160 * (gfrom_rel + job) - N_FFT / 2 - back;       // |offset| < P: one wrap
                if (pos < 0) pos += B.P; else if (pos >= B.P) pos -= B.P;
                rd.start = pos; rd.len = N_FFT + back;
                f0 = back;
                int r = gfrom_row + job; if (r >= A.DG) r -= A.DG;
                row = G + r * ROW;
                if (lane == 0) g2tag[r] = 0x7fc00001;
            } else {
                int rem = job - n_g, k = 0;
                while (rem >= DH * (2 + A.t[k].r)) { rem -= DH * (2 + A.t[k].r); k++; }
                const DenseTmplDev& tp = A.t[k];
                const int e = rem / DH, hl = rem % DH;              // DH is a power of two
                if (hl >= nh || hs + hl - tp.n < 0) continue;       // unused slot / window starts before the stream
                const int t = e < 2 ? e : tp.t_hi + 1 + (e - 2);
                int pos = hs_pos + 160 * (hl - tp.n);
                if (pos < 0) pos += B.P; else if (pos >= B.P) pos -= B.P;
                rd.start = pos; rd.len = tp.L;                      // zeros outside the window
                f0 = t * HOP - N_FFT / 2;
                row = edge + ((k * DH + hl) * 4 + e) * ROW;
            }
            float2 x[8];
            load_frame_pairs_at<PRE>(rd, f0, lane, x);
            float mn, mx;
            warp_frame_mfcc(x, *ft, scr, lane, -INFINITY, row, mn, mx);
            if (lane == 0) { row[N_MFCC] = mn; row[N_MFCC + 1] = mx; }
        }
        g_done = g_hi + 1;
        __syncthreads();

        // ---- the floor of the newest window of template 0 is the stream's current floor f*: stream-grid frames it
        // changes are recomputed once into G2 (only rows not yet tagged with f*: in steady state the new rows)
        {
            const DenseTmplDev& tp0 = A.t[0];
            const long long jl0 = hs + nh - 1 - tp0.n;
            if (warp == 0) {
                float wmax = -INFINITY;
                if (jl0 >= 0) {
                    const int j0 = (int)(jl0 % A.DG);
                    const float* ek0 = edge + (nh - 1) * 4 * ROW;
                    for (int t = lane; t < tp0.F; t += 32) wmax = fmaxf(wmax, dense_row(G, ek0, tp0, A.DG, j0, t)[N_MFCC + 1]);
                }
#pragma unroll
                for (int o = 16; o; o >>= 1) wmax = fmaxf(wmax, __shfl_xor_sync(FULL, wmax, o));
                if (lane == 0) *fstar_s = jl0 >= 0 ? __float_as_int(wmax - 80.0f) : 0x7fc00002;
            }
            __syncthreads();
            const int fbits = *fstar_s;
            const float fstar = __int_as_float(fbits);
            const int span = (int)(g_hi - g_lo + 1);
            const int row_lo = span > 0 ? (int)(g_lo % A.DG) : 0;
            const int glo_rel = (int)(g_lo - hs);
            if (fbits != 0x7fc00002)
                for (int i = warp; i < span; i += DENSE_WARPS) {
                    int r = row_lo + i; if (r >= A.DG) r -= A.DG;
                    if (!(G[r * ROW + N_MFCC] < fstar) || g2tag[r] == fbits) continue;      // warp-uniform
                    int pos = hs_pos + 160 * (glo_rel + i) - N_FFT / 2 - back;
                    if (pos < 0) pos += B.P; else if (pos >= B.P) pos -= B.P;
                    rd.start = pos; rd.len = N_FFT + back;
                    float2 x[8];
                    load_frame_pairs_at<PRE>(rd, back, lane, x);
                    float mn, mx;
                    warp_frame_mfcc(x, *ft, scr, lane, fstar, G2 + r * N_MFCC, mn, mx);
                    if (lane == 0) g2tag[r] = fbits;
                }
            __syncthreads();
        }

        // ---- statistics of the span's aligned blocks (both variants), one warp per block
        // blocks [b_lo, b_lo + n_blk) lie inside the span; 32-bit ring positions of the first one
        const long long b_lo = (g_lo + BLK - 1) / BLK;
        const int n_blk = (int)max(0LL, (g_hi + 1) / BLK - b_lo);
        const int blk_row_lo = (int)((b_lo * BLK) % A.DG), blk_idx_lo = (int)(b_lo % NBLK);
        {
            const int fb = *fstar_s;
            const float fstar = __int_as_float(fb);
            for (int i = warp; i < 2 * n_blk; i += DENSE_WARPS) {
                const int var = i & 1;
                int r = blk_row_lo + (i >> 1) * BLK; if (r >= A.DG) r -= A.DG;       // n_blk * BLK <= DG
                int bi = blk_idx_lo + (i >> 1); if (bi >= NBLK) bi -= NBLK;
                float v[BLK];
#pragma unroll
                for (int q = 0; q < BLK; q++) {
                    const float* src = G + r * ROW;
                    if (var && src[N_MFCC] < fstar && g2tag[r] == fb) src = G2 + r * N_MFCC;
                    v[q] = lane < N_MFCC ? src[lane] : 0.f;
                    r++; if (r >= A.DG) r -= A.DG;
                }
                float mu, m2;
                block_mean_m2(v, mu, m2);
                float* dst = BS + ((size_t)var * NBLK + (size_t)bi) * 2 * N_MFCC;
                if (lane < N_MFCC) { dst[lane] = mu; dst[N_MFCC + lane] = m2; }
            }
        }
        __syncthreads();

        // ---- windows of this sub-chunk, one warp per (template, hop): window max -> floor; class of the window
        //   0: no stream-grid frame of the window is floored            -> block statistics over G
        //   1: floored, and the floor is the stream's current one (f*)  -> block statistics over G2 | G
        //   2: floored with another floor: blocks are formed here, floored rows recomputed on the fly
        // All three give the same bits for the same window: same items, same order, same arithmetic.
        for (int w = warp; w < A.T * nh; w += DENSE_WARPS) {
            const int k = w / nh, hl = w - k * nh;
            const DenseTmplDev tp = A.t[k];
            const long long jl = hs + hl - tp.n;
            float* outp = A.out + ((size_t)s * A.n_hops + (size_t)(hs - A.hop0 + hl)) * A.T + k;
            const TemplateFeat& tf = tmpl[tp.slot];
            if (jl < 0 || !tf.valid) { if (lane == 0) *outp = __int_as_float(0x7fc00000); continue; }
            const int j = (int)(jl % A.DG);                            // ring row base of this window
            const float* ekh = edge + (k * DH + hl) * 4 * ROW;
            float wmax = -INFINITY, rmin = INFINITY;                   // max over all frames, min over the ring frames
            for (int t = 2 + lane; t <= tp.t_hi; t += 32) {
                int r = j + t; if (r >= A.DG) r -= A.DG;
                rmin = fminf(rmin, G[r * ROW + N_MFCC]);
                wmax = fmaxf(wmax, G[r * ROW + N_MFCC + 1]);
            }
            const int n_edge = 2 + tp.r;
            const float emin = lane < n_edge ? ekh[lane * ROW + N_MFCC] : INFINITY;
            if (lane < n_edge) wmax = fmaxf(wmax, ekh[lane * ROW + N_MFCC + 1]);
#pragma unroll
            for (int o = 16; o; o >>= 1) {
                wmax = fmaxf(wmax, __shfl_xor_sync(FULL, wmax, o));
                rmin = fminf(rmin, __shfl_xor_sync(FULL, rmin, o));
            }
            const float floor_db = wmax - 80.0f;                       // librosa.power_to_db(top_db=80) on this window
            const unsigned emask = __ballot_sync(FULL, emin < floor_db);   // floored edge frames
            const int cls = !(rmin < floor_db) ? 0 : (__float_as_int(floor_db) == *fstar_s ? 1 : 2);
            { int pos = hs_pos + 160 * (hl - tp.n); if (pos < 0) pos += B.P; else if (pos >= B.P) pos -= B.P; rd.start = pos; }
            rd.len = tp.L;
            // floored edge frames are recomputed with the floor (window-local PCM view) into patch slots 0..3
            for (unsigned mm = emask; mm; mm &= mm - 1) {
                const int e = __ffs(mm) - 1;
                const int t = e < 2 ? e : tp.t_hi + 1 + (e - 2);
                float2 x[8];
                load_frame_pairs_at<PRE>(rd, t * HOP - N_FFT / 2, lane, x);
                float mn, mx;
                warp_frame_mfcc(x, *ft, scr, lane, floor_db, patch + e * N_MFCC, mn, mx);
            }
            __syncwarp();
            // value of ring frame t (row r) under this window's floor, coefficient = lane
            auto ring_value = [&](int t, int r) -> float {
                if (cls == 0 || !(G[r * ROW + N_MFCC] < floor_db)) return lane < N_MFCC ? G[r * ROW + lane] : 0.f;
                if (cls == 1) return lane < N_MFCC ? G2[r * N_MFCC + lane] : 0.f;
                float2 x[8];
                load_frame_pairs_at<PRE>(rd, t * HOP - N_FFT / 2, lane, x);
                float mn, mx;
                __syncwarp();
                warp_frame_mfcc(x, *ft, scr, lane, floor_db, patch + 4 * N_MFCC, mn, mx);
                __syncwarp();
                return lane < N_MFCC ? patch[4 * N_MFCC + lane] : 0.f;
            };
            auto edge_value = [&](int e) -> float {
                if (lane >= N_MFCC) return 0.f;
                return ((emask >> e) & 1u) ? patch[e * N_MFCC + lane] : ekh[e * ROW + lane];
            };
            // pooled statistics, fixed order: left edges, loose rows, aligned blocks, loose rows, right edges.
            // Ring frames t = 2 .. t_hi are grid frames jl + t; t0 = first t on a block boundary.
            float n = 1.f, mean = edge_value(0), M2 = 0.f;
            pool_item(n, mean, M2, 1.f, edge_value(1), 0.f);
            const int t0 = min(tp.t_hi + 1, 2 + ((BLK - (((int)(jl & (BLK - 1)) + 2) & (BLK - 1))) & (BLK - 1)));
            int t = 2, r = j + 2; if (r >= A.DG) r -= A.DG;
            for (; t < t0; t++) {
                pool_item(n, mean, M2, 1.f, ring_value(t, r), 0.f);
                r++; if (r >= A.DG) r -= A.DG;
            }
            const float* bs = BS + (size_t)(cls == 1 ? NBLK : 0) * 2 * N_MFCC;
            int bi = blk_idx_lo + (int)(((jl + t0) >> 3) - b_lo); if (bi >= NBLK) bi -= NBLK;
            static_assert(BLK == 8, "block index uses >> 3");
            for (; t + BLK - 1 <= tp.t_hi; t += BLK) {
                float mu, m2;
                if (cls < 2) {
                    const float* src = bs + (size_t)bi * 2 * N_MFCC;
                    mu = lane < N_MFCC ? src[lane] : 0.f;
                    m2 = lane < N_MFCC ? src[N_MFCC + lane] : 0.f;
                    r += BLK; if (r >= A.DG) r -= A.DG;
                } else {
                    float v[BLK];
#pragma unroll
                    for (int q = 0; q < BLK; q++) { v[q] = ring_value(t + q, r); r++; if (r >= A.DG) r -= A.DG; }
                    block_mean_m2(v, mu, m2);
                }
                pool_item(n, mean, M2, (float)BLK, mu, m2);
                bi++; if (bi >= NBLK) bi -= NBLK;
            }
            for (; t <= tp.t_hi; t++) {
                pool_item(n, mean, M2, 1.f, ring_value(t, r), 0.f);
                r++; if (r >= A.DG) r -= A.DG;
            }
            for (int e = 2; e < n_edge; e++) pool_item(n, mean, M2, 1.f, edge_value(e), 0.f);
            const float sd = sqrtf(M2 / (float)tp.F);
            const int nk = ft->n_mfcc;
            const float sc = similarity_score_warp(lane < nk ? tf.mean[lane] : 0.f, lane < nk ? tf.std[lane] : 0.f,
                                                   lane < nk ? mean : 0.f, lane < nk ? sd : 0.f);
            if (lane == 0) *outp = sc;
        }
        __syncthreads();
    }
    if (A.keep_rows && g_done > 2) {
        // keep the newest rows for the next call
        float* kr = A.keep_rows + (size_t)s * DENSE_KEEP * ROW;
        const long long lo = max(g_valid_lo, g_done - min(DENSE_KEEP, A.DG));
        const int cnt = (int)max(0LL, g_done - lo);
        for (int i = tid; i < cnt * ROW; i += DENSE_THREADS) {
            const long long g = lo + i / ROW;
            kr[(size_t)(g % DENSE_KEEP) * ROW + i % ROW] = G[(size_t)(g % A.DG) * ROW + i % ROW];
        }
        if (tid == 0) { A.keep_end[2 * s] = lo; A.keep_end[2 * s + 1] = g_done; }
    }
}

}  // namespace ewk

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
//   * mean / std over the F frames (two-pass), cosine vs the template, p^1.5/10.
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
constexpr int PATCH_CAP = 6;           // floored frames kept per warp before falling back to recomputation

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

__host__ __device__ inline size_t dense_smem_bytes(int DG, int T) {
    return sizeof(FrameTables) +
           sizeof(float) * ((size_t)DENSE_WARPS * SCR_WARP + (size_t)DG * ROW + (size_t)T * DH * 4 * ROW +
                            (size_t)DENSE_WARPS * (PATCH_CAP * N_MFCC + 2 * N_MFCC + 16) +
                            (size_t)DG * (N_MFCC + 1) + 4);
}

// frame `t` of the window of template `tp` starting at grid index j: pointer to its ROW
// (jb = j mod DG, so that the ring index is one add and one conditional subtract)
__device__ __forceinline__ const float* dense_row(const float* G, const float* edge_kh, const DenseTmplDev& tp, int DG,
                                                  int jb, int t) {
    if (t < 2) return edge_kh + t * ROW;
    if (t <= tp.t_hi) { int r = jb + t; if (r >= DG) r -= DG; return G + r * ROW; }
    return edge_kh + (2 + t - tp.t_hi - 1) * ROW;
}

__global__ void __launch_bounds__(DENSE_THREADS, 2)
dense_score_kernel(const DeviceTables* __restrict__ T, BankView B, const TemplateFeat* __restrict__ tmpl, DenseArgs A) {
    extern __shared__ __align__(16) float smem[];
    FrameTables* ft = reinterpret_cast<FrameTables*>(smem);
    float* scratch = smem + sizeof(FrameTables) / sizeof(float);
    float* G = scratch + DENSE_WARPS * SCR_WARP;                 // [DG][ROW]
    float* edge = G + (size_t)A.DG * ROW;                        // [T][DH][4][ROW]
    float* wbuf = edge + (size_t)A.T * DH * 4 * ROW;             // per warp: patch[PATCH_CAP][20], feat[40], masks[16]
    // alternate ring: rows of G recomputed with ONE floor value (the current one of the stream), tagged per row, so
    // that a floored stream-grid frame is recomputed once per floor value and not once per window that contains it
    float* G2 = wbuf + (size_t)DENSE_WARPS * (PATCH_CAP * N_MFCC + 2 * N_MFCC + 16);   // [DG][20]
    int* g2tag = reinterpret_cast<int*>(G2 + (size_t)A.DG * N_MFCC);                    // [DG] floor bits of the row
    int* fstar_s = g2tag + A.DG;                                                        // [1] floor bits served by G2

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int s = blockIdx.x;
    load_frame_tables(*ft, T, tid, DENSE_THREADS);
    float* scr = scratch + warp * SCR_WARP;
    init_warp_scratch(scr, lane);
    for (int i = tid; i < A.DG; i += DENSE_THREADS) g2tag[i] = 0x7fc00001;     // matches no floor
    __syncthreads();

    float* patch = wbuf + (size_t)warp * (PATCH_CAP * N_MFCC + 2 * N_MFCC + 16);
    float* feat = patch + PATCH_CAP * N_MFCC;
    unsigned* masks = reinterpret_cast<unsigned*>(feat + 2 * N_MFCC);   // [8] floored-frame bit masks, [8] prefix counts

    const size_t esz = B.fmt == 1 ? 2 : 4;
    PcmReader rd;
    rd.f = B.fmt == 0 ? (const float*)((const char*)B.ring + (size_t)s * B.P * esz) : nullptr;
    rd.q = B.fmt == 1 ? (const short*)((const char*)B.ring + (size_t)s * B.P * esz) : nullptr;
    rd.ring = B.P;

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
        for (int k = 0; k < A.T; k++) n_e += nh * (2 + A.t[k].r);
        // 32-bit bases for this sub-chunk: ring position of sample 160*hs and ring row of grid frame g_from
        const int hs_pos = (int)((160 * hs) % B.P);
        const int gfrom_row = n_g > 0 ? (int)(g_from % A.DG) : 0;
        const int gfrom_rel = (int)(g_from - hs);                  // grid index relative to hs
        for (int job = warp; job < n_g + n_e; job += DENSE_WARPS) {
            float* row;
            int f0;
            if (job < n_g) {
                // absolute first sample 160 g - 256 (unmasked frame of the stream grid)
                int pos = hs_pos + 160 * (gfrom_rel + job) - N_FFT / 2;
                pos %= B.P; if (pos < 0) pos += B.P;
                rd.start = pos; rd.len = N_FFT;
                f0 = 0;
                int r = gfrom_row + job; if (r >= A.DG) r -= A.DG;
                row = G + r * ROW;
                if (lane == 0) g2tag[r] = 0x7fc00001;
            } else {
                int rem = job - n_g, k = 0;
                while (rem >= nh * (2 + A.t[k].r)) { rem -= nh * (2 + A.t[k].r); k++; }
                const DenseTmplDev& tp = A.t[k];
                const int per = 2 + tp.r;
                const int hl = rem / per, e = rem - hl * per;
                if (hs + hl - tp.n < 0) continue;                   // window starts before the stream
                const int t = e < 2 ? e : tp.t_hi + 1 + (e - 2);
                int pos = hs_pos + 160 * (hl - tp.n);
                pos %= B.P; if (pos < 0) pos += B.P;
                rd.start = pos; rd.len = tp.L;                      // zeros outside the window
                f0 = t * HOP - N_FFT / 2;
                row = edge + ((k * DH + hl) * 4 + e) * ROW;
            }
            float2 x[8];
            load_frame_pairs_at(rd, f0, lane, x);
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
                    int pos = hs_pos + 160 * (glo_rel + i) - N_FFT / 2;
                    pos %= B.P; if (pos < 0) pos += B.P;
                    rd.start = pos; rd.len = N_FFT;
                    float2 x[8];
                    load_frame_pairs_at(rd, 0, lane, x);
                    float mn, mx;
                    warp_frame_mfcc(x, *ft, scr, lane, fstar, G2 + r * N_MFCC, mn, mx);
                    if (lane == 0) g2tag[r] = fbits;
                }
            __syncthreads();
        }

        // ---- windows of this sub-chunk, one warp per (template, hop)
        for (int w = warp; w < A.T * nh; w += DENSE_WARPS) {
            const int k = w / nh, hl = w - k * nh;
            const DenseTmplDev tp = A.t[k];
            const long long jl = hs + hl - tp.n;
            float* outp = A.out + ((size_t)s * A.n_hops + (size_t)(hs - A.hop0 + hl)) * A.T + k;
            if (jl < 0) { if (lane == 0) *outp = __int_as_float(0x7fc00000); continue; }
            const int j = (int)(jl % A.DG);                            // ring row base of this window
            const float* ekh = edge + (k * DH + hl) * 4 * ROW;
            // window max of the frames' log-mel max -> floor (librosa.power_to_db(top_db=80) on this window);
            // lanes walk the frames 32 at a time and keep their rows' min for the floored-frame test
            float wmax = -INFINITY;
            float fmin_c[DENSE_MAX_F / 32];
#pragma unroll
            for (int c = 0; c < DENSE_MAX_F / 32; c++) {
                const int t = c * 32 + lane;
                fmin_c[c] = INFINITY;
                if (t < tp.F) {
                    const float* row = dense_row(G, ekh, tp, A.DG, j, t);
                    fmin_c[c] = row[N_MFCC];
                    wmax = fmaxf(wmax, row[N_MFCC + 1]);
                }
            }
#pragma unroll
            for (int o = 16; o; o >>= 1) wmax = fmaxf(wmax, __shfl_xor_sync(FULL, wmax, o));
            const float floor_db = wmax - 80.0f;
            // frames that the floor changes
            int n_aff = 0;
#pragma unroll
            for (int c = 0; c < DENSE_MAX_F / 32; c++) {
                const unsigned m = __ballot_sync(FULL, fmin_c[c] < floor_db);
                if (lane == 0) { masks[c] = m; masks[8 + c] = (unsigned)n_aff; }
                n_aff += __popc(m);
            }
            __syncwarp();
            // G2 serves this window's floored stream-grid frames when its floor is the stream's current one
            const bool ringp = n_aff && __float_as_int(floor_db) == *fstar_s;
            if (n_aff) {
                // recompute the floored frames (window-local PCM view), keep the first PATCH_CAP of them
                // (ring-patched windows: only their edge frames, into slot = edge index)
                { int pos = hs_pos + 160 * (hl - tp.n); pos %= B.P; if (pos < 0) pos += B.P; rd.start = pos; }
                rd.len = tp.L;
                int slot = 0;
                if (ringp) {
                    for (int e = 0; e < 2 + tp.r; e++) {
                        const int t = e < 2 ? e : tp.t_hi + 1 + (e - 2);
                        if (!((masks[t >> 5] >> (t & 31)) & 1u)) continue;
                        float2 x[8];
                        load_frame_pairs_at(rd, t * HOP - N_FFT / 2, lane, x);
                        float mn, mx;
                        warp_frame_mfcc(x, *ft, scr, lane, floor_db, patch + e * N_MFCC, mn, mx);
                    }
                }
                for (int c = 0; !ringp && c * 32 < tp.F && slot < PATCH_CAP; c++) {
                    unsigned m = masks[c];
                    while (m && slot < PATCH_CAP) {
                        const int t = c * 32 + __ffs(m) - 1;
                        m &= m - 1;
                        float2 x[8];
                        load_frame_pairs_at(rd, t * HOP - N_FFT / 2, lane, x);
                        float mn, mx;
                        warp_frame_mfcc(x, *ft, scr, lane, floor_db, patch + slot * N_MFCC, mn, mx);
                        slot++;
                    }
                }
                __syncwarp();
            }
            // mean / std over the F frames: lane = (group g3 of 3, coefficient pair c2 of 10); three frames per step
            const int g3 = lane / 10, c2 = lane - 10 * g3;
            const bool act = lane < 30;
            float2 mean = make_float2(0.f, 0.f), var = make_float2(0.f, 0.f);
            for (int pass = 0; pass < 2; pass++) {
                float2 acc = make_float2(0.f, 0.f);
                if (act) {
                    if (!n_aff) {
                        // fast path: two left-edge rows, a run of ring rows, the right-edge rows
                        auto add = [&](const float* row) {
                            const float2 v = *reinterpret_cast<const float2*>(row + 2 * c2);
                            if (pass == 0) { acc.x += v.x; acc.y += v.y; }
                            else { const float dx = v.x - mean.x, dy = v.y - mean.y; acc.x = fmaf(dx, dx, acc.x); acc.y = fmaf(dy, dy, acc.y); }
                        };
                        if (g3 < 2) add(ekh + g3 * ROW);                               // t = 0, 1
                        if (g3 < tp.r) add(ekh + (2 + g3) * ROW);                      // t = t_hi+1 ..
                        int r = j + 2 + g3; if (r >= A.DG) r -= A.DG;
                        for (int t = 2 + g3; t <= tp.t_hi; t += 3) {
                            add(G + r * ROW);
                            r += 3; if (r >= A.DG) r -= A.DG;
                        }
                    } else {
                        for (int t = g3; t < tp.F; t += 3) {
                            float2 v;
                            bool patched = false;
                            const unsigned m = masks[t >> 5];
                            if ((m >> (t & 31)) & 1u) {
                                if (ringp) {
                                    if (t >= 2 && t <= tp.t_hi) { int r = j + t; if (r >= A.DG) r -= A.DG; v = *reinterpret_cast<const float2*>(G2 + r * N_MFCC + 2 * c2); }
                                    else v = *reinterpret_cast<const float2*>(patch + (t < 2 ? t : 2 + t - tp.t_hi - 1) * N_MFCC + 2 * c2);
                                    patched = true;
                                } else {
                                    const int sl = (int)masks[8 + (t >> 5)] + __popc(m & ((1u << (t & 31)) - 1u));
                                    if (sl < PATCH_CAP) { v = *reinterpret_cast<const float2*>(patch + sl * N_MFCC + 2 * c2); patched = true; }
                                }
                            }
                            if (!patched) v = *reinterpret_cast<const float2*>(dense_row(G, ekh, tp, A.DG, j, t) + 2 * c2);
                            if (pass == 0) { acc.x += v.x; acc.y += v.y; }
                            else { const float dx = v.x - mean.x, dy = v.y - mean.y; acc.x = fmaf(dx, dx, acc.x); acc.y = fmaf(dy, dy, acc.y); }
                        }
                    }
                }
                // combine the three frame groups (lanes c2, c2+10, c2+20)
                const float ax = __shfl_sync(FULL, acc.x, c2) + __shfl_sync(FULL, acc.x, c2 + 10) + __shfl_sync(FULL, acc.x, c2 + 20);
                const float ay = __shfl_sync(FULL, acc.y, c2) + __shfl_sync(FULL, acc.y, c2 + 10) + __shfl_sync(FULL, acc.y, c2 + 20);
                if (pass == 0) mean = make_float2(ax / (float)tp.F, ay / (float)tp.F);
                else var = make_float2(ax / (float)tp.F, ay / (float)tp.F);
            }
            // floored frames beyond the patch capacity: fold their corrections in by recomputation (rare)
            if (!ringp && n_aff > PATCH_CAP) {
                // second-order exactness is kept by redoing both passes with on-the-fly recomputation
                float2 m2 = make_float2(0.f, 0.f), v2 = make_float2(0.f, 0.f);
                for (int pass = 0; pass < 2; pass++) {
                    float accx = 0.f, accy = 0.f;
                    for (int t = 0; t < tp.F; t++) {
                        const unsigned m = masks[t >> 5];
                        const bool a = (m >> (t & 31)) & 1u;
                        const float* src;
                        if (a) {
                            float2 x[8];
                            load_frame_pairs_at(rd, t * HOP - N_FFT / 2, lane, x);
                            float mn, mx;
                            warp_frame_mfcc(x, *ft, scr, lane, floor_db, patch, mn, mx);
                            __syncwarp();
                            src = patch;
                        } else src = dense_row(G, ekh, tp, A.DG, j, t);
                        if (lane < 10) {
                            const float2 v = *reinterpret_cast<const float2*>(src + 2 * lane);
                            if (pass == 0) { accx += v.x; accy += v.y; }
                            else { const float dx = v.x - m2.x, dy = v.y - m2.y; accx = fmaf(dx, dx, accx); accy = fmaf(dy, dy, accy); }
                        }
                        __syncwarp();
                    }
                    if (pass == 0) m2 = make_float2(accx / (float)tp.F, accy / (float)tp.F);
                    else v2 = make_float2(accx / (float)tp.F, accy / (float)tp.F);
                }
                mean = make_float2(__shfl_sync(FULL, m2.x, c2), __shfl_sync(FULL, m2.y, c2));
                var = make_float2(__shfl_sync(FULL, v2.x, c2), __shfl_sync(FULL, v2.y, c2));
            }
            if (lane < 10) {
                feat[2 * lane] = mean.x; feat[2 * lane + 1] = mean.y;
                feat[N_MFCC + 2 * lane] = sqrtf(var.x); feat[N_MFCC + 2 * lane + 1] = sqrtf(var.y);
            }
            __syncwarp();
            if (lane == 0) {
                const TemplateFeat& tf = tmpl[tp.slot];
                *outp = tf.valid ? similarity_score(tf.mean, tf.std, feat, feat + N_MFCC) : __int_as_float(0x7fc00000);
            }
            __syncwarp();
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

// libewk: C ABI (include/ewk.h) over the sm_100a kernels.  No torch types, no CPU fallback.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/ewk.h"
#include "ewk_ctx.hpp"

using namespace ewk;

static thread_local std::string g_create_error = "";

void ewk_ctx::fail(const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    err = buf;
}

extern "C" int ewk_abi_version(void) { return EWK_ABI_VERSION; }

extern "C" const char* ewk_last_error(const ewk_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

extern "C" int ewk_host_table(int which, float* out, int cap) {
    static DeviceTables T;
    static std::vector<float> mel;
    static bool built = false;
    if (!built) { build_tables(T); build_mel_dense(mel); built = true; }
    if (which == 0) {
        if (out && cap >= N_FFT) std::memcpy(out, T.hann, sizeof(float) * N_FFT);
        return N_FFT;
    }
    if (which == 1) {
        if (out && cap >= N_MELS * N_BINS) std::memcpy(out, mel.data(), sizeof(float) * N_MELS * N_BINS);
        return N_MELS * N_BINS;
    }
    if (which == 2) {
        if (out && cap >= N_MFCC * N_MELS)
            for (int k = 0; k < N_MFCC; k++)
                for (int b = 0; b < N_MELS; b++) out[k * N_MELS + b] = T.dct_t[b * N_MFCC + k];
        return N_MFCC * N_MELS;
    }
    return EWK_ERR_ARG;
}

#define CK(call)                                                                                    \
    do {                                                                                            \
        cudaError_t e_ = (call);                                                                    \
        if (e_ != cudaSuccess) {                                                                    \
            ctx->fail("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__);  \
            return EWK_ERR_CUDA;                                                                    \
        }                                                                                           \
    } while (0)

extern "C" int ewk_create(int device, const ewk_config* cfg, ewk_ctx** out) {
    if (!cfg || !out) { g_create_error = "ewk_create: null argument"; return EWK_ERR_ARG; }
    *out = nullptr;
    if (cfg->n_streams < 0 || cfg->max_templates < 1 || cfg->max_templates > EWK_MAX_TEMPLATES ||
        (cfg->pcm_format != EWK_PCM_F32 && cfg->pcm_format != EWK_PCM_I16) ||
        (cfg->n_streams > 0 && (cfg->ring_samples < EWK_TICK_SAMPLES || cfg->slack_samples < 0))) {
        g_create_error = "ewk_create: invalid ewk_config";
        return EWK_ERR_ARG;
    }
    if (!(cfg->preemphasis >= 0.f && cfg->preemphasis < 1.f)) {
        g_create_error = "ewk_create: preemphasis must be in [0, 1)";
        return EWK_ERR_ARG;
    }
    if (cfg->n_mfcc < 0 || cfg->n_mfcc > EWK_N_MFCC || cfg->reserved0 != 0 || cfg->reserved1 != 0) {
        g_create_error = "ewk_create: n_mfcc must be 0 (= 20) or in [1, 20]; reserved fields must be 0";
        return EWK_ERR_ARG;
    }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        g_create_error = std::string("ewk_create: no CUDA device (") + cudaGetErrorString(e) +
                         "); libewk has no CPU fallback";
        return EWK_ERR_CUDA;
    }
    if (device < 0 || device >= ndev) { g_create_error = "ewk_create: device index out of range"; return EWK_ERR_ARG; }
    ewk_ctx* ctx = new ewk_ctx();
    ctx->device = device;
    ctx->cfg = *cfg;
    int rc = ctx->init();
    if (rc != EWK_OK) {
        g_create_error = ctx->err;
        ctx->release();
        delete ctx;
        return rc;
    }
    *out = ctx;
    return EWK_OK;
}

extern "C" int ewk_destroy(ewk_ctx* ctx) {
    if (!ctx) return EWK_ERR_ARG;
    cudaSetDevice(ctx->device);
    ctx->release();
    delete ctx;
    return EWK_OK;
}

extern "C" int ewk_set_cuda_stream(ewk_ctx* ctx, void* s) {
    if (!ctx) return EWK_ERR_ARG;
    if (ctx->match_inflight) {                       // results of the old stream's ticks stay ordered before the switch
        CK(cudaSetDevice(ctx->device));
        CK(cudaStreamSynchronize(ctx->match_stream));
        ctx->match_inflight = false;
    }
    ctx->stream = s ? (cudaStream_t)s : ctx->own_stream;
    return EWK_OK;
}

extern "C" int ewk_synchronize(ewk_ctx* ctx) {
    if (!ctx) return EWK_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    if (ctx->bank.n_streams) { int rc = ctx->flush_pending(); if (rc) return rc; }
    int jr = ctx->join_match();
    if (jr) return jr;
    CK(cudaStreamSynchronize(ctx->stream));
    if (ctx->pub_stream) CK(cudaStreamSynchronize(ctx->pub_stream));
    return EWK_OK;
}

int ewk_ctx::init() {
    ewk_ctx* ctx = this;
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    sm_count = prop.multiProcessorCount;
    if (const char* e = getenv("EWK_SEG_LM")) use_lm = atoi(e) != 0;       // 0: recompute floored frames from the PCM
    if (prop.major < 10) {
        fail("ewk_create: device %d is sm_%d%d; libewk is built for sm_100a only", device, prop.major, prop.minor);
        return EWK_ERR_CUDA;
    }
    CK(cudaStreamCreateWithFlags(&own_stream, cudaStreamNonBlocking));
    stream = own_stream;
    CK(cudaStreamCreateWithFlags(&copy_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; i++) {
        CK(cudaEventCreateWithFlags(&ev_ready[i], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&ev_free[i], cudaEventDisableTiming));
    }
    DeviceTables* h = new DeviceTables();
    int nnz = build_tables(*h);
    if (nnz < 0) { delete h; fail("mel table overflow"); return EWK_ERR_STATE; }
    // flat tables, then the per-CTA layout of them (FrameTables image) that K3 / K4 copy into shared memory
    FrameTables* img = new FrameTables();
    load_frame_tables(*img, h, 0, 1);
    img->preemph = cfg.preemphasis;
    img->n_mfcc = cfg.n_mfcc > 0 ? cfg.n_mfcc : N_MFCC;
    img->pad_[0] = img->pad_[1] = 0;
    CK(cudaMalloc((void**)&d_tables, FRAME_IMAGE_OFFSET + sizeof(FrameTables)));
    CK(cudaMemcpy(d_tables, h, sizeof(DeviceTables), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(reinterpret_cast<char*>(d_tables) + FRAME_IMAGE_OFFSET, img, sizeof(FrameTables), cudaMemcpyHostToDevice));
    delete img;
    delete h;
    h_tmpl.assign(cfg.max_templates, TemplateFeat{});
    CK(cudaMalloc(&d_tmpl, sizeof(TemplateFeat) * cfg.max_templates));
    CK(cudaMemset(d_tmpl, 0, sizeof(TemplateFeat) * cfg.max_templates));
    CK(cudaFuncSetAttribute(segment_mfcc_match_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            (int)seg_smem_bytes(SEG_SMEM_FRAMES)));
    CK(cudaFuncSetAttribute(segment_mfcc_match_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            (int)seg_smem_bytes(SEG_SMEM_FRAMES)));
    int rc = init_streams();
    if (rc != EWK_OK) return rc;
    return EWK_OK;
}

void ewk_ctx::release() {
    cudaSetDevice(device);
    if (own_stream) cudaStreamSynchronize(own_stream);
    release_streams();
    for (DevBuf* b : {&b_pcm, &b_desc, &b_ws, &b_lm, &b_feat, &b_scores, &b_matched, &b_frames, &b_off}) b->free();
    for (auto& kv : rs_tables) if (kv.second.H) cudaFree(kv.second.H);
    rs_tables.clear();
    b_rs_in.free(); b_rs_out.free();
    if (d_wait_flag) { cudaFree(d_wait_flag); d_wait_flag = nullptr; }
    if (pub_stream) { cudaStreamSynchronize(pub_stream); cudaStreamDestroy(pub_stream); pub_stream = nullptr; }
    if (ev_k3) { cudaEventDestroy(ev_k3); ev_k3 = nullptr; }
    for (int i = 0; i < 2; i++) if (ev_pub[i]) { cudaEventDestroy(ev_pub[i]); ev_pub[i] = nullptr; }
    if (d_pub_snap) { cudaFree(d_pub_snap); d_pub_snap = nullptr; }
    if (d_tables) cudaFree(d_tables);
    if (d_tmpl) cudaFree(d_tmpl);
    if (copy_stream) { cudaStreamSynchronize(copy_stream); cudaStreamDestroy(copy_stream); }
    if (match_stream) { cudaStreamSynchronize(match_stream); cudaStreamDestroy(match_stream); match_stream = nullptr; }
    if (ev_gate) { cudaEventDestroy(ev_gate); ev_gate = nullptr; }
    if (ev_match) { cudaEventDestroy(ev_match); ev_match = nullptr; }
    for (int i = 0; i < 2; i++) {
        if (ev_ready[i]) cudaEventDestroy(ev_ready[i]);
        if (ev_free[i]) cudaEventDestroy(ev_free[i]);
        ev_ready[i] = ev_free[i] = nullptr;
        b_stage2[i].free();
        b_raw[i].free();
    }
    if (own_stream) cudaStreamDestroy(own_stream);
    d_tables = nullptr; d_tmpl = nullptr; own_stream = nullptr; copy_stream = nullptr;
}

// ------------------------------------------------------------------------------------------
// Launch K3 over `n_seg` descriptors already on the device.
int ewk_ctx::launch_segments(const SegDesc* d_segs, int n_seg, int max_frames, long long spill_frames,
                             long long lm_frames, int n_tmpl, int tmpl_first, float threshold, float* d_feat,
                             float* d_frames, float* d_scores, unsigned char* d_matched) {
    ewk_ctx* ctx = this;
    const int cap = std::min(max_frames, SEG_SMEM_FRAMES);
    float* ws = nullptr;
    if (spill_frames > 0) {
        CK(b_ws.ensure(sizeof(float) * (size_t)spill_frames * FR_STRIDE));
        ws = (float*)b_ws.p;
    }
    // log-mel rows of every frame (512 B each), so that the top_db floor costs one DCT per floored frame; beyond
    // LM_WS_MAX_BYTES the kernel recomputes floored frames from the PCM instead
    float* lm = nullptr;
    if (use_lm && lm_frames > 0 && sizeof(float) * (size_t)lm_frames * LM_ROW <= LM_WS_MAX_BYTES) {
        CK(b_lm.ensure(sizeof(float) * (size_t)lm_frames * LM_ROW));
        lm = (float*)b_lm.p;
    }
    cudaEvent_t pe = prof_begin(3);
    auto k3 = cfg.preemphasis != 0.f ? segment_mfcc_match_kernel<true> : segment_mfcc_match_kernel<false>;
    k3<<<n_seg, SEG_THREADS, seg_smem_bytes(cap), stream>>>(
        d_tables, d_segs, cap, ws, lm, d_tmpl, n_tmpl, tmpl_first, threshold, d_feat, d_frames, d_scores, d_matched);
    prof_end(pe, 3);
    CK(cudaGetLastError());
    launches++;
    return EWK_OK;
}

// WordMatcher.extract_mfcc on one buffer; dense40 (nullable): the same features from K4's integer statistics of the frames.
static int extract_impl(ewk_ctx* ctx, const float* pcm, int64_t n, int where, float* mean20, float* std20,
                        float* frames, int64_t frames_cap, float* dense40) {
    if (!pcm || n < 1 || n > (int64_t)1 << 30 || !mean20 || !std20) {
        ctx->fail("ewk_extract_mfcc: need pcm, 1 <= n < 2^30, mean20, std20 (n=%lld)", (long long)n);
        return EWK_ERR_ARG;
    }
    const int64_t F = 1 + n / HOP;
    if (frames && frames_cap < F) { ctx->fail("ewk_extract_mfcc: frames_cap %lld < %lld frames", (long long)frames_cap, (long long)F); return EWK_ERR_ARG; }
    CK(cudaSetDevice(ctx->device));
    const float* d_pcm = pcm;
    if (where == EWK_HOST) {
        CK(ctx->b_pcm.ensure(sizeof(float) * (size_t)n));
        CK(cudaMemcpyAsync(ctx->b_pcm.p, pcm, sizeof(float) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
        d_pcm = (const float*)ctx->b_pcm.p;
    }
    SegDesc sd{};
    sd.base = d_pcm; sd.start = 0; sd.ring = 0; sd.len = (int)n; sd.fmt = 0; sd.ws_frame_off = 0; sd.frames_off = 0;
    sd.lm_off = 0;
    CK(ctx->b_desc.ensure(sizeof(SegDesc)));
    CK(cudaMemcpyAsync(ctx->b_desc.p, &sd, sizeof(sd), cudaMemcpyHostToDevice, ctx->stream));
    CK(ctx->b_feat.ensure(sizeof(float) * 2 * FEAT));
    float* d_frames = nullptr;
    if (frames || dense40) { CK(ctx->b_frames.ensure(sizeof(float) * (size_t)F * N_MFCC)); d_frames = (float*)ctx->b_frames.p; }
    int rc = ctx->launch_segments((const SegDesc*)ctx->b_desc.p, 1, (int)F, F > SEG_SMEM_FRAMES ? F : 0, F, 0, 0, 0.f,
                                  (float*)ctx->b_feat.p, d_frames, nullptr, nullptr);
    if (rc != EWK_OK) return rc;
    if (dense40) {
        dense_template_features_kernel<<<1, 32, 0, ctx->stream>>>(d_frames, (int)F, ctx->cfg.n_mfcc > 0 ? ctx->cfg.n_mfcc : N_MFCC,
                                                                  (float*)ctx->b_feat.p + FEAT);
        CK(cudaGetLastError());
        ctx->launches++;
    }
    float feat[2 * FEAT];
    CK(cudaMemcpyAsync(feat, ctx->b_feat.p, sizeof(float) * (dense40 ? 2 * FEAT : FEAT), cudaMemcpyDeviceToHost, ctx->stream));
    if (frames) CK(cudaMemcpyAsync(frames, d_frames, sizeof(float) * (size_t)F * N_MFCC, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    std::memcpy(mean20, feat, sizeof(float) * N_MFCC);
    std::memcpy(std20, feat + N_MFCC, sizeof(float) * N_MFCC);
    if (dense40) std::memcpy(dense40, feat + FEAT, sizeof(float) * FEAT);
    return EWK_OK;
}

extern "C" int ewk_extract_mfcc(ewk_ctx* ctx, const float* pcm, int64_t n, int where, float* mean20, float* std20,
                                float* frames, int64_t frames_cap) {
    if (!ctx) return EWK_ERR_ARG;
    return extract_impl(ctx, pcm, n, where, mean20, std20, frames, frames_cap, nullptr);
}

static int check_slot(ewk_ctx* ctx, int slot, const char* who) {
    if (slot < 0 || slot >= ctx->cfg.max_templates) {
        ctx->fail("%s: template slot %d out of range [0, %d)", who, slot, ctx->cfg.max_templates);
        return EWK_ERR_ARG;
    }
    return EWK_OK;
}

static int install_template(ewk_ctx* ctx, int slot, const float* mean20, const float* std20, const float* dense40, int64_t n_samples) {
    CK(cudaSetDevice(ctx->device));
    TemplateFeat tf{};
    std::memcpy(tf.mean, mean20, sizeof(tf.mean));
    std::memcpy(tf.std, std20, sizeof(tf.std));
    // features for the dense kernel: its own integer statistics of the template's frames when the audio is known,
    // the given features otherwise
    std::memcpy(tf.dmean, dense40 ? dense40 : mean20, sizeof(tf.dmean));
    std::memcpy(tf.dstd, dense40 ? dense40 + N_MFCC : std20, sizeof(tf.dstd));
    tf.n_samples = n_samples;
    tf.valid = 1;
    ctx->h_tmpl[slot] = tf;
    CK(cudaMemcpyAsync(ctx->d_tmpl + slot, &ctx->h_tmpl[slot], sizeof(tf), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return EWK_OK;
}

extern "C" int ewk_set_template_features(ewk_ctx* ctx, int slot, const float* mean20, const float* std20, int64_t n_samples) {
    if (!ctx) return EWK_ERR_ARG;
    int rc = check_slot(ctx, slot, "ewk_set_template_features");
    if (rc) return rc;
    if (!mean20 || !std20 || n_samples < 1) { ctx->fail("ewk_set_template_features: null features or n_samples < 1"); return EWK_ERR_ARG; }
    return install_template(ctx, slot, mean20, std20, nullptr, n_samples);
}

extern "C" int ewk_set_template(ewk_ctx* ctx, int slot, const float* pcm, int64_t n) {
    if (!ctx) return EWK_ERR_ARG;
    int rc = check_slot(ctx, slot, "ewk_set_template");
    if (rc) return rc;
    float mean[N_MFCC], sd[N_MFCC], dense[FEAT];
    rc = extract_impl(ctx, pcm, n, EWK_HOST, mean, sd, nullptr, 0, dense);
    if (rc) return rc;
    return install_template(ctx, slot, mean, sd, dense, n);
}

extern "C" int ewk_get_template(ewk_ctx* ctx, int slot, float* mean20, float* std20, int64_t* n_samples) {
    if (!ctx) return EWK_ERR_ARG;
    int rc = check_slot(ctx, slot, "ewk_get_template");
    if (rc) return rc;
    const TemplateFeat& tf = ctx->h_tmpl[slot];
    if (!tf.valid) { ctx->fail("No reference word set. Call set_reference() first."); return EWK_ERR_NO_TEMPLATE; }
    if (mean20) std::memcpy(mean20, tf.mean, sizeof(tf.mean));
    if (std20) std::memcpy(std20, tf.std, sizeof(tf.std));
    if (n_samples) *n_samples = tf.n_samples;
    return EWK_OK;
}

extern "C" int ewk_clear_template(ewk_ctx* ctx, int slot) {
    if (!ctx) return EWK_ERR_ARG;
    int rc = check_slot(ctx, slot, "ewk_clear_template");
    if (rc) return rc;
    CK(cudaSetDevice(ctx->device));
    ctx->h_tmpl[slot] = TemplateFeat{};
    CK(cudaMemcpyAsync(ctx->d_tmpl + slot, &ctx->h_tmpl[slot], sizeof(TemplateFeat), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return EWK_OK;
}

extern "C" int ewk_similarity_batch(ewk_ctx* ctx, int slot, const void* pcm, int pcm_format, int where,
                                    const int64_t* offsets, const int64_t* lens, int n_seg, float threshold,
                                    float* scores, uint8_t* matched, float* features) {
    if (!ctx) return EWK_ERR_ARG;
    int rc = check_slot(ctx, slot, "ewk_similarity_batch");
    if (rc) return rc;
    if (!ctx->h_tmpl[slot].valid) { ctx->fail("No reference word set. Call set_reference() first."); return EWK_ERR_NO_TEMPLATE; }
    if (n_seg == 0) return EWK_OK;
    if (!pcm || !offsets || !lens || n_seg < 0 || !scores || (pcm_format != EWK_PCM_F32 && pcm_format != EWK_PCM_I16)) {
        ctx->fail("ewk_similarity_batch: bad arguments");
        return EWK_ERR_ARG;
    }
    const size_t esz = pcm_format == EWK_PCM_I16 ? 2 : 4;
    int64_t extent = 0, spill = 0, lm_frames = 0;
    int max_frames = 1;
    std::vector<SegDesc> sd(n_seg);
    for (int i = 0; i < n_seg; i++) {
        if (offsets[i] < 0 || lens[i] < 1 || lens[i] > (int64_t)1 << 30) {
            ctx->fail("ewk_similarity_batch: segment %d has offset %lld len %lld", i, (long long)offsets[i], (long long)lens[i]);
            return EWK_ERR_ARG;
        }
        extent = std::max(extent, offsets[i] + lens[i]);
        const int F = 1 + (int)(lens[i] / HOP);
        max_frames = std::max(max_frames, F);
        sd[i] = SegDesc{};
        sd[i].start = offsets[i]; sd[i].ring = 0; sd[i].len = (int)lens[i]; sd[i].fmt = pcm_format == EWK_PCM_I16 ? 1 : 0;
        sd[i].ws_frame_off = (int)spill;
        if (F > SEG_SMEM_FRAMES) spill += F;
        sd[i].lm_off = lm_frames;
        lm_frames += F;
    }
    CK(cudaSetDevice(ctx->device));
    const void* d_pcm = pcm;
    if (where == EWK_HOST) {
        CK(ctx->b_pcm.ensure(esz * (size_t)extent));
        CK(cudaMemcpyAsync(ctx->b_pcm.p, pcm, esz * (size_t)extent, cudaMemcpyHostToDevice, ctx->stream));
        d_pcm = ctx->b_pcm.p;
    }
    for (int i = 0; i < n_seg; i++) sd[i].base = d_pcm;
    CK(ctx->b_desc.ensure(sizeof(SegDesc) * (size_t)n_seg));
    CK(cudaMemcpyAsync(ctx->b_desc.p, sd.data(), sizeof(SegDesc) * (size_t)n_seg, cudaMemcpyHostToDevice, ctx->stream));
    CK(ctx->b_feat.ensure(sizeof(float) * FEAT * (size_t)n_seg));
    CK(ctx->b_scores.ensure(sizeof(float) * (size_t)n_seg));
    CK(ctx->b_matched.ensure((size_t)n_seg));
    rc = ctx->launch_segments((const SegDesc*)ctx->b_desc.p, n_seg, max_frames, spill, lm_frames, 1, slot, threshold,
                              (float*)ctx->b_feat.p, nullptr, (float*)ctx->b_scores.p, (unsigned char*)ctx->b_matched.p);
    if (rc != EWK_OK) return rc;
    CK(cudaMemcpyAsync(scores, ctx->b_scores.p, sizeof(float) * (size_t)n_seg, cudaMemcpyDeviceToHost, ctx->stream));
    if (matched) CK(cudaMemcpyAsync(matched, ctx->b_matched.p, (size_t)n_seg, cudaMemcpyDeviceToHost, ctx->stream));
    if (features) CK(cudaMemcpyAsync(features, ctx->b_feat.p, sizeof(float) * FEAT * (size_t)n_seg, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return EWK_OK;
}

extern "C" int ewk_analyze_templates(ewk_ctx* ctx, const float* pcm, int where, const int64_t* offsets,
                                     const int64_t* lens, int n, ewk_vad_result* out, float* rms_out, int64_t rms_cap) {
    static_assert(sizeof(ewk_vad_result) == sizeof(VadResult) && sizeof(VadResult) == 32, "ewk_vad_result layout");
    if (!ctx) return EWK_ERR_ARG;
    if (n == 0) return EWK_OK;
    if (!pcm || !offsets || !lens || n < 0 || !out) { ctx->fail("ewk_analyze_templates: bad arguments"); return EWK_ERR_ARG; }
    int64_t extent = 0, frames = 0;
    std::vector<VadDesc> vd(n);
    for (int i = 0; i < n; i++) {
        if (offsets[i] < 0 || lens[i] < 0 || lens[i] > (int64_t)1 << 32) {
            ctx->fail("ewk_analyze_templates: template %d has offset %lld len %lld", i, (long long)offsets[i], (long long)lens[i]);
            return EWK_ERR_ARG;
        }
        extent = std::max(extent, offsets[i] + lens[i]);
        vd[i].start = offsets[i]; vd[i].len = lens[i]; vd[i].rms_off = frames;
        frames += 1 + lens[i] / VAD_HOP;
    }
    if (rms_out && rms_cap < frames) { ctx->fail("ewk_analyze_templates: rms_cap %lld < %lld frames", (long long)rms_cap, (long long)frames); return EWK_ERR_ARG; }
    CK(cudaSetDevice(ctx->device));
    const float* d_pcm = pcm;
    if (where == EWK_HOST) {
        CK(ctx->b_pcm.ensure(sizeof(float) * (size_t)std::max<int64_t>(extent, 1)));
        CK(cudaMemcpyAsync(ctx->b_pcm.p, pcm, sizeof(float) * (size_t)extent, cudaMemcpyHostToDevice, ctx->stream));
        d_pcm = (const float*)ctx->b_pcm.p;
    }
    for (int i = 0; i < n; i++) vd[i].base = d_pcm;
    CK(ctx->b_desc.ensure(sizeof(VadDesc) * (size_t)n));
    CK(cudaMemcpyAsync(ctx->b_desc.p, vd.data(), sizeof(VadDesc) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    CK(ctx->b_feat.ensure(sizeof(VadResult) * (size_t)n));
    CK(ctx->b_frames.ensure(sizeof(float) * (size_t)frames));
    cudaEvent_t pe = ctx->prof_begin(3);
    template_vad_kernel<<<n, VAD_THREADS, 0, ctx->stream>>>((const VadDesc*)ctx->b_desc.p, (VadResult*)ctx->b_feat.p,
                                                            (float*)ctx->b_frames.p);
    ctx->prof_end(pe, 3);
    CK(cudaGetLastError());
    ctx->launches++;
    CK(cudaMemcpyAsync(out, ctx->b_feat.p, sizeof(VadResult) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    if (rms_out) CK(cudaMemcpyAsync(rms_out, ctx->b_frames.p, sizeof(float) * (size_t)frames, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return EWK_OK;
}

// ------------------------------------------------------------------------------------------ K7 resampling
extern "C" int ewk_resample_info(int sr_in, int32_t* half_width, int32_t* up, int32_t* down) {
    ResampleDesign d;
    if (sr_in < 1000 || sr_in > 768000 || !rs_design(sr_in, d)) return EWK_ERR_ARG;
    if (half_width) *half_width = d.W;
    if (up) *up = d.L;
    if (down) *down = d.M;
    return EWK_OK;
}

extern "C" int64_t ewk_resample_out_len(int64_t n_in, int sr_in) {
    if (n_in <= 0 || sr_in <= 0) return 0;
    return (int64_t)std::ceil((double)n_in * RS_TARGET / (double)sr_in);     // librosa.resample: int(np.ceil(n * ratio))
}

int ewk_ctx::resample_table(int sr_in, const ResampleTable** out) {
    ewk_ctx* ctx = this;
    auto it = rs_tables.find(sr_in);
    if (it == rs_tables.end()) {
        ResampleTable t;
        if (sr_in < 1000 || sr_in > 768000 || !rs_design(sr_in, t.d)) {
            fail("ewk_resample: %d Hz -> 16000 Hz is not supported (needs more than %d filter phases)", sr_in, RS_MAX_PHASES);
            return EWK_ERR_ARG;
        }
        std::vector<float> H;
        rs_build_table(t.d, H);
        CK(cudaMalloc(&t.H, sizeof(float) * H.size()));
        CK(cudaMemcpyAsync(t.H, H.data(), sizeof(float) * H.size(), cudaMemcpyHostToDevice, stream));
        CK(cudaStreamSynchronize(stream));
        it = rs_tables.emplace(sr_in, t).first;
    }
    *out = &it->second;
    return EWK_OK;
}

extern "C" int ewk_resample(ewk_ctx* ctx, const void* in, int pcm_format, int where_in, int n_rows, int64_t n_in,
                            int64_t in_stride, int sr_in, int64_t in_first, int64_t out_first, int64_t n_out, float* out,
                            int64_t out_stride, int where_out) {
    if (!ctx) return EWK_ERR_ARG;
    if (n_rows == 0 || n_out == 0) return EWK_OK;
    if (!in || !out || n_rows < 0 || n_in < 0 || n_out < 0 || in_stride < n_in || out_stride < n_out || out_first < 0 ||
        (pcm_format != EWK_PCM_F32 && pcm_format != EWK_PCM_I16) || n_in > (int64_t)1 << 34 || n_out > (int64_t)1 << 34) {
        ctx->fail("ewk_resample: bad arguments");
        return EWK_ERR_ARG;
    }
    if (sr_in == RS_TARGET) { ctx->fail("ewk_resample: input is already at 16000 Hz"); return EWK_ERR_ARG; }
    CK(cudaSetDevice(ctx->device));
    const ewk_ctx::ResampleTable* t = nullptr;
    int rc = ctx->resample_table(sr_in, &t);
    if (rc) return rc;
    const size_t esz = pcm_format == EWK_PCM_I16 ? 2 : 4;
    ResampleArgs A{};
    A.in = in; A.in_stride = in_stride;
    if (where_in == EWK_HOST) {
        CK(ctx->b_rs_in.ensure(esz * (size_t)n_rows * (size_t)std::max<int64_t>(n_in, 1)));
        if (n_in > 0)
            CK(cudaMemcpy2DAsync(ctx->b_rs_in.p, esz * (size_t)n_in, in, esz * (size_t)in_stride, esz * (size_t)n_in,
                                 (size_t)n_rows, cudaMemcpyHostToDevice, ctx->stream));
        A.in = ctx->b_rs_in.p; A.in_stride = n_in;
    }
    A.out = out; A.out_stride = out_stride;
    if (where_out == EWK_HOST) {
        CK(ctx->b_rs_out.ensure(sizeof(float) * (size_t)n_rows * (size_t)n_out));
        A.out = (float*)ctx->b_rs_out.p; A.out_stride = n_out;
    }
    A.H = t->H; A.n_in = n_in; A.in_first = in_first; A.out_first = out_first; A.n_out = n_out;
    A.L = t->d.L; A.M = t->d.M; A.Minv = t->d.Minv; A.W = t->d.W; A.fmt = pcm_format == EWK_PCM_I16 ? 1 : 0;
    const long long g0 = out_first / A.L, g1 = (out_first + n_out + A.L - 1) / A.L;
    const long long threads = (g1 - g0) * A.L;
    const dim3 grid((unsigned)((threads + RS_THREADS - 1) / RS_THREADS), (unsigned)n_rows);
    if (n_rows > 65535) { ctx->fail("ewk_resample: at most 65535 rows per call"); return EWK_ERR_ARG; }
    cudaEvent_t pe = ctx->prof_begin(3);
    if (A.fmt == 1) resample_kernel<short><<<grid, RS_THREADS, 0, ctx->stream>>>(A);
    else resample_kernel<float><<<grid, RS_THREADS, 0, ctx->stream>>>(A);
    ctx->prof_end(pe, 3);
    CK(cudaGetLastError());
    ctx->launches++;
    if (where_out == EWK_HOST) {
        CK(cudaMemcpy2DAsync(out, sizeof(float) * (size_t)out_stride, ctx->b_rs_out.p, sizeof(float) * (size_t)n_out,
                             sizeof(float) * (size_t)n_out, (size_t)n_rows, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    return EWK_OK;
}

// ==========================================================================================
// stream bank
// ==========================================================================================
static_assert(sizeof(StreamParams) == sizeof(ewk_stream_params), "ewk_stream_params layout");
static_assert(sizeof(EventRec) == sizeof(ewk_event), "ewk_event layout");
static_assert(sizeof(StreamResult) == sizeof(ewk_stream_result), "ewk_stream_result layout");
constexpr int MIN_FRAME_SIZE = 160;
constexpr int GATE_SMEM_CHUNKS = 2048;   // storage-order chunks a warp of K2 can hold in shared memory (3 arrays x 4 warps)

extern "C" int ewk_default_stream_params(ewk_stream_params* p) {
    if (!p) return EWK_ERR_ARG;
    std::memset(p, 0, sizeof(*p));
    p->similarity_threshold = 75.0f;
    p->frame_size = 0;
    p->pre_speech_silence = 0.8;
    p->speech_duration_min = 0.3;
    p->speech_duration_max = 2.0;
    p->post_speech_silence = 0.4;
    p->timeout = 30.0;
    p->min_threshold = 0.005;
    p->template_first = 0;
    p->template_count = 1;
    p->live = 0;
    return EWK_OK;
}

int ewk_ctx::init_streams() {
    ewk_ctx* ctx = this;
    const int n = cfg.n_streams;
    if (n == 0) return EWK_OK;
    const size_t esz = cfg.pcm_format == EWK_PCM_I16 ? 2 : 4;
    bank.n_streams = n;
    bank.R = cfg.ring_samples;
    bank.P = ((cfg.ring_samples + cfg.slack_samples + 63) / 64) * 64;
    bank.fmt = cfg.pcm_format == EWK_PCM_I16 ? 1 : 0;
    chunk_cap = std::max(1, cfg.ring_samples / MIN_FRAME_SIZE);
    bank.chunk_cap = chunk_cap;
    bank.max_events = std::max(16, cfg.max_events);
    CK(cudaMalloc(&bank.ring, esz * (size_t)n * bank.P));
    CK(cudaMemsetAsync(bank.ring, 0, esz * (size_t)n * bank.P, stream));          // np.zeros   wakeword.py:428
    CK(cudaMalloc(&bank.st, sizeof(StreamState) * (size_t)n));
    CK(cudaMalloc(&bank.prm, sizeof(StreamParams) * (size_t)n));
    CK(cudaMalloc(&bank.chunk_ms, sizeof(double) * (size_t)n * 2 * chunk_cap));   // per stream: ms[cap] ++ sorted[cap]
    CK(cudaMalloc(&bank.events, sizeof(EventRec) * (size_t)bank.max_events));
    CK(cudaMalloc(&bank.ev_count, sizeof(int) * 8));
    CK(cudaMemsetAsync(bank.ev_count, 0, sizeof(int) * 8, stream));
    CK(cudaMalloc(&bank.bk_count, sizeof(int) * SEG_NB));
    CK(cudaMemsetAsync(bank.bk_count, 0, sizeof(int) * SEG_NB, stream));
    CK(cudaMalloc(&bank.bk_list, sizeof(int) * (size_t)SEG_NB * bank.max_events));
#ifdef EWK_K3_TRACE
    if (getenv("EWK_K3_TRACE")) {
        CK(cudaMalloc(&bank.k3_trace, sizeof(long long) * K3_TRACE_WORDS * (size_t)queue_grid()));
        CK(cudaMemsetAsync(bank.k3_trace, 0, sizeof(long long) * K3_TRACE_WORDS * (size_t)queue_grid(), stream));
    }
#endif

    bank.NB = bank.P / TICK;
    CK(cudaMalloc(&bank.block_ss, sizeof(double) * (size_t)n * std::max(1, bank.NB)));
    CK(cudaMalloc(&own_results, sizeof(StreamResult) * (size_t)n));
    bank.results = (StreamResult*)own_results;
    if (const char* e = getenv("EWK_K3")) k3_frames = atoi(e) == 2;      // EWK_K3=2: the frame-parallel form (measured, not the default)
    if (!k3_frames) {
        if (use_lm) CK(cudaMalloc(&bank.lm_ws, sizeof(float) * (size_t)queue_grid() * SEG_SMEM_FRAMES * LM_ROW));
    } else {
        // frame-parallel K3: a frame table for the candidates of ONE ewk_tick call (it restarts after every K3 launch)
        const size_t cap = std::min<size_t>((size_t)bank.max_events * SEG_SMEM_FRAMES, ((size_t)2 << 30) / (FROW * sizeof(float)));
        bank.frow_cap = (int)cap;
        CK(cudaMalloc(&bank.frow, sizeof(float) * FROW * cap));
        CK(cudaMalloc(&bank.frame_ev, sizeof(int) * cap));
        CK(cudaMalloc(&bank.ev_done, sizeof(int) * (size_t)bank.max_events));
    }
    std::vector<StreamState> st(n);
    std::memset(st.data(), 0, sizeof(StreamState) * n);
    for (auto& x : st) { x.thr = 0.01; x.last_ev = -1; }                            // wakeword.py:431
    CK(cudaMemcpyAsync(bank.st, st.data(), sizeof(StreamState) * n, cudaMemcpyHostToDevice, stream));
    ewk_stream_params dp;
    ewk_default_stream_params(&dp);
    StreamParams sp;
    std::memcpy(&sp, &dp, sizeof(sp));
    h_prm.assign(n, sp);
    max_post = sp.post_speech_silence;
    CK(cudaMemcpyAsync(bank.prm, h_prm.data(), sizeof(StreamParams) * n, cudaMemcpyHostToDevice, stream));
    std::vector<StreamResult> rs(n);
    for (auto& r : rs) { r.score = std::nanf(""); r.flags = 0; }
    CK(cudaMemcpyAsync(bank.results, rs.data(), sizeof(StreamResult) * n, cudaMemcpyHostToDevice, stream));
    CK(cudaStreamSynchronize(stream));
    h_written.assign(n, 0);
    h_visible_lb.assign(n, 0);
    h_tick.assign(n, 0);
    h_frame_size.assign(n, 0);
    CK(cudaFuncSetAttribute(segment_queue_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            (int)seg_smem_bytes(SEG_SMEM_FRAMES)));
    CK(cudaFuncSetAttribute(segment_queue_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            (int)seg_smem_bytes(SEG_SMEM_FRAMES)));
    CK(cudaFuncSetAttribute(segment_frames_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fq_smem_bytes()));
    CK(cudaFuncSetAttribute(segment_frames_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fq_smem_bytes()));
    CK(cudaFuncSetAttribute(ring_push_bulk_kernel<short>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024));
    CK(cudaFuncSetAttribute(ring_push_bulk_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024));
    {
        const size_t big = sizeof(double) * 3 * (size_t)(std::min(chunk_cap, GATE_SMEM_CHUNKS) + 1) * GATE_WARPS;
        const size_t staged = sizeof(double) * 3 * (size_t)130 * GATE_WARPS + (size_t)2 * TICK * 4 * GATE_WARPS;
        CK(cudaFuncSetAttribute(tick_gate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max(big, staged)));
    }
    return EWK_OK;
}

void ewk_ctx::release_streams() {
#ifdef EWK_K3_TRACE
    if (bank.k3_trace) {                          // the last launch's timeline -> $EWK_K3_TRACE (raw int64 [CTAs][K3_TRACE_WORDS])
        cudaDeviceSynchronize();
        std::vector<long long> h((size_t)K3_TRACE_WORDS * queue_grid());
        cudaMemcpy(h.data(), bank.k3_trace, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
        if (FILE* f = fopen(getenv("EWK_K3_TRACE"), "wb")) { fwrite(h.data(), sizeof(long long), h.size(), f); fclose(f); }
        cudaFree(bank.k3_trace);
    }
#endif
    for (void* p : {bank.ring, (void*)bank.st, (void*)bank.prm, (void*)bank.chunk_ms, (void*)bank.events,
                    (void*)bank.ev_count, (void*)bank.block_ss, (void*)bank.lm_ws, own_results, (void*)bank.frow,
                    (void*)bank.frame_ev, (void*)bank.ev_done, (void*)bank.bk_count, (void*)bank.bk_list})
        if (p) cudaFree(p);
    bank = BankView{};
    own_results = nullptr;
    b_trace.free(); b_read.free(); b_dense.free(); b_keep_rows.free(); b_keep_end.free(); b_g2.free();
    for (auto& p : prof_pairs) { cudaEventDestroy(p.a); cudaEventDestroy(p.b); }
    for (auto e : prof_free) cudaEventDestroy(e);
    prof_pairs.clear(); prof_free.clear();
}

int ewk_ctx::gate_chunks() const {
    int m = 1;
    for (int s = 0; s < bank.n_streams; s++) {
        const int fs = h_frame_size[s] > 0 ? h_frame_size[s] : h_prm[s].frame_size;
        if (fs > 0) m = std::max(m, std::min(bank.chunk_cap, bank.R / fs));
    }
    return m;
}

// Makes the context's stream wait for a K3 launch that is still running on the match stream (overlap mode).
int ewk_ctx::join_match() {
    ewk_ctx* ctx = this;
    if (!match_inflight) return EWK_OK;
    match_inflight = false;
    CK(cudaStreamWaitEvent(stream, ev_match, 0));
    return EWK_OK;
}

// Every stream-bank entry point starts here.  All of them except ewk_push order themselves after an in-flight K3.
static int need_streams(ewk_ctx* ctx, const char* who, bool join = true) {
    if (ctx->bank.n_streams == 0) { ctx->fail("%s: context was created with n_streams = 0", who); return EWK_ERR_STATE; }
    if (join) return ctx->join_match();
    return EWK_OK;
}

extern "C" int ewk_set_overlap(ewk_ctx* ctx, int enable) {
    if (!ctx) return EWK_ERR_ARG;
    int rc = need_streams(ctx, "ewk_set_overlap");
    if (rc) return rc;
    CK(cudaSetDevice(ctx->device));
    if (enable && !ctx->match_stream) {
        CK(cudaStreamCreateWithFlags(&ctx->match_stream, cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&ctx->ev_gate, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&ctx->ev_match, cudaEventDisableTiming));
    }
    ctx->overlap = enable != 0;
    return EWK_OK;
}

extern "C" int ewk_join(ewk_ctx* ctx) {
    if (!ctx) return EWK_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    return ctx->join_match();
}

extern "C" int64_t ewk_launch_count(const ewk_ctx* ctx) { return ctx ? ctx->launches : 0; }

extern "C" int ewk_set_stream_params(ewk_ctx* ctx, int stream, const ewk_stream_params* p) {
    if (!ctx) return EWK_ERR_ARG;
    int rc = need_streams(ctx, "ewk_set_stream_params");
    if (rc) return rc;
    const int n = ctx->bank.n_streams;
    if (!p || stream < -1 || stream >= n) { ctx->fail("ewk_set_stream_params: bad stream %d", stream); return EWK_ERR_ARG; }
    // the reference's constructor checks (wakeword.py:752-763), same messages
    if (!(p->pre_speech_silence > 0)) { ctx->fail("pre_speech_silence must be positive"); return EWK_ERR_ARG; }
    if (!(p->speech_duration_min > 0)) { ctx->fail("speech_duration_min must be positive"); return EWK_ERR_ARG; }
    if (!(p->speech_duration_max > 0)) { ctx->fail("speech_duration_max must be positive"); return EWK_ERR_ARG; }
    if (p->speech_duration_min > p->speech_duration_max) { ctx->fail("speech_duration_min must be <= speech_duration_max"); return EWK_ERR_ARG; }
    if (!(p->post_speech_silence > 0)) { ctx->fail("post_speech_silence must be positive"); return EWK_ERR_ARG; }
    if (p->frame_size != 0 && (p->frame_size < MIN_FRAME_SIZE || p->frame_size > ctx->bank.R)) {
        ctx->fail("frame_size must be 0 or in [%d, %d]", MIN_FRAME_SIZE, ctx->bank.R);
        return EWK_ERR_ARG;
    }
    if (p->template_first < 0 || p->template_count < 0 || p->template_first + p->template_count > ctx->cfg.max_templates) {
        ctx->fail("template range [%d, %d) outside [0, %d)", p->template_first, p->template_first + p->template_count, ctx->cfg.max_templates);
        return EWK_ERR_ARG;
    }
    CK(cudaSetDevice(ctx->device));
    StreamParams sp;
    std::memcpy(&sp, p, sizeof(sp));
    const int a = stream < 0 ? 0 : stream, b = stream < 0 ? n : stream + 1;
    for (int s = a; s < b; s++) ctx->h_prm[s] = sp;
    ctx->max_post = 0.0;
    for (int s = 0; s < n; s++) ctx->max_post = std::max(ctx->max_post, ctx->h_prm[s].post_speech_silence);
    CK(cudaMemcpyAsync(ctx->bank.prm + a, ctx->h_prm.data() + a, sizeof(StreamParams) * (size_t)(b - a),
                       cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return EWK_OK;
}

// K1 on device-resident PCM (caller's buffer or a staging slot): rings <- src, per-block sums, bookkeeping.
int ewk_ctx::land(int stream0, int n_streams, const void* d_src, long long d_stride, long long n, int stage_slot) {
    ewk_ctx* ctx = this;
    BankView& B = bank;
    const size_t esz = B.fmt == 1 ? 2 : 4;
    int with_sums = 0;
    // fused copy + per-block sums when every stream of the push sits on a 0.1 s block boundary
    bool aligned = (n % TICK) == 0 && (B.P % TICK) == 0 && B.NB > 0 && ((size_t)d_src & 15) == 0 && ((d_stride * esz) & 15) == 0;
    for (int s = stream0; aligned && s < stream0 + n_streams; s++) aligned = (h_written[s] % TICK) == 0;
    if (stage_slot >= 0) CK(cudaStreamWaitEvent(stream, ev_ready[stage_slot], 0));
    cudaEvent_t pe = prof_begin(0);
    static const int bulk_env = [] { const char* e = getenv("EWK_PUSH_BULK"); return e ? atoi(e) : 1; }();
    if (aligned && bulk_env && match_inflight && (long long)n_streams * (n / TICK) >= 4LL * sm_count * PUSH_BULK_WARPS) {
        // TMA form while K3 is running on the match stream: persistent CTAs of two warps whose bytes in flight sit in
        // shared-memory stages, small enough (2.5 k registers, 26 KB) to share the SMs with K3's CTAs.  Alone, the
        // register-staged kernel below is faster (47 vs 60 us at 4096 x 1.0 s), so it keeps the sequential case.
        static const int env_stages = [] { const char* e = getenv("EWK_PUSH_STAGES"); return e ? atoi(e) : 0; }();
        static const int env_grid = [] { const char* e = getenv("EWK_PUSH_GRID"); return e ? atoi(e) : 0; }();
        const int stages = env_stages > 0 ? env_stages : (B.fmt == 1 ? 4 : 2);
        const size_t smem = (size_t)PUSH_BULK_WARPS * stages * TICK * esz + sizeof(unsigned long long) * PUSH_BULK_WARPS * stages;
        const int grid = (env_grid > 0 ? env_grid : 2) * sm_count;
        if (B.fmt == 1) ring_push_bulk_kernel<short><<<grid, PUSH_BULK_WARPS * 32, smem, stream>>>(B, stream0, n_streams, (const short*)d_src, d_stride, (int)n, stages);
        else ring_push_bulk_kernel<float><<<grid, PUSH_BULK_WARPS * 32, smem, stream>>>(B, stream0, n_streams, (const float*)d_src, d_stride, (int)n, stages);
        with_sums = 1;
    } else if (aligned) {
        dim3 grid((unsigned)((n / TICK + 3) / 4), (unsigned)n_streams);
        if (B.fmt == 1) ring_push_sums_kernel<short><<<grid, 128, 0, stream>>>(B, stream0, (const short*)d_src, d_stride, (int)n);
        else ring_push_sums_kernel<float><<<grid, 128, 0, stream>>>(B, stream0, (const float*)d_src, d_stride, (int)n);
        with_sums = 1;
    } else {
        const int per = esz == 2 ? 8 : 4;
        dim3 grid((unsigned)std::max<long long>(1, std::min<long long>(64, (n / per + 255) / 256)), (unsigned)n_streams);
        if (B.fmt == 1) ring_push_kernel<short><<<grid, 256, 0, stream>>>(B, stream0, (const short*)d_src, d_stride, (int)n);
        else ring_push_kernel<float><<<grid, 256, 0, stream>>>(B, stream0, (const float*)d_src, d_stride, (int)n);
    }
    prof_end(pe, 0);
    CK(cudaGetLastError());
    launches++;
    if (stage_slot >= 0) {
        CK(cudaEventRecord(ev_free[stage_slot], stream));
        ev_free_valid[stage_slot] = true;
    }
    ring_commit_kernel<<<(n_streams + 255) / 256, 256, 0, stream>>>(B, stream0, n_streams, (int)n, with_sums);
    CK(cudaGetLastError());
    launches++;
    // do the per-block sums of K1 cover everything pushed since the last tick, for every stream?
    if (!with_sums) all_presummed = false;
    else if (stream0 == 0 && n_streams == B.n_streams && pushes_since_tick == 0) all_presummed = true;
    pushes_since_tick++;
    for (int s = stream0; s < stream0 + n_streams; s++) {
        h_written[s] += n;
        if (h_frame_size[s] == 0) h_frame_size[s] = h_prm[s].frame_size > 0 ? h_prm[s].frame_size : (int)n;
    }
    return EWK_OK;
}

// A host push whose H2D copy is in flight on the copy stream lands (K1) only when something needs its samples:
// the next push, a tick that reaches into it, or any read of stream state.  That keeps K1(i+1) — which must wait
// for its copy — out of the way of the kernels of step i on the compute stream, so copy and compute overlap.
int ewk_ctx::flush_pending() {
    if (!pending.valid) return EWK_OK;
    pending.valid = false;
    return land(pending.stream0, pending.n_streams, b_stage2[pending.slot].p, pending.n, pending.n, pending.slot);
}

extern "C" int ewk_push(ewk_ctx* ctx, int stream0, int n_streams, const void* pcm, int64_t n, int64_t stride, int where) {
    if (!ctx) return EWK_ERR_ARG;
    int rc = need_streams(ctx, "ewk_push", false);        // a push does not wait for the matching of the previous ticks
    if (rc) return rc;
    BankView& B = ctx->bank;
    if (!pcm || n_streams < 1 || stream0 < 0 || stream0 + n_streams > B.n_streams || n < 1 || stride < n) {
        ctx->fail("ewk_push: bad arguments (stream0=%d n_streams=%d n=%lld stride=%lld)", stream0, n_streams, (long long)n, (long long)stride);
        return EWK_ERR_ARG;
    }
    if (n > B.P - B.R && n > B.R) { ctx->fail("ewk_push: %lld samples exceed the ring (%d)", (long long)n, B.R); return EWK_ERR_ARG; }
    CK(cudaSetDevice(ctx->device));
    rc = ctx->flush_pending();
    if (rc) return rc;
    // audio-clock streams: the push may not overwrite samples a pending tick still has to see
    for (int s = stream0; s < stream0 + n_streams; s++) {
        if (ctx->h_prm[s].live) continue;
        const long long ahead = ctx->h_written[s] + n - ctx->h_visible_lb[s];
        if (ctx->h_visible_lb[s] >= B.R && ahead > (long long)(B.P - B.R)) {
            ctx->fail("ewk_push: stream %d would hold %lld un-gated samples, more than slack_samples=%d; call ewk_tick first",
                      s, ahead, B.P - B.R);
            return EWK_ERR_STATE;
        }
        // before the first full tick every sample from 0 upward is still needed (the first-full tick forms all chunks
        // from samples): nothing may wrap onto them
        if (ctx->h_visible_lb[s] < B.R && ctx->h_written[s] + n > (long long)B.P) {
            ctx->fail("ewk_push: stream %d would hold %lld samples before its first full tick, more than the physical ring "
                      "(%d); call ewk_tick first", s, ctx->h_written[s] + n, B.P);
            return EWK_ERR_STATE;
        }
    }
    const size_t esz = B.fmt == 1 ? 2 : 4;
    if (where != EWK_HOST) return ctx->land(stream0, n_streams, pcm, stride, n, -1);
    // Host PCM: H2D on a dedicated copy stream into one of two staging buffers.  Contiguous sources
    // ([n_streams][n]) go as ONE linear copy (55 GB/s over PCIe 5 vs 47 GB/s pitched).
    const int b = ctx->stage_idx;
    ctx->stage_idx ^= 1;
    CK(ctx->b_stage2[b].ensure(esz * (size_t)n * n_streams));
    if (ctx->ev_free_valid[b]) CK(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_free[b], 0));
    if (stride == n)
        CK(cudaMemcpyAsync(ctx->b_stage2[b].p, pcm, esz * (size_t)n * n_streams, cudaMemcpyHostToDevice, ctx->copy_stream));
    else
        CK(cudaMemcpy2DAsync(ctx->b_stage2[b].p, (size_t)n * esz, pcm, (size_t)stride * esz, (size_t)n * esz, n_streams,
                             cudaMemcpyHostToDevice, ctx->copy_stream));
    CK(cudaEventRecord(ctx->ev_ready[b], ctx->copy_stream));
    ctx->pending.valid = true; ctx->pending.slot = b; ctx->pending.stream0 = stream0; ctx->pending.n_streams = n_streams;
    ctx->pending.n = n;
    return EWK_OK;
}

// G.711 feed: 8-bit mu-law (law 0) / A-law (law 1) codes, expanded on the device to the 16-bit samples of the standard
// and pushed like PCM16.  Host codes cross PCIe at one byte per sample; the expansion (K0) runs on the copy stream right
// behind the copy, into the same double-buffered staging the PCM path uses, and lands lazily like any host push.
extern "C" int ewk_push_g711(ewk_ctx* ctx, int stream0, int n_streams, const uint8_t* codes, int64_t n, int64_t stride, int where, int law) {
    if (!ctx) return EWK_ERR_ARG;
    int rc = need_streams(ctx, "ewk_push_g711", false);
    if (rc) return rc;
    BankView& B = ctx->bank;
    if (B.fmt != 1) { ctx->fail("ewk_push_g711: the context's rings must be int16 (pcm_format EWK_PCM_I16)"); return EWK_ERR_STATE; }
    if (!codes || n_streams < 1 || stream0 < 0 || stream0 + n_streams > B.n_streams || n < 1 || stride < n || law < 0 || law > 1) {
        ctx->fail("ewk_push_g711: bad arguments (stream0=%d n_streams=%d n=%lld stride=%lld law=%d)", stream0, n_streams,
                  (long long)n, (long long)stride, law);
        return EWK_ERR_ARG;
    }
    if (n > B.P - B.R && n > B.R) { ctx->fail("ewk_push_g711: %lld samples exceed the ring (%d)", (long long)n, B.R); return EWK_ERR_ARG; }
    CK(cudaSetDevice(ctx->device));
    rc = ctx->flush_pending();
    if (rc) return rc;
    for (int s = stream0; s < stream0 + n_streams; s++) {     // same guard as ewk_push
        if (ctx->h_prm[s].live) continue;
        const long long ahead = ctx->h_written[s] + n - ctx->h_visible_lb[s];
        if (ctx->h_visible_lb[s] >= B.R && ahead > (long long)(B.P - B.R)) {
            ctx->fail("ewk_push_g711: stream %d would hold %lld un-gated samples, more than slack_samples=%d; call ewk_tick first",
                      s, ahead, B.P - B.R);
            return EWK_ERR_STATE;
        }
        // before the first full tick every sample from 0 upward is still needed (the first-full tick forms all chunks
        // from samples): nothing may wrap onto them
        if (ctx->h_visible_lb[s] < B.R && ctx->h_written[s] + n > (long long)B.P) {
            ctx->fail("ewk_push_g711: stream %d would hold %lld samples before its first full tick, more than the physical ring "
                      "(%d); call ewk_tick first", s, ctx->h_written[s] + n, B.P);
            return EWK_ERR_STATE;
        }
    }
    const int b = ctx->stage_idx;
    ctx->stage_idx ^= 1;
    CK(ctx->b_stage2[b].ensure(sizeof(short) * (size_t)n * n_streams));
    if (ctx->ev_free_valid[b]) CK(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_free[b], 0));
    const unsigned char* d_codes = codes;
    long long d_stride = stride;
    if (where == EWK_HOST) {
        CK(ctx->b_raw[b].ensure((size_t)n * n_streams));
        if (stride == n)
            CK(cudaMemcpyAsync(ctx->b_raw[b].p, codes, (size_t)n * n_streams, cudaMemcpyHostToDevice, ctx->copy_stream));
        else
            CK(cudaMemcpy2DAsync(ctx->b_raw[b].p, (size_t)n, codes, (size_t)stride, (size_t)n, n_streams, cudaMemcpyHostToDevice, ctx->copy_stream));
        d_codes = (const unsigned char*)ctx->b_raw[b].p;
        d_stride = n;
    } else {
        // device-resident codes were produced on the context's stream: the copy stream starts behind it
        CK(cudaEventRecord(ctx->ev_ready[b], ctx->stream));
        CK(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_ready[b], 0));
    }
    const long long work = (long long)n * n_streams / 16 + 1;
    const int grid = (int)std::min<long long>((work + 255) / 256, 8LL * ctx->sm_count);
    g711_decode_kernel<<<grid, 256, 0, ctx->copy_stream>>>(d_codes, (short*)ctx->b_stage2[b].p, n_streams, (long long)n, d_stride, law);
    CK(cudaGetLastError());
    ctx->launches++;
    CK(cudaEventRecord(ctx->ev_ready[b], ctx->copy_stream));
    ctx->pending.valid = true; ctx->pending.slot = b; ctx->pending.stream0 = stream0; ctx->pending.n_streams = n_streams;
    ctx->pending.n = n;
    return EWK_OK;
}

static int tick_impl(ewk_ctx* ctx, int n_ticks, uint8_t* silent, uint8_t* state, double* thr, double* rms) {
    int rc = need_streams(ctx, "ewk_tick");
    if (rc) return rc;
    if (n_ticks < 1 || n_ticks > 4096) { ctx->fail("ewk_tick: n_ticks must be in [1, 4096]"); return EWK_ERR_ARG; }
    BankView& B = ctx->bank;
    CK(cudaSetDevice(ctx->device));
    if (ctx->pending.valid) {
        // land the in-flight host push only if these ticks reach into its samples
        bool need = false;
        for (int s = ctx->pending.stream0; !need && s < ctx->pending.stream0 + ctx->pending.n_streams; s++) {
            const int fs = ctx->h_frame_size[s] > 0 ? ctx->h_frame_size[s] : ctx->h_prm[s].frame_size;
            if (ctx->h_prm[s].live || fs <= 0) need = true;
            else need = ((ctx->h_tick[s] + n_ticks) * TICK / fs) * fs > ctx->h_written[s];
        }
        if (need) { rc = ctx->flush_pending(); if (rc) return rc; }
    }
    TraceView tr{};
    const bool want = silent || state || thr || rms;
    const size_t cells = (size_t)B.n_streams * n_ticks;
    if (want) {
        CK(ctx->b_trace.ensure(cells * (2 + 16)));
        char* base = (char*)ctx->b_trace.p;
        tr.thr = (double*)base;
        tr.rms = (double*)(base + cells * 8);
        tr.silent = (unsigned char*)(base + cells * 16);
        tr.state = (unsigned char*)(base + cells * 17);
    }
    int smem_chunks = ctx->gate_chunks();
    if (smem_chunks > GATE_SMEM_CHUNKS) {
        ctx->fail("ewk_tick: ring_samples / frame_size = %d chunks; the gate keeps them in shared memory and supports at most %d "
                  "(use a larger frame_size or a shorter ring)", smem_chunks, GATE_SMEM_CHUNKS);
        return EWK_ERR_ARG;
    }
    if (smem_chunks & 1) smem_chunks++;                   // keeps the staging area 16-byte aligned
    // bulk staging of whole ticks pays when chunk arrays are small (frame_size 1600: 100 chunks)
    static const int stage_env = [] { const char* e = getenv("EWK_GATE_STAGE"); return e ? atoi(e) : 1; }();
    const int stage_bytes = (stage_env && smem_chunks <= 128 && !ctx->all_presummed) ? TICK * (B.fmt == 1 ? 2 : 4) : 0;
    if (B.n_pub > 0) {                                    // this call's sequence number and parity buffer at every destination
        ctx->publish_seq++;
        B.pub_seq = (unsigned long long)ctx->publish_seq;
        B.pub_parity = (int)((ctx->publish_seq - 1) & 1);
    }
    if (B.n_pub > 0 && ctx->ev_pub_valid[B.pub_parity])   // K2 rewrites this parity's copy: its sender (two calls ago) is long done
        CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_pub[B.pub_parity], 0));
    cudaEvent_t pe = ctx->prof_begin(1);
    for (int done = 0; done < n_ticks; done += GATE_MAX_TICKS) {
        const int nt = std::min(GATE_MAX_TICKS, n_ticks - done);
        tick_gate_kernel<<<(B.n_streams + GATE_WARPS - 1) / GATE_WARPS, GATE_THREADS,
                           sizeof(double) * 3 * (size_t)smem_chunks * GATE_WARPS + (size_t)2 * stage_bytes * GATE_WARPS,
                           ctx->stream>>>(B, nt, tr, n_ticks, done, smem_chunks, stage_bytes);
        ctx->launches++;
    }
    ctx->prof_end(pe, 1);
    CK(cudaGetLastError());
    const int grid = ctx->queue_grid();
    // Overlap mode: K3 only reads ring samples of the last 3 s before the gated position and writes event / result
    // records, so the next push (K1) may run beside it; the next gate, poll or result read joins it first.  The push
    // guard keeps un-gated samples within the slack, which protects [visible - R, visible); the segments of these ticks
    // start at most 3 s + n_ticks x 0.1 s before it, hence the ring-length condition.
    cudaStream_t ks = ctx->stream;
    // ... plus the tail a cut drops: n_back = segment + n_drop, n_drop ~ (post_speech_silence + 0.05 s) and one tick of slop
    const long long drop_reserve = (long long)std::ceil((ctx->max_post + 0.15) * 16000.0);
    if (ctx->overlap && (long long)B.R >= MAX_SEG + (long long)(n_ticks + 2) * TICK + drop_reserve && !want) {
        CK(cudaEventRecord(ctx->ev_gate, ctx->stream));
        CK(cudaStreamWaitEvent(ctx->match_stream, ctx->ev_gate, 0));
        ks = ctx->match_stream;
    }
    ctx->last_match_stream = ks;
    pe = ctx->prof_begin(2, ks);
    if (ctx->k3_frames) {
        auto k3f = ctx->cfg.preemphasis != 0.f ? segment_frames_kernel<true> : segment_frames_kernel<false>;
        k3f<<<grid, SEG_THREADS, fq_smem_bytes(), ks>>>(ctx->d_tables, B, ctx->d_tmpl, ctx->cfg.max_templates);
    } else {
        auto k3q = ctx->cfg.preemphasis != 0.f ? segment_queue_kernel<true> : segment_queue_kernel<false>;
        k3q<<<grid, SEG_THREADS, seg_smem_bytes(SEG_SMEM_FRAMES), ks>>>(ctx->d_tables, B, ctx->d_tmpl, ctx->cfg.max_templates);
    }
    ctx->prof_end(pe, 2, ks);
    CK(cudaGetLastError());
    if (ks != ctx->stream) {
        CK(cudaEventRecord(ctx->ev_match, ks));
        ctx->match_inflight = true;
    }
    ctx->launches += 1;
    if (B.n_pub > 0) {
        // the sender: off everybody's critical path on its own stream, behind K3's snapshot of the records
        CK(cudaEventRecord(ctx->ev_k3, ks));
        CK(cudaStreamWaitEvent(ctx->pub_stream, ctx->ev_k3, 0));
        cudaEvent_t pp = ctx->prof_begin(6, ctx->pub_stream);
        publish_records_kernel<<<B.n_pub, 256, 0, ctx->pub_stream>>>(B);
        ctx->prof_end(pp, 6, ctx->pub_stream);
        CK(cudaGetLastError());
        ctx->launches += 1;
        CK(cudaEventRecord(ctx->ev_pub[B.pub_parity], ctx->pub_stream));
        ctx->ev_pub_valid[B.pub_parity] = true;
        ctx->launches += 1;
    }
    ctx->pushes_since_tick = 0;
    // host mirrors (audio clock): V after these ticks, given what has been pushed
    for (int s = 0; s < B.n_streams; s++) {
        ctx->h_tick[s] += n_ticks;
        const StreamParams& p = ctx->h_prm[s];
        if (p.live) { ctx->h_visible_lb[s] = ctx->h_written[s]; continue; }
        const int fs = ctx->h_frame_size[s];
        if (fs > 0) {
            const long long v = std::min((ctx->h_tick[s] * TICK / fs) * fs, (ctx->h_written[s] / fs) * fs);
            ctx->h_visible_lb[s] = std::max(ctx->h_visible_lb[s], v);
        }
    }
    if (want) {
        if (thr) CK(cudaMemcpyAsync(thr, tr.thr, cells * 8, cudaMemcpyDeviceToHost, ctx->stream));
        if (rms) CK(cudaMemcpyAsync(rms, tr.rms, cells * 8, cudaMemcpyDeviceToHost, ctx->stream));
        if (silent) CK(cudaMemcpyAsync(silent, tr.silent, cells, cudaMemcpyDeviceToHost, ctx->stream));
        if (state) CK(cudaMemcpyAsync(state, tr.state, cells, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    return EWK_OK;
}

extern "C" int ewk_tick(ewk_ctx* ctx, int n_ticks) {
    if (!ctx) return EWK_ERR_ARG;
    return tick_impl(ctx, n_ticks, nullptr, nullptr, nullptr, nullptr);
}

extern "C" int ewk_tick_trace(ewk_ctx* ctx, int n_ticks, uint8_t* silent, uint8_t* state, double* thr, double* rms) {
    if (!ctx) return EWK_ERR_ARG;
    return tick_impl(ctx, n_ticks, silent, state, thr, rms);
}

extern "C" int ewk_poll(ewk_ctx* ctx, ewk_event* out, int cap, int* dropped) {
    if (!ctx) return EWK_ERR_ARG;
    int rc = need_streams(ctx, "ewk_poll");
    if (rc) return rc;
    if (cap < 0 || (cap > 0 && !out)) { ctx->fail("ewk_poll: bad output buffer"); return EWK_ERR_ARG; }
    BankView& B = ctx->bank;
    CK(cudaSetDevice(ctx->device));
    int cnt[2] = {0, 0};
    CK(cudaMemcpyAsync(cnt, B.ev_count, sizeof(cnt), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    int n = std::min(cnt[0], B.max_events);
    if (dropped) *dropped = cnt[1];
    if (n > cap) { ctx->fail("ewk_poll: %d events pending but cap is %d", n, cap); return EWK_ERR_ARG; }
    if (n > 0) {
        CK(cudaMemcpyAsync(out, B.events, sizeof(EventRec) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    }
    CK(cudaMemsetAsync(B.ev_count, 0, sizeof(int) * 7, ctx->stream));     // count, dropped, K3 work counter, scored watermark, frame table
    CK(cudaStreamSynchronize(ctx->stream));
    std::sort(out, out + n, [](const ewk_event& a, const ewk_event& b) {
        if (a.tick != b.tick) return a.tick < b.tick;
        if (a.stream != b.stream) return a.stream < b.stream;
        return a.kind < b.kind;
    });
    return n;
}

extern "C" int ewk_stream_status_get(ewk_ctx* ctx, int stream, ewk_stream_status* out) {
    if (!ctx) return EWK_ERR_ARG;
    int rc = need_streams(ctx, "ewk_stream_status_get");
    if (rc) return rc;
    cudaSetDevice(ctx->device);
    rc = ctx->flush_pending();
    if (rc) return rc;
    if (!out || stream < 0 || stream >= ctx->bank.n_streams) { ctx->fail("ewk_stream_status_get: bad stream %d", stream); return EWK_ERR_ARG; }
    CK(cudaSetDevice(ctx->device));
    StreamState st;
    CK(cudaMemcpyAsync(&st, ctx->bank.st + stream, sizeof(st), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    out->written = st.written; out->visible = st.visible; out->tick = st.tick;
    out->silence_threshold = st.thr; out->last_rms = st.last_rms; out->frame_size = st.frame_size;
    out->state = st.state; out->started = st.started; out->is_silent = st.last_silent;
    out->n_timeouts = st.n_timeouts; out->n_events = st.n_events;
    return EWK_OK;
}

// absolute samples [a0, a0+len) of one stream as float32 into `out` (host)
static int read_abs(ewk_ctx* ctx, int stream, long long a0, long long len, float* out) {
    BankView& B = ctx->bank;
    const size_t esz = B.fmt == 1 ? 2 : 4;
    std::vector<char> tmp((size_t)len * esz);
    const char* ring = (const char*)B.ring + (size_t)stream * B.P * esz;
    const long long p0 = ((a0 % B.P) + B.P) % B.P;
    const long long first = std::min<long long>(len, B.P - p0);
    CK(cudaMemcpyAsync(tmp.data(), ring + p0 * esz, (size_t)first * esz, cudaMemcpyDeviceToHost, ctx->stream));
    if (first < len)
        CK(cudaMemcpyAsync(tmp.data() + first * esz, ring, (size_t)(len - first) * esz, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (B.fmt == 1) {
        const short* q = (const short*)tmp.data();
        for (long long i = 0; i < len; i++) out[i] = (float)q[i] * (1.0f / 32768.0f);
    } else std::memcpy(out, tmp.data(), (size_t)len * 4);
    return EWK_OK;
}

extern "C" int ewk_read_last(ewk_ctx* ctx, int stream, int64_t n_samples, float* out) {
    if (!ctx) return EWK_ERR_ARG;
    int rc = need_streams(ctx, "ewk_read_last");
    if (rc) return rc;
    cudaSetDevice(ctx->device);
    rc = ctx->flush_pending();
    if (rc) return rc;
    BankView& B = ctx->bank;
    if (!out || stream < 0 || stream >= B.n_streams || n_samples < 0 || n_samples > B.R) {
        ctx->fail("ewk_read_last: bad arguments"); return EWK_ERR_ARG;
    }
    if (n_samples == 0) return EWK_OK;
    CK(cudaSetDevice(ctx->device));
    StreamState st;
    CK(cudaMemcpyAsync(&st, B.st + stream, sizeof(st), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    const long long V = ctx->h_prm[stream].live ? st.written : st.visible;
    // positions never written still hold the zeros of np.zeros (wakeword.py:428)
    const long long a0 = V - n_samples;
    long long lead = a0 < 0 ? -a0 : 0;
    for (long long i = 0; i < lead; i++) out[i] = 0.f;
    if (n_samples - lead > 0) return read_abs(ctx, stream, a0 + lead, n_samples - lead, out + lead);
    return EWK_OK;
}

extern "C" int ewk_read_segment(ewk_ctx* ctx, int stream, int64_t seg_start, int64_t seg_len, float* out) {
    if (!ctx) return EWK_ERR_ARG;
    int rc = need_streams(ctx, "ewk_read_segment");
    if (rc) return rc;
    cudaSetDevice(ctx->device);
    rc = ctx->flush_pending();
    if (rc) return rc;
    BankView& B = ctx->bank;
    if (!out || stream < 0 || stream >= B.n_streams || seg_len < 1 || seg_len > B.P || seg_start < 0) {
        ctx->fail("ewk_read_segment: bad arguments"); return EWK_ERR_ARG;
    }
    if (seg_start + seg_len > ctx->h_written[stream]) { ctx->fail("ewk_read_segment: segment not pushed yet"); return EWK_ERR_STATE; }
    if (ctx->h_written[stream] - seg_start > B.P) { ctx->fail("ewk_read_segment: segment already overwritten in the ring"); return EWK_ERR_STATE; }
    CK(cudaSetDevice(ctx->device));
    return read_abs(ctx, stream, seg_start, seg_len, out);
}

extern "C" int ewk_prepare_segments(ewk_ctx* ctx, int n_seg, const int32_t* streams, const int64_t* starts, const int64_t* lens,
                                    const int64_t* out_offsets, float* out, int64_t out_len, int where) {
    if (!ctx) return EWK_ERR_ARG;
    int rc = need_streams(ctx, "ewk_prepare_segments");
    if (rc) return rc;
    cudaSetDevice(ctx->device);
    rc = ctx->flush_pending();
    if (rc) return rc;
    if (n_seg == 0) return EWK_OK;
    BankView& B = ctx->bank;
    if (n_seg < 0 || !streams || !starts || !lens || !out_offsets || !out || out_len < 1) {
        ctx->fail("ewk_prepare_segments: bad arguments");
        return EWK_ERR_ARG;
    }
    std::vector<PrepDesc> d(n_seg);
    for (int i = 0; i < n_seg; i++) {
        if (streams[i] < 0 || streams[i] >= B.n_streams || lens[i] < 1 || lens[i] > B.P || starts[i] < 0 ||
            out_offsets[i] < 0 || out_offsets[i] + lens[i] > out_len) {
            ctx->fail("ewk_prepare_segments: segment %d is out of range", i);
            return EWK_ERR_ARG;
        }
        if (starts[i] + lens[i] > ctx->h_written[streams[i]]) { ctx->fail("ewk_prepare_segments: segment %d not pushed yet", i); return EWK_ERR_STATE; }
        if (ctx->h_written[streams[i]] - starts[i] > B.P) { ctx->fail("ewk_prepare_segments: segment %d already overwritten in the ring", i); return EWK_ERR_STATE; }
        d[i] = PrepDesc{streams[i], (int)lens[i], starts[i], out_offsets[i]};
    }
    CK(cudaSetDevice(ctx->device));
    CK(ctx->b_desc.ensure(sizeof(PrepDesc) * (size_t)n_seg));
    CK(cudaMemcpyAsync(ctx->b_desc.p, d.data(), sizeof(PrepDesc) * (size_t)n_seg, cudaMemcpyHostToDevice, ctx->stream));
    float* d_out = out;
    if (where == EWK_HOST) { CK(ctx->b_read.ensure(sizeof(float) * (size_t)out_len)); d_out = (float*)ctx->b_read.p; }
    cudaEvent_t pe = ctx->prof_begin(5);
    segment_prepare_kernel<<<n_seg, 256, 0, ctx->stream>>>(B, (const PrepDesc*)ctx->b_desc.p, d_out);
    ctx->prof_end(pe, 5);
    CK(cudaGetLastError());
    ctx->launches++;
    if (where == EWK_HOST) {
        CK(cudaMemcpyAsync(out, d_out, sizeof(float) * (size_t)out_len, cudaMemcpyDeviceToHost, ctx->stream));
    }
    CK(cudaStreamSynchronize(ctx->stream));
    return EWK_OK;
}

// Launch geometry of K4 for a template set: hops per sub-chunk, threads per CTA and CTAs per SM that fit shared memory.
struct DensePlan { int DH, threads, per_sm; size_t smem; };

static bool dense_plan(const DenseArgs& A0, int n_min, int n_max, int n_re_u, DensePlan& out) {
    const size_t SM_TOTAL = 233472, CTA_MAX = 232448, RESERVED = 1024;     // sm_100: 228 KB per SM, 227 KB per CTA
    auto bytes = [&](int DH, int threads) {
        return dense_smem_bytes(threads / 32, A0.T, DH, DH + n_max + 2, DH + n_max - n_min, n_re_u);
    };
    for (int DH : {64, 32})
        for (int threads : {512, 448}) {
            const size_t b = bytes(DH, threads);
            if (2 * (b + RESERVED) <= SM_TOTAL) { out = {DH, threads, 2, b}; return true; }
        }
    for (int threads : {1024, 512})
        for (int DH : {64, 32}) {
            const size_t b = bytes(DH, threads);
            if (b <= CTA_MAX) { out = {DH, threads, 1, b}; return true; }
        }
    return false;
}

extern "C" int ewk_dense_scores(ewk_ctx* ctx, int64_t hop0, int n_hops, int tmpl_first, int tmpl_count, float* out, int where) {
    if (!ctx) return EWK_ERR_ARG;
    int rc = need_streams(ctx, "ewk_dense_scores");
    if (rc) return rc;
    cudaSetDevice(ctx->device);
    BankView& B = ctx->bank;
    if (ctx->pending.valid) {
        // a host push whose copy is still in flight lands now only if these hops reach into its samples: otherwise the copy
        // of the NEXT block keeps overlapping this call's kernel (offline sweeps: push block j + 1, then score block j)
        bool need = false;
        for (int s = ctx->pending.stream0; !need && s < ctx->pending.stream0 + ctx->pending.n_streams; s++)
            need = ctx->h_written[s] < 160LL * (hop0 + n_hops - 1);
        if (need) { rc = ctx->flush_pending(); if (rc) return rc; }
    }
    if (!out || hop0 < 0 || n_hops < 1 || tmpl_count < 1 || tmpl_count > DENSE_MAX_T || tmpl_first < 0 ||
        tmpl_first + tmpl_count > ctx->cfg.max_templates) {
        ctx->fail("ewk_dense_scores: bad arguments (hop0=%lld n_hops=%d templates [%d, %d), at most %d per call)", (long long)hop0,
                  n_hops, tmpl_first, tmpl_first + tmpl_count, DENSE_MAX_T);
        return EWK_ERR_ARG;
    }
    DenseArgs A{};
    A.hop0 = hop0; A.n_hops = n_hops; A.T = tmpl_count;
    int max_n = 0, min_n = 1 << 30, g_back = 1 << 30, n_re_u = 0;
    for (int k = 0; k < tmpl_count; k++) {
        const TemplateFeat& tf = ctx->h_tmpl[tmpl_first + k];
        if (!tf.valid) { ctx->fail("No reference word set. Call set_reference() first."); return EWK_ERR_NO_TEMPLATE; }
        const int L = (int)tf.n_samples;
        if (tf.n_samples < DENSE_MIN_L || tf.n_samples > DENSE_MAX_L) {
            ctx->fail("ewk_dense_scores: template %d has %lld samples; dense scoring needs %d..%d", tmpl_first + k,
                      (long long)tf.n_samples, DENSE_MIN_L, DENSE_MAX_L);
            return EWK_ERR_ARG;
        }
        DenseTmplDev& t = A.t[k];
        t.L = L; t.n = (L + HOP - 1) / HOP; t.F = 1 + L / HOP; t.t_hi = (L - N_FFT / 2) / HOP; t.r = t.F - 1 - t.t_hi;
        t.slot = tmpl_first + k;
        t.inv_f = 1.0 / (double)t.F;
        max_n = std::max(max_n, t.n); min_n = std::min(min_n, t.n);
        g_back = std::min(g_back, t.n - t.t_hi);
        // right-edge frame e: stream-grid frame (hop - goff) cut at sample 160 hop - delta.  Templates with the same
        // (goff, delta) share it; the group is computed through the window of its shortest member
        const int delta = HOP * t.n - L;
        for (int e = 0; e < t.r; e++) {
            const int goff = t.n - (t.t_hi + 1 + e);
            int u = 0;
            while (u < n_re_u && !(A.re_goff[u] == goff && A.re_delta[u] == delta)) u++;
            if (u == n_re_u) {
                A.re_goff[u] = goff; A.re_delta[u] = delta; A.re_rep[u] = k; A.re_rep_e[u] = e; A.re_nfirst[u] = t.n;
                n_re_u++;
            } else if (t.n < A.re_nfirst[u]) { A.re_rep[u] = k; A.re_rep_e[u] = e; A.re_nfirst[u] = t.n; }
            t.re_u[e] = u;
        }
    }
    DensePlan plan;
    if (!dense_plan(A, min_n, max_n, n_re_u, plan)) {
        ctx->fail("ewk_dense_scores: this template set does not fit shared memory (%d templates, windows of %d..%d hops)",
                  tmpl_count, min_n, max_n);
        return EWK_ERR_ARG;
    }
    A.DH = plan.DH; A.DG = plan.DH + max_n + 2; A.DLE = plan.DH + max_n - min_n; A.n_re_u = n_re_u;
    A.n_min = min_n; A.n_max = max_n; A.g_back = g_back;
    if ((long long)B.P < 160LL * (max_n + plan.DH + 4) + N_FFT) {        // the kernel wraps ring positions once
        ctx->fail("ewk_dense_scores: ring of %d samples is too short for a template of %d hops", B.P, max_n);
        return EWK_ERR_ARG;
    }
    // the audio of every requested window must be resident
    const long long need_hi = 160LL * (hop0 + n_hops - 1);
    const long long need_lo = std::max<long long>(0, 160LL * (hop0 - max_n) - N_FFT / 2);
    for (int s = 0; s < B.n_streams; s++) {
        if (ctx->h_written[s] < need_hi) {
            ctx->fail("ewk_dense_scores: stream %d has %lld samples, hop %lld needs %lld", s, ctx->h_written[s],
                      (long long)(hop0 + n_hops - 1), need_hi);
            return EWK_ERR_STATE;
        }
        if (ctx->h_written[s] - need_lo > B.P) {
            ctx->fail("ewk_dense_scores: stream %d no longer holds sample %lld (ring of %d)", s, need_lo, B.P);
            return EWK_ERR_STATE;
        }
    }
    CK(cudaSetDevice(ctx->device));
    const size_t n_out = (size_t)B.n_streams * n_hops * tmpl_count;
    float* d_out = out;
    if (where == EWK_HOST) { CK(ctx->b_dense.ensure(sizeof(float) * n_out)); d_out = (float*)ctx->b_dense.p; }
    A.out = d_out;
    if (!ctx->b_keep_rows.p) {
        CK(ctx->b_keep_rows.ensure(sizeof(float) * (size_t)B.n_streams * DENSE_KEEP * ROW));
        CK(ctx->b_keep_end.ensure(sizeof(long long) * 2 * (size_t)B.n_streams));
        CK(cudaMemsetAsync(ctx->b_keep_end.p, 0, sizeof(long long) * 2 * (size_t)B.n_streams, ctx->stream));
    }
    CK(ctx->b_g2.ensure(sizeof(float) * (size_t)B.n_streams * DENSE_WAYS * A.DG * N_MFCC));
    A.keep_rows = (float*)ctx->b_keep_rows.p;
    A.keep_end = (long long*)ctx->b_keep_end.p;
    A.g2 = (float*)ctx->b_g2.p;
    auto k4 = ctx->cfg.preemphasis != 0.f ? dense_score_kernel<true> : dense_score_kernel<false>;
    CK(cudaFuncSetAttribute(k4, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem));
    CK(cudaFuncSetAttribute(k4, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    cudaEvent_t pe = ctx->prof_begin(4);
    k4<<<B.n_streams, plan.threads, plan.smem, ctx->stream>>>(ctx->d_tables, B, ctx->d_tmpl, A);
    ctx->prof_end(pe, 4);
    CK(cudaGetLastError());
    ctx->launches++;
    ctx->dense_plan_info[0] = plan.DH; ctx->dense_plan_info[1] = plan.threads; ctx->dense_plan_info[2] = plan.per_sm;
    ctx->dense_plan_info[3] = (int)plan.smem;
    if (where == EWK_HOST) {
        CK(cudaMemcpyAsync(out, d_out, sizeof(float) * n_out, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    return EWK_OK;
}

extern "C" int ewk_stream_results(ewk_ctx* ctx, ewk_stream_result* out) {
    if (!ctx) return EWK_ERR_ARG;
    int rc = need_streams(ctx, "ewk_stream_results");
    if (rc) return rc;
    if (!out) { ctx->fail("ewk_stream_results: null output"); return EWK_ERR_ARG; }
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync(out, ctx->bank.results, sizeof(StreamResult) * (size_t)ctx->bank.n_streams, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return EWK_OK;
}

extern "C" int ewk_results_device_ptr(ewk_ctx* ctx, void** out) {
    if (!ctx || !out) return EWK_ERR_ARG;
    int rc = need_streams(ctx, "ewk_results_device_ptr");
    if (rc) return rc;
    *out = ctx->bank.results;
    return EWK_OK;
}

extern "C" int ewk_set_results_buffer(ewk_ctx* ctx, void* device_ptr) {
    if (!ctx) return EWK_ERR_ARG;
    int rc = need_streams(ctx, "ewk_set_results_buffer");
    if (rc) return rc;
    if ((size_t)device_ptr & 7) { ctx->fail("ewk_set_results_buffer: the buffer must be 8-byte aligned (records are stored as one 64-bit word)"); return EWK_ERR_ARG; }
    CK(cudaSetDevice(ctx->device));
    StreamResult* dst = device_ptr ? (StreamResult*)device_ptr : (StreamResult*)ctx->own_results;
    if (dst != ctx->bank.results) {
        CK(cudaMemcpyAsync(dst, ctx->bank.results, sizeof(StreamResult) * (size_t)ctx->bank.n_streams, cudaMemcpyDeviceToDevice, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        ctx->bank.results = dst;
    }
    return EWK_OK;
}

extern "C" int ewk_set_results_peers(ewk_ctx* ctx, void* const* bases, int n_bases, int64_t stride_records, int64_t offset_records,
                                     void* const* signals, int slot) {
    if (!ctx) return EWK_ERR_ARG;
    int rc = need_streams(ctx, "ewk_set_results_peers");
    if (rc) return rc;
    BankView& B = ctx->bank;
    if (n_bases < 0 || n_bases > MAX_PUB || (n_bases > 0 && !bases)) {
        ctx->fail("ewk_set_results_peers: n_bases must be in [0, %d]", MAX_PUB);
        return EWK_ERR_ARG;
    }
    if (n_bases > 0 && (offset_records < 0 || stride_records < offset_records + B.n_streams)) {
        ctx->fail("ewk_set_results_peers: stride %lld cannot hold %d records at offset %lld",
                  (long long)stride_records, B.n_streams, (long long)offset_records);
        return EWK_ERR_ARG;
    }
    if (n_bases > 0 && signals && (slot < 0 || slot >= n_bases)) {
        ctx->fail("ewk_set_results_peers: slot %d outside [0, %d)", slot, n_bases);
        return EWK_ERR_ARG;
    }
    for (int p = 0; p < n_bases; p++) {
        if (!bases[p] || ((size_t)bases[p] & 7)) { ctx->fail("ewk_set_results_peers: destination %d is null or not 8-byte aligned", p); return EWK_ERR_ARG; }
        if (signals && (!signals[p] || ((size_t)signals[p] & 7))) { ctx->fail("ewk_set_results_peers: signal row %d is null or not 8-byte aligned", p); return EWK_ERR_ARG; }
    }
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));                // kernels in flight keep the destinations they were launched with
    for (int p = 0; p < MAX_PUB; p++) {
        B.pub[p] = p < n_bases ? (StreamResult*)bases[p] : nullptr;
        B.pub_sig[p] = (p < n_bases && signals) ? (unsigned long long*)signals[p] : nullptr;
    }
    B.n_pub = n_bases;
    B.pub_stride = n_bases ? stride_records : 0;
    B.pub_off = n_bases ? offset_records : 0;
    B.pub_slot = n_bases && signals ? slot : 0;
    B.pub_seq = 0;
    B.pub_parity = 0;
    ctx->publish_seq = 0;
    if (ctx->pub_stream) CK(cudaStreamSynchronize(ctx->pub_stream));
    ctx->ev_pub_valid[0] = ctx->ev_pub_valid[1] = false;
    if (n_bases && !ctx->pub_stream) {
        // highest priority: the sender is a handful of CTAs that should start the moment K3 ends, not queue behind the
        // full grids of the next push
        int prio_lo = 0, prio_hi = 0;
        CK(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
        CK(cudaStreamCreateWithPriority(&ctx->pub_stream, cudaStreamNonBlocking, prio_hi));
        CK(cudaEventCreateWithFlags(&ctx->ev_k3, cudaEventDisableTiming));
        for (int i = 0; i < 2; i++) CK(cudaEventCreateWithFlags(&ctx->ev_pub[i], cudaEventDisableTiming));
        CK(cudaMalloc(&ctx->d_pub_snap, sizeof(StreamResult) * 2 * (size_t)B.n_streams));
    }
    B.pub_snap = (StreamResult*)ctx->d_pub_snap;
    if (n_bases && signals && !ctx->d_wait_flag) {
        CK(cudaMalloc(&ctx->d_wait_flag, sizeof(int)));
        CK(cudaMemsetAsync(ctx->d_wait_flag, 0, sizeof(int), ctx->stream));
    }
    return EWK_OK;
}

extern "C" int ewk_publish_parity(const ewk_ctx* ctx) { return ctx && ctx->publish_seq > 0 ? (int)((ctx->publish_seq - 1) & 1) : -1; }
extern "C" int64_t ewk_publish_seq(const ewk_ctx* ctx) { return ctx ? ctx->publish_seq : 0; }

extern "C" int ewk_wait_published(ewk_ctx* ctx, int n_slots, int64_t seq, int timeout_ms) {
    if (!ctx) return EWK_ERR_ARG;
    int rc = need_streams(ctx, "ewk_wait_published", false);
    if (rc) return rc;
    BankView& B = ctx->bank;
    if (!B.n_pub || !B.pub_sig[0]) { ctx->fail("ewk_wait_published: no signal rows installed (ewk_set_results_peers)"); return EWK_ERR_STATE; }
    if (n_slots < 1 || n_slots > B.n_pub || seq < 1 || timeout_ms < 1) { ctx->fail("ewk_wait_published: bad arguments"); return EWK_ERR_ARG; }
    CK(cudaSetDevice(ctx->device));
    // this rank's own sender (side stream, same GPU) first: the wait kernel then only spins on other GPUs' signals
    if (ctx->ev_pub_valid[(seq - 1) & 1]) CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_pub[(seq - 1) & 1], 0));
    const unsigned long long* row = B.pub_sig[B.pub_slot] + (size_t)((seq - 1) & 1) * MAX_PUB;
    peer_wait_kernel<<<1, 32, 0, ctx->stream>>>(row, n_slots, (unsigned long long)seq, (unsigned long long)timeout_ms * 1000000ULL,
                                                ctx->d_wait_flag);
    CK(cudaGetLastError());
    ctx->launches++;
    return EWK_OK;
}

extern "C" int ewk_published_seq(ewk_ctx* ctx, int parity, uint64_t* out, int n_slots) {
    if (!ctx) return EWK_ERR_ARG;
    int rc = need_streams(ctx, "ewk_published_seq", false);
    if (rc) return rc;
    BankView& B = ctx->bank;
    if (!B.n_pub || !B.pub_sig[0]) { ctx->fail("ewk_published_seq: no signal rows installed (ewk_set_results_peers)"); return EWK_ERR_STATE; }
    if (!out || n_slots < 1 || n_slots > MAX_PUB || parity < 0 || parity > 1) { ctx->fail("ewk_published_seq: bad arguments"); return EWK_ERR_ARG; }
    CK(cudaSetDevice(ctx->device));
    int flag = 0;
    CK(cudaMemcpyAsync(out, B.pub_sig[B.pub_slot] + (size_t)parity * MAX_PUB, sizeof(uint64_t) * (size_t)n_slots, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(&flag, ctx->d_wait_flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemsetAsync(ctx->d_wait_flag, 0, sizeof(int), ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return flag ? 1 : 0;
}

extern "C" int ewk_match_stream(ewk_ctx* ctx, void** out) {
    if (!ctx || !out) return EWK_ERR_ARG;
    if (ctx->bank.n_pub > 0 && ctx->pub_stream) { *out = (void*)ctx->pub_stream; return EWK_OK; }   // where the call's records are sent
    *out = (void*)(ctx->last_match_stream ? ctx->last_match_stream : ctx->stream);
    return EWK_OK;
}

extern "C" int ewk_host_alloc(void** out, int64_t bytes) {
    if (!out || bytes < 1) return EWK_ERR_ARG;
    return cudaHostAlloc(out, (size_t)bytes, cudaHostAllocDefault) == cudaSuccess ? EWK_OK : EWK_ERR_NOMEM;
}

extern "C" int ewk_host_free(void* p) {
    if (!p) return EWK_ERR_ARG;
    return cudaFreeHost(p) == cudaSuccess ? EWK_OK : EWK_ERR_CUDA;
}

// ==========================================================================================
// per-kernel event timing
// ==========================================================================================
cudaEvent_t ewk_ctx::prof_begin(int, cudaStream_t on) {
    if (!prof_on) return nullptr;
    cudaEvent_t e;
    if (!prof_free.empty()) { e = prof_free.back(); prof_free.pop_back(); }
    else if (cudaEventCreate(&e) != cudaSuccess) return nullptr;
    cudaEventRecord(e, on ? on : stream);
    return e;
}

void ewk_ctx::prof_end(cudaEvent_t a, int cls, cudaStream_t on) {
    if (!a) return;
    cudaEvent_t b;
    if (!prof_free.empty()) { b = prof_free.back(); prof_free.pop_back(); }
    else if (cudaEventCreate(&b) != cudaSuccess) { prof_free.push_back(a); return; }
    cudaEventRecord(b, on ? on : stream);
    prof_pairs.push_back({a, b, cls});
}

int ewk_ctx::prof_collect() {
    ewk_ctx* ctx = this;
    int jr = join_match();
    if (jr) return jr;
    CK(cudaStreamSynchronize(stream));
    if (pub_stream) CK(cudaStreamSynchronize(pub_stream));
    for (auto& p : prof_pairs) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) { prof_ms[p.cls] += ms; prof_n[p.cls]++; }
        prof_free.push_back(p.a);
        prof_free.push_back(p.b);
    }
    prof_pairs.clear();
    return EWK_OK;
}

extern "C" int ewk_profile(ewk_ctx* ctx, int enable) {
    if (!ctx) return EWK_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    int rc = ctx->prof_collect();
    if (rc) return rc;
    for (int i = 0; i < 8; i++) { ctx->prof_ms[i] = 0; ctx->prof_n[i] = 0; }
    ctx->prof_on = enable != 0;
    return EWK_OK;
}

extern "C" int ewk_profile_read(ewk_ctx* ctx, double* ms, int64_t* launches) {
    if (!ctx || !ms || !launches) return EWK_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    int rc = ctx->prof_collect();
    if (rc) return rc;
    for (int i = 0; i < 8; i++) { ms[i] = ctx->prof_ms[i]; launches[i] = ctx->prof_n[i]; ctx->prof_ms[i] = 0; ctx->prof_n[i] = 0; }
    return EWK_OK;
}

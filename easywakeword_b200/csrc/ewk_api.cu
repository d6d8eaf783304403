// libewk: C ABI (include/ewk.h) over the sm_100a kernels.  No torch types, no CPU fallback.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
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
    ctx->stream = s ? (cudaStream_t)s : ctx->own_stream;
    return EWK_OK;
}

extern "C" int ewk_synchronize(ewk_ctx* ctx) {
    if (!ctx) return EWK_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    return EWK_OK;
}

int ewk_ctx::init() {
    ewk_ctx* ctx = this;
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    sm_count = prop.multiProcessorCount;
    if (prop.major < 10) {
        fail("ewk_create: device %d is sm_%d%d; libewk is built for sm_100a only", device, prop.major, prop.minor);
        return EWK_ERR_CUDA;
    }
    CK(cudaStreamCreateWithFlags(&own_stream, cudaStreamNonBlocking));
    stream = own_stream;
    DeviceTables* h = new DeviceTables();
    int nnz = build_tables(*h);
    if (nnz < 0) { delete h; fail("mel table overflow"); return EWK_ERR_STATE; }
    CK(cudaMalloc(&d_tables, sizeof(DeviceTables)));
    CK(cudaMemcpy(d_tables, h, sizeof(DeviceTables), cudaMemcpyHostToDevice));
    delete h;
    h_tmpl.assign(cfg.max_templates, TemplateFeat{});
    CK(cudaMalloc(&d_tmpl, sizeof(TemplateFeat) * cfg.max_templates));
    CK(cudaMemset(d_tmpl, 0, sizeof(TemplateFeat) * cfg.max_templates));
    CK(cudaFuncSetAttribute(segment_mfcc_match_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            (int)seg_smem_bytes(SEG_SMEM_FRAMES)));
    int rc = init_streams();
    if (rc != EWK_OK) return rc;
    return EWK_OK;
}

void ewk_ctx::release() {
    cudaSetDevice(device);
    if (own_stream) cudaStreamSynchronize(own_stream);
    release_streams();
    for (DevBuf* b : {&b_pcm, &b_desc, &b_ws, &b_feat, &b_scores, &b_matched, &b_frames, &b_off}) b->free();
    if (d_tables) cudaFree(d_tables);
    if (d_tmpl) cudaFree(d_tmpl);
    if (own_stream) cudaStreamDestroy(own_stream);
    d_tables = nullptr; d_tmpl = nullptr; own_stream = nullptr;
}

// ------------------------------------------------------------------------------------------
// Launch K3 over `n_seg` descriptors already on the device.
int ewk_ctx::launch_segments(const SegDesc* d_segs, int n_seg, int max_frames, long long spill_frames,
                             int n_tmpl, int tmpl_first, float threshold, float* d_feat, float* d_frames,
                             float* d_scores, unsigned char* d_matched) {
    ewk_ctx* ctx = this;
    const int cap = std::min(max_frames, SEG_SMEM_FRAMES);
    float* ws = nullptr;
    if (spill_frames > 0) {
        CK(b_ws.ensure(sizeof(float) * (size_t)spill_frames * (LM_STRIDE + N_MFCC)));
        ws = (float*)b_ws.p;
    }
    segment_mfcc_match_kernel<<<n_seg, SEG_THREADS, seg_smem_bytes(cap), stream>>>(
        d_tables, d_segs, cap, ws, d_tmpl, n_tmpl, tmpl_first, threshold, d_feat, d_frames, d_scores, d_matched);
    CK(cudaGetLastError());
    return EWK_OK;
}

extern "C" int ewk_extract_mfcc(ewk_ctx* ctx, const float* pcm, int64_t n, int where, float* mean20, float* std20,
                                float* frames, int64_t frames_cap) {
    if (!ctx) return EWK_ERR_ARG;
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
    CK(ctx->b_desc.ensure(sizeof(SegDesc)));
    CK(cudaMemcpyAsync(ctx->b_desc.p, &sd, sizeof(sd), cudaMemcpyHostToDevice, ctx->stream));
    CK(ctx->b_feat.ensure(sizeof(float) * FEAT));
    float* d_frames = nullptr;
    if (frames) { CK(ctx->b_frames.ensure(sizeof(float) * (size_t)F * N_MFCC)); d_frames = (float*)ctx->b_frames.p; }
    int rc = ctx->launch_segments((const SegDesc*)ctx->b_desc.p, 1, (int)F, F > SEG_SMEM_FRAMES ? F : 0, 0, 0, 0.f,
                                  (float*)ctx->b_feat.p, d_frames, nullptr, nullptr);
    if (rc != EWK_OK) return rc;
    float feat[FEAT];
    CK(cudaMemcpyAsync(feat, ctx->b_feat.p, sizeof(feat), cudaMemcpyDeviceToHost, ctx->stream));
    if (frames) CK(cudaMemcpyAsync(frames, d_frames, sizeof(float) * (size_t)F * N_MFCC, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    std::memcpy(mean20, feat, sizeof(float) * N_MFCC);
    std::memcpy(std20, feat + N_MFCC, sizeof(float) * N_MFCC);
    return EWK_OK;
}

static int check_slot(ewk_ctx* ctx, int slot, const char* who) {
    if (slot < 0 || slot >= ctx->cfg.max_templates) {
        ctx->fail("%s: template slot %d out of range [0, %d)", who, slot, ctx->cfg.max_templates);
        return EWK_ERR_ARG;
    }
    return EWK_OK;
}

extern "C" int ewk_set_template_features(ewk_ctx* ctx, int slot, const float* mean20, const float* std20, int64_t n_samples) {
    if (!ctx) return EWK_ERR_ARG;
    int rc = check_slot(ctx, slot, "ewk_set_template_features");
    if (rc) return rc;
    if (!mean20 || !std20 || n_samples < 1) { ctx->fail("ewk_set_template_features: null features or n_samples < 1"); return EWK_ERR_ARG; }
    CK(cudaSetDevice(ctx->device));
    TemplateFeat tf{};
    std::memcpy(tf.mean, mean20, sizeof(tf.mean));
    std::memcpy(tf.std, std20, sizeof(tf.std));
    tf.n_samples = n_samples;
    tf.valid = 1;
    ctx->h_tmpl[slot] = tf;
    CK(cudaMemcpyAsync(ctx->d_tmpl + slot, &ctx->h_tmpl[slot], sizeof(tf), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return EWK_OK;
}

extern "C" int ewk_set_template(ewk_ctx* ctx, int slot, const float* pcm, int64_t n) {
    if (!ctx) return EWK_ERR_ARG;
    int rc = check_slot(ctx, slot, "ewk_set_template");
    if (rc) return rc;
    float mean[N_MFCC], sd[N_MFCC];
    rc = ewk_extract_mfcc(ctx, pcm, n, EWK_HOST, mean, sd, nullptr, 0);
    if (rc) return rc;
    return ewk_set_template_features(ctx, slot, mean, sd, n);
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
    int64_t extent = 0, spill = 0;
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
    rc = ctx->launch_segments((const SegDesc*)ctx->b_desc.p, n_seg, max_frames, spill, 1, slot, threshold,
                              (float*)ctx->b_feat.p, nullptr, (float*)ctx->b_scores.p, (unsigned char*)ctx->b_matched.p);
    if (rc != EWK_OK) return rc;
    CK(cudaMemcpyAsync(scores, ctx->b_scores.p, sizeof(float) * (size_t)n_seg, cudaMemcpyDeviceToHost, ctx->stream));
    if (matched) CK(cudaMemcpyAsync(matched, ctx->b_matched.p, (size_t)n_seg, cudaMemcpyDeviceToHost, ctx->stream));
    if (features) CK(cudaMemcpyAsync(features, ctx->b_feat.p, sizeof(float) * FEAT * (size_t)n_seg, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return EWK_OK;
}

/* libewk — C ABI of the B200-native EasyWakeWord hot path (level-1 energy gate + level-2 MFCC/cosine
 * matcher, batched over independent 16 kHz audio streams).
 *
 * The reference (raymondclowe/EasyWakeWord) has no FFI of its own: its seam is two duck-typed Python
 * attributes of WakeWord, `_sound_buffer` and `_matcher` (easywakeword/wakeword.py:989-997; used at
 * 1004, 1055, 1066, 1105, 1121, 1225).  Each entry point below cites the reference method it replaces.
 * easywakeword_b200/_lib.py is the ctypes binding; INTEGRATION.md shows the stub a reference
 * maintainer would add.
 *
 * Conventions: every call returns 0 (EWK_OK) or a negative ewk_status; the message is read with
 * ewk_last_error().  The caller owns all host buffers; the library owns device memory; no callbacks
 * cross the ABI.  One context per GPU; one consumer thread per context.  There is no CPU fallback:
 * without a CUDA device ewk_create fails with EWK_ERR_CUDA.
 */
#ifndef EWK_H
#define EWK_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EWK_ABI_VERSION 1

typedef struct ewk_ctx ewk_ctx;

typedef enum ewk_status {
    EWK_OK = 0,
    EWK_ERR_ARG = -1,         /* bad argument                                    -> ValueError   */
    EWK_ERR_NO_TEMPLATE = -2, /* wakeword.py:608-609 "No reference word set..."  -> ValueError   */
    EWK_ERR_CUDA = -3,        /* CUDA runtime failure / no device                -> RuntimeError */
    EWK_ERR_STATE = -4,       /* call made in the wrong state                    -> RuntimeError */
    EWK_ERR_NOMEM = -5
} ewk_status;

typedef enum ewk_pcm_format { EWK_PCM_F32 = 0, EWK_PCM_I16 = 1 } ewk_pcm_format;
typedef enum ewk_mem { EWK_HOST = 0, EWK_DEVICE = 1 } ewk_mem;

enum { EWK_N_MFCC = 20, EWK_N_MELS = 128, EWK_N_FFT = 512, EWK_HOP = 160, EWK_SAMPLE_RATE = 16000,
       EWK_TICK_SAMPLES = 1600 };

typedef struct ewk_config {
    int32_t n_streams;      /* independent audio streams resident on this GPU (0: matcher only)       */
    int32_t ring_samples;   /* SoundBuffer.buffer_length = seconds * 16000       wakeword.py:426-428  */
    int32_t slack_samples;  /* extra physical ring so that one ewk_push may run ahead of ewk_tick      */
    int32_t pcm_format;     /* ewk_pcm_format of the device rings                                     */
    int32_t max_templates;  /* template slots (reference: one WordMatcher holds one template)         */
    int32_t max_events;     /* capacity of the device event queue between two ewk_poll calls          */
} ewk_config;

/* ---- library ------------------------------------------------------------------------------ */
int ewk_abi_version(void);
/* message of the last failing call on `ctx` (or, with ctx == NULL, of the last failing ewk_create
 * on this thread).  Never NULL. */
const char* ewk_last_error(const ewk_ctx* ctx);
/* Host-only (no GPU needed): copies a read-only table the kernels use into out[cap] and returns its
 * length: 0 hann[512], 1 mel filterbank dense [128*257], 2 dct [20*128] (row k, column b). */
int ewk_host_table(int which, float* out, int cap);

/* ---- context ------------------------------------------------------------------------------ */
/* WakeWord._initialize_audio (wakeword.py:989-1000) for a whole bank of streams on CUDA `device`. */
int ewk_create(int device, const ewk_config* cfg, ewk_ctx** out);
int ewk_destroy(ewk_ctx* ctx);
/* Run all kernels of `ctx` on an existing cudaStream_t (e.g. torch's current stream); NULL restores
 * the context's own stream. */
int ewk_set_cuda_stream(ewk_ctx* ctx, void* cuda_stream);
int ewk_synchronize(ewk_ctx* ctx);

/* ---- level 2: WordMatcher (wakeword.py:520-639) -------------------------------------------- */
/* WordMatcher.extract_mfcc (wakeword.py:544-567): n float32 samples -> mean[20], std[20] over the
 * 1 + n/160 MFCC frames; when frames != NULL also the frame matrix, row-major [n_frames][20]
 * (frames_cap = capacity in frames).  pcm is a host or device pointer according to `where`. */
int ewk_extract_mfcc(ewk_ctx* ctx, const float* pcm, int64_t n, int where, float* mean20, float* std20,
                     float* frames, int64_t frames_cap);
/* WordMatcher.set_reference (wakeword.py:569-578) into template slot `slot`. */
int ewk_set_template(ewk_ctx* ctx, int slot, const float* pcm, int64_t n);
/* Install precomputed features (mean[20], std[20]) for a template of n_samples samples. */
int ewk_set_template_features(ewk_ctx* ctx, int slot, const float* mean20, const float* std20, int64_t n_samples);
int ewk_get_template(ewk_ctx* ctx, int slot, float* mean20, float* std20, int64_t* n_samples);
int ewk_clear_template(ewk_ctx* ctx, int slot);
/* WordMatcher.calculate_similarity / matches (wakeword.py:591-639) for n_seg segments of one PCM
 * buffer (segment i = pcm[offsets[i] : offsets[i] + lens[i]]) against template `slot`:
 * scores[i] in [0, 100] or NaN; matched[i] = scores[i] >= threshold (NaN -> 0).  features may be NULL;
 * otherwise [n_seg][40] = mean, std per segment.  EWK_ERR_NO_TEMPLATE if the slot is empty. */
int ewk_similarity_batch(ewk_ctx* ctx, int slot, const void* pcm, int pcm_format, int where,
                         const int64_t* offsets, const int64_t* lens, int n_seg, float threshold,
                         float* scores, uint8_t* matched, float* features);

#ifdef __cplusplus
}
#endif
#endif /* EWK_H */

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

#define EWK_ABI_VERSION 2

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
    /* --- ABI 2: front-end parameters the reference hard-codes (wakeword.py:561-563; exposing them is on its own
     * wish list, LEARNINGS.md:87, README-CODE-ALIGNMENT.md:84-107).  0 selects the reference's value. */
    float preemphasis;      /* coefficient a of y[n] = x[n] - a x[n-1], applied to every segment / dense window
                               handed to the MFCC front-end, before centring (librosa.effects.preemphasis incl.
                               its zi = 2 x[0] - x[1] initial state).  The reference applies none: 0 = identity
                               is the parity default and leaves every result bit-identical                    */
    int32_t n_mfcc;         /* MFCC coefficients kept, 1..20; 0 = 20                      wakeword.py:562     */
    int32_t reserved0;
    int32_t reserved1;
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

/* ---- level 1: SoundBuffer (wakeword.py:405-517) + _detect_word timing (wakeword.py:1036-1159) --- */
typedef struct ewk_stream_params {
    float similarity_threshold;     /* WakeWord(similarity_threshold=75.0)              wakeword.py:676  */
    int32_t frame_size;             /* PortAudio callback length; 0: length of the first push  :457-458 */
    double pre_speech_silence;      /* 0.8                                                :38, 677       */
    double speech_duration_min;     /* auto from the template WAV (fallback 0.3)          :39, 678       */
    double speech_duration_max;     /* 2 x min (fallback 2.0)                             :40, 679       */
    double post_speech_silence;     /* 0.4                                                :41, 680       */
    double timeout;                 /* seconds before _detect_word raises and the listen loop re-enters
                                       it (:1061-1062, 1205-1211); <= 0: never                          */
    double min_threshold;           /* SoundBuffer.MIN_THRESHOLD = 0.005                  :409           */
    int32_t template_first;         /* the stream is scored against template slots                      */
    int32_t template_count;         /*   [template_first, template_first + template_count), best wins   */
    int32_t live;                   /* 1: a tick sees every pushed sample (SoundBuffer facade driven by
                                       wall-clock polls); 0: audio clock, tick k sees
                                       floor(1600 k / frame_size) * frame_size samples                  */
    int32_t reserved;
} ewk_stream_params;

typedef struct ewk_event {
    int32_t stream;
    int32_t kind;          /* 2: level-2 evaluation of a level-1 candidate (wakeword.py:1121); 1: timeout */
    int64_t tick;          /* tick index k (time = k * 0.1 s) at which the reference would have acted     */
    int64_t seg_start;     /* absolute sample index of the first sample of the extracted word_audio       */
    int32_t seg_len;       /* len(word_audio)                                       wakeword.py:1109-1111 */
    int32_t template_slot; /* best-scoring template slot                                                  */
    float score;           /* WordMatcher.calculate_similarity                      wakeword.py:591-625   */
    int32_t matched;       /* score >= similarity_threshold                         wakeword.py:638-639   */
} ewk_event;

typedef struct ewk_stream_status {
    int64_t written;            /* samples pushed                                                       */
    int64_t visible;            /* samples the last tick saw                                            */
    int64_t tick;               /* ticks processed                                                      */
    double silence_threshold;   /* SoundBuffer.silence_threshold                     wakeword.py:431,486 */
    double last_rms;            /* RMS of the last 0.1 s at the last tick            wakeword.py:495     */
    int32_t frame_size;         /* SoundBuffer.frame_size                                               */
    int32_t state;              /* 0 waiting, 1 in_silence, 2 in_sound, 3 after_sound                   */
    int32_t started;            /* buffer was full and _detect_word is running                          */
    int32_t is_silent;          /* SoundBuffer.is_silent() at the last tick          wakeword.py:488-496 */
    int32_t n_timeouts;
    int32_t n_events;
} ewk_stream_status;

typedef struct ewk_stream_result {  /* dense per-stream record, 8 bytes: what multi-GPU runs gather      */
    float score;                    /* latest level-2 score of the stream (NaN before the first)         */
    uint32_t flags;                 /* bit0 matched | bit1 silent | bits2-3 state | bit4 event in the last
                                       ewk_tick | bits 8..31 events so far                               */
} ewk_stream_result;

typedef struct ewk_vad_result {     /* template voice-activity analysis, 32 bytes                          */
    double duration_s;              /* max((last - first) * 160 / 16000, 0.2); 0 when voiced == 0         */
    float max_rms, threshold;       /* max frame RMS and max_rms * 0.1 (float32)                          */
    int32_t first_frame, last_frame;/* first / last frame with rms > threshold (-1 when none)             */
    int32_t n_frames;               /* 1 + len / 160                                                      */
    int32_t voiced;                 /* 0: no frame above the threshold (the reference returns None)       */
} ewk_vad_result;

/* defaults of the reference (wakeword.py:31-48, 408-409, 676) with audio-clock ticks */
int ewk_default_stream_params(ewk_stream_params* out);
/* stream == -1 applies to every stream. */
int ewk_set_stream_params(ewk_ctx* ctx, int stream, const ewk_stream_params* p);
/* SoundBuffer._add_sound_to_buffer (wakeword.py:454-465) for n_streams streams at once: n samples per
 * stream, stream s reads pcm + s*stride (in samples, format = the ring's).  `where` tells whether pcm is
 * a host pointer (pinned memory makes the copy asynchronous) or a device pointer.  Host PCM is copied on a
 * dedicated copy stream into one of two staging buffers and lands in the rings (K1) when something needs it:
 * the next push, a tick that reaches into it, or a read of stream state; a PINNED host buffer must therefore
 * stay unchanged until one of those calls (or ewk_synchronize) returns.  Pageable buffers are safe on return. */
int ewk_push(ewk_ctx* ctx, int stream0, int n_streams, const void* pcm, int64_t n, int64_t stride, int where);
/* The same for a G.711 feed: 8-bit mu-law (law = 0) / A-law (law = 1) codes — the telephony wire format, and WAV format
 * tags 7 / 6 that libsndfile decodes under librosa.load (wakeword.py:588).  The codes are expanded on the device to the
 * standard's 16-bit linear samples and pushed like PCM16, so a host feed crosses PCIe at one byte per sample.  The
 * context's rings must be EWK_PCM_I16.  Results equal those of pushing the expanded samples with ewk_push, bit for bit. */
int ewk_push_g711(ewk_ctx* ctx, int stream0, int n_streams, const uint8_t* codes, int64_t n, int64_t stride, int where, int law);
/* n_ticks polls of WakeWord._detect_word (wakeword.py:1064-1157) for every stream: adaptive threshold,
 * is_silent, timing state machine, segment cut, then the fused MFCC+match kernel on every candidate.
 * Asynchronous; results are read with ewk_poll / ewk_stream_results. */
int ewk_tick(ewk_ctx* ctx, int n_ticks);
/* Same, also recording per-tick traces [n_streams][n_ticks] (any pointer may be NULL; host memory). */
int ewk_tick_trace(ewk_ctx* ctx, int n_ticks, uint8_t* silent, uint8_t* state, double* thr, double* rms);
/* Drain the event queue into out[cap] (sorted by tick, then stream); returns the number of events
 * (>= 0) or a negative status.  *dropped (may be NULL) counts events lost to a full queue. */
int ewk_poll(ewk_ctx* ctx, ewk_event* out, int cap, int* dropped);
int ewk_stream_status_get(ewk_ctx* ctx, int stream, ewk_stream_status* out);
/* SoundBuffer.return_last_n_seconds (wakeword.py:498-513): the last n_samples visible samples. */
int ewk_read_last(ewk_ctx* ctx, int stream, int64_t n_samples, float* out);
/* word_audio of an event (wakeword.py:1105-1111) for the level-3 hand-off; EWK_ERR_STATE if overwritten. */
int ewk_read_segment(ewk_ctx* ctx, int stream, int64_t seg_start, int64_t seg_len, float* out);
/* Level-3 hand-off, batched: what WakeWord._transcribe_audio does to word_audio before the
 * speech-to-text call (wakeword.py:1020-1025): x - mean(x), / max|.| when > 0, * 1.5, clip to [-1, 1] — for
 * n_seg segments of the rings at once (typically the matched events of one ewk_poll).  Segment i is
 * streams[i], absolute samples [starts[i], starts[i] + lens[i]); its float32 result is written at
 * out + out_offsets[i].  `where` tells whether `out` is host or device memory.  EWK_ERR_STATE if a segment was
 * already overwritten in its ring. */
int ewk_prepare_segments(ewk_ctx* ctx, int n_seg, const int32_t* streams, const int64_t* starts, const int64_t* lens,
                         const int64_t* out_offsets, float* out, int64_t out_len, int where);
/* Overlap of level-2 matching with the next push (off by default).  When enabled, ewk_tick launches K3 on a second
 * stream right after the gate: the next ewk_push (K1, HBM-bound) then runs beside K3's tail instead of behind it.
 * K3 only reads ring samples that the push guard already protects and writes event / result records, and every other
 * stream-bank call (the next ewk_tick, ewk_poll, ewk_stream_results, ewk_read_*, ...) first makes the context's
 * stream wait for it, so results are unchanged.  The one visible difference: work that the CALLER enqueues on the
 * context's stream right after ewk_tick (e.g. an NCCL all-gather of ewk_results_device_ptr) must be preceded by
 * ewk_join, which enqueues that wait without blocking the host.  Rings shorter than 3 s + the ticks of a call keep the
 * sequential order. */
int ewk_set_overlap(ewk_ctx* ctx, int enable);
int ewk_join(ewk_ctx* ctx);
/* Template analysis, batched: WakeWord._analyze_reference_audio_duration
 * (wakeword.py:872-893) for n templates at once — librosa.feature.rms with 25 ms frames every 10 ms (zero-padded
 * centring), frames above 0.1 * max RMS, first-to-last span, floor 0.2 s.  Template i is the float32 samples
 * pcm[offsets[i] : offsets[i] + lens[i]] (16 kHz; `where` says host or device).  rms_out (nullable, host) receives
 * the frame RMS values of all templates back to back (n_frames each; rms_cap = its capacity in floats).  The
 * caller derives speech_duration_min/max from duration_s as the reference's tests pin it
 * (tests/test_wakeword_simulated.py:687-775: min = duration or 0.3, max = 2 * min or 2.0). */
int ewk_analyze_templates(ewk_ctx* ctx, const float* pcm, int where, const int64_t* offsets, const int64_t* lens,
                          int n, ewk_vad_result* out, float* rms_out, int64_t rms_cap);
/* Ingest at other sample rates: conversion to 16 kHz on the device, the step the reference
 * delegates to librosa.load(sr=16000) / librosa.resample (wakeword.py:588, 866-870; examples/tune_threshold.py:
 * 33-47; soxr HQ).  The filter meets soxr HQ's published specification (linear phase, pass-band to 0.913 of the
 * lower Nyquist, stop-band from 1.0, 125 dB) as one Kaiser-windowed-sinc polyphase stage; it is NOT soxr bit for
 * bit (oracle/resample_restated.py: parity unpinned).  Framing follows librosa.resample: output sample n sits at
 * input time n * sr_in / 16000, zeros outside the input, no gain rescale, one-shot length ceil(n_in*16000/sr_in).
 *
 * ewk_resample computes, for each of n_rows rows, absolute output samples [out_first, out_first + n_out) from
 * the absolute input samples [in_first, in_first + n_in) it is given (zeros elsewhere): one-shot use passes
 * in_first = out_first = 0; streaming use passes each chunk with half_width samples of history and emits only
 * outputs whose half_width samples of look-ahead are present, which reproduces the one-shot result exactly.
 * in: float32 or int16 (pcm_format), rows in_stride samples apart; out: float32, rows out_stride apart.
 * Supported rates: those with 16000 / gcd(sr_in, 16000) <= 4096 (every standard rate). */
int ewk_resample_info(int sr_in, int32_t* half_width, int32_t* up, int32_t* down);
int64_t ewk_resample_out_len(int64_t n_in, int sr_in);
int ewk_resample(ewk_ctx* ctx, const void* in, int pcm_format, int where_in, int n_rows, int64_t n_in, int64_t in_stride,
                 int sr_in, int64_t in_first, int64_t out_first, int64_t n_out, float* out, int64_t out_stride,
                 int where_out);
/* Dense per-hop scoring (usage shape of examples/tune_threshold.py:86-116 at hop
 * granularity): for every stream, every hop h in [hop0, hop0 + n_hops) (hop h <-> 160*h samples pushed)
 * and every template slot k in [template_first, template_first + template_count), the value
 * WordMatcher.calculate_similarity (wakeword.py:591-625) returns for the window
 *     x[160*(h - n_k) : 160*(h - n_k) + L_k],  n_k = ceil(L_k / 160)
 * i.e. the latest template-length window that starts on the hop grid and is complete at hop h
 * (window-local zero-pad centring and top_db floor).  out[(stream * n_hops + (h - hop0)) * template_count
 * + k], NaN where the window starts before the stream.  The audio must have been pushed and still be in
 * the ring (EWK_ERR_STATE otherwise).  Templates must have 640 <= L_k <= 48000 samples (the reference's own 3.0 s
 * segment cap, wakeword.py:1114-1118); template_count <= 8 per call.  Left-edge frames are shared by all templates, right-
 * edge frames by templates with the same padding to the hop grid; every call may continue where the previous one stopped.
 * `where` tells whether `out` is host or device memory. */
int ewk_dense_scores(ewk_ctx* ctx, int64_t hop0, int n_hops, int template_first, int template_count,
                     float* out, int where);
/* Copy the dense per-stream results [n_streams] to host memory. */
int ewk_stream_results(ewk_ctx* ctx, ewk_stream_result* out);
/* Device pointer of that array (ewk_stream_result[n_streams]); or make the kernels write into a
 * caller-owned device buffer instead (e.g. a registered NCCL send buffer). */
int ewk_results_device_ptr(ewk_ctx* ctx, void** out);
int ewk_set_results_buffer(ewk_ctx* ctx, void* device_ptr);
/* Peer publication — the multi-GPU "gather" without a collective (the reference's multiroom shape,
 * examples/multiroom_async.py:14-35, has no exchange at all: N objects report to one host thread).
 * bases[p], p < n_bases <= 16, are device pointers — local, or peer-mapped over NVLink (CUDA IPC / VMM /
 * torch symmetric memory) — to arrays of 2 * stride_records ewk_stream_result.  After this call every ewk_tick ends
 * with the last K3 CTA snapshotting the per-stream records, and a small sender kernel on a side stream of the context
 * (nothing of the next push / gate waits for it) storing the snapshot at
 *     bases[p][parity * stride_records + offset_records + stream]      for every p,
 * where the k-th ewk_tick call after this one (k = 1, 2, ...; ewk_publish_seq returns the latest k) uses
 * parity = (k - 1) & 1 (ewk_publish_parity), so a consumer reads a complete, stable copy of call k while call k + 1
 * is being produced.
 * signals (optional, may be NULL): signals[p] points to uint64_t[2][16] at destination p.  Behind its stores to
 * destination p the sender stores k at signals[p][parity][slot] with release semantics at system scope
 * (a put-with-signal): whoever reads k in slot r of its own copy holds every record rank r published up to call k.
 * bases[slot] / signals[slot] must be this context's own copy (ewk_wait_published / ewk_published_seq read it).
 * Without signals the records are globally visible once the sender of the call has completed on ewk_match_stream()
 * and a cross-GPU barrier enqueued there does the job.  n_bases = 0 switches publication off. */
int ewk_set_results_peers(ewk_ctx* ctx, void* const* bases, int n_bases, int64_t stride_records, int64_t offset_records,
                          void* const* signals, int slot);
int ewk_publish_parity(const ewk_ctx* ctx);
int64_t ewk_publish_seq(const ewk_ctx* ctx);
/* Enqueue, on the context's stream, a wait until slots [0, n_slots) of the local signal row of call `seq` hold a
 * value >= seq — everything enqueued afterwards sees all those ranks' records of that call.  The wait gives up after
 * timeout_ms (it never hangs the device); ewk_published_seq then shows which slot is behind. */
int ewk_wait_published(ewk_ctx* ctx, int n_slots, int64_t seq, int timeout_ms);
/* Synchronous read of the local signal row of `parity`: out[0 .. n_slots).  Returns 1 if an earlier
 * ewk_wait_published timed out since the last call of this function, else 0 (negative: error). */
int ewk_published_seq(ewk_ctx* ctx, int parity, uint64_t* out, int n_slots);
/* The cudaStream_t on which the latest ewk_tick's results become complete: the sender's stream while peer publication is
 * on, else the stream K3 was launched on (the match stream in overlap mode, else the context's stream). */
int ewk_match_stream(ewk_ctx* ctx, void** out);
/* Pinned host memory for asynchronous pushes. */
int ewk_host_alloc(void** out, int64_t bytes);
int ewk_host_free(void* p);
/* Kernel launches issued by this context so far (for bench accounting). */
int64_t ewk_launch_count(const ewk_ctx* ctx);
/* Per-kernel device timing with CUDA events on the context's stream.  ewk_profile(ctx, 1) starts
 * bracketing every kernel launch with an event pair; ewk_profile_read synchronises and returns, per
 * kernel class (0 ring_push, 1 tick_gate, 2 segment_queue, 3 segment_batch, 4 dense_score, 5 segment_prepare, 6 publish), the summed
 * milliseconds and the number of launches since profiling was enabled, then resets the sums. */
enum { EWK_PROF_CLASSES = 8 };
int ewk_profile(ewk_ctx* ctx, int enable);
int ewk_profile_read(ewk_ctx* ctx, double* ms, int64_t* launches);

#ifdef __cplusplus
}
#endif
#endif /* EWK_H */

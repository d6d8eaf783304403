// Context object behind the opaque ewk_ctx handle of include/ewk.h.
#pragma once
#include <cuda_runtime.h>

#include <map>
#include <string>
#include <vector>

#include "ewk_frame.cuh"
#include "ewk_segment.cuh"
#include "ewk_streams.cuh"
#include "ewk_dense.cuh"
#include "ewk_vad.cuh"
#include "ewk_resample.cuh"
#include "ewk_tables.hpp"

#define EWK_MAX_TEMPLATES 64

namespace ewk {

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 4 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void free() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

}  // namespace ewk

struct ewk_ctx {
    int device = 0;
    int sm_count = 0;
    bool use_lm = true;          // keep log-mel rows of K3 frames in a global workspace (EWK_SEG_LM=0 disables)
    bool k3_frames = false;      // EWK_K3=2: queue-form K3 with frames as the unit of work (default: one CTA per segment)
    ewk_config cfg{};
    std::string err;
    cudaStream_t own_stream = nullptr, stream = nullptr, copy_stream = nullptr;
    cudaEvent_t ev_ready[2] = {nullptr, nullptr}, ev_free[2] = {nullptr, nullptr};
    bool ev_free_valid[2] = {false, false};
    // optional overlap of K3 (level-2 matching) with the next push: K3 runs on match_stream after the gate (ev_gate) and the
    // context's stream joins it (ev_match) at the next call that needs level-2 results
    cudaStream_t match_stream = nullptr;
    cudaEvent_t ev_gate = nullptr, ev_match = nullptr;
    bool overlap = false, match_inflight = false;
    long long publish_seq = 0;               // ewk_tick calls since ewk_set_results_peers (peer publication); parity = (seq - 1) & 1
    int* d_wait_flag = nullptr;              // set by peer_wait_kernel when it gave up
    cudaStream_t pub_stream = nullptr;       // publish_records_kernel: the call's records -> every destination, then the signal
    cudaEvent_t ev_k3 = nullptr, ev_pub[2] = {nullptr, nullptr};
    bool ev_pub_valid[2] = {false, false};
    void* d_pub_snap = nullptr;              // StreamResult [2][n_streams]: K3's snapshots, one per parity
    cudaStream_t last_match_stream = nullptr;   // where the latest ewk_tick launched K3
    int join_match();
    int stage_idx = 0;
    struct Pending { bool valid = false; int slot = 0, stream0 = 0, n_streams = 0; long long n = 0; } pending;
    int land(int stream0, int n_streams, const void* d_src, long long d_stride, long long n, int stage_slot);
    int flush_pending();
    ewk::DevBuf b_stage2[2];
    ewk::DevBuf b_raw[2];                    // G.711 codes of a host push (ewk_push_g711), decoded into b_stage2
    ewk::DeviceTables* d_tables = nullptr;
    ewk::TemplateFeat* d_tmpl = nullptr;
    std::vector<ewk::TemplateFeat> h_tmpl;
    ewk::DevBuf b_pcm, b_desc, b_ws, b_lm, b_feat, b_scores, b_matched, b_frames, b_off;

    void fail(const char* fmt, ...);
    int init();
    void release();
    // stream bank
    ewk::BankView bank{};
    void* own_results = nullptr;
    ewk::DevBuf b_trace, b_read, b_dense, b_keep_rows, b_keep_end, b_g2;
    int dense_plan_info[4] = {0, 0, 0, 0};  // K4 geometry of the latest ewk_dense_scores: hops per sub-chunk, threads, CTAs per SM, smem bytes
    int chunk_cap = 0;
    bool all_presummed = false;            // K1's block sums cover every sample pushed since the last tick
    int pushes_since_tick = 0;
    long long launches = 0;
    std::vector<ewk::StreamParams> h_prm;
    double max_post = 0.4;                 // largest post_speech_silence of any stream (overlap-mode ring reserve)
    std::vector<long long> h_written;      // host mirror of StreamState.written
    std::vector<long long> h_visible_lb;   // lower bound of StreamState.visible (audio-clock overrun check)
    std::vector<long long> h_tick;
    std::vector<int> h_frame_size;         // latched frame size per stream (0: no push yet)
    int gate_chunks() const;               // largest R / frame_size in use
    // per-kernel event timing (ewk_profile)
    bool prof_on = false;
    struct ProfPair { cudaEvent_t a, b; int cls; };
    std::vector<ProfPair> prof_pairs;
    std::vector<cudaEvent_t> prof_free;
    double prof_ms[8] = {0};
    long long prof_n[8] = {0};
    cudaEvent_t prof_begin(int cls, cudaStream_t on = nullptr);
    void prof_end(cudaEvent_t a, int cls, cudaStream_t on = nullptr);
    int prof_collect();
    int init_streams();
    void release_streams();
    int launch_segments(const ewk::SegDesc* d_segs, int n_seg, int max_frames, long long spill_frames,
                        long long lm_frames, int n_tmpl, int tmpl_first, float threshold, float* d_feat, float* d_frames,
                        float* d_scores, unsigned char* d_matched);
    int queue_grid() const { return sm_count > 0 ? 2 * sm_count : 1; }     // persistent K3 CTAs (2 per SM)
    static constexpr size_t LM_WS_MAX_BYTES = (size_t)4 << 30;
    // K7 filter tables, one per input rate seen
    struct ResampleTable { ewk::ResampleDesign d; float* H = nullptr; };
    std::map<int, ResampleTable> rs_tables;
    ewk::DevBuf b_rs_in, b_rs_out;
    int resample_table(int sr_in, const ResampleTable** out);
};

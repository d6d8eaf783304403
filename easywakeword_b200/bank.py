"""WakeWordBank — the batched ("multiroom") form of the reference's detector.

The reference runs N rooms as N independent WakeWord objects and N threads
(/root/reference/examples/multiroom_async.py:14-35).  Here N streams share one device context:
rings in HBM, and per 100 ms tick of audio one K2 launch (adaptive threshold, is_silent, timing state
machine for every stream) plus one persistent K3 launch (fused MFCC + template match on every candidate
segment).  Host code only pushes PCM, ticks, and polls events; level 3 (Whisper) is handed
`read_segment(event)` audio on the host, exactly where the reference calls `_transcribe_audio`
(wakeword.py:1126-1128).
"""
from __future__ import annotations

from typing import Callable, Iterable, Optional, Sequence

import numpy as np

from . import _lib
from .wakeword import WakeWord, load_wav_16k

TICK_SAMPLES = 1600
EV_TIMEOUT, EV_SCORED = _lib.EV_TIMEOUT, _lib.EV_SCORED


class WakeWordBank:
    """N independent 16 kHz streams, one GPU.

    templates: list of float32 arrays or WAV paths (slot i = templates[i]); every stream is scored
    against all of them (best score wins) unless set_stream(..., template_first=, template_count=).
    Timing defaults are the reference's (wakeword.py:38-41); speech_duration_min/max default to the
    auto-calculated values of the first template (tests/test_wakeword_simulated.py:687-775)."""

    def __init__(self, n_streams: int, templates: Sequence, *, device: int = 0, buffer_seconds: int = 10,
                 pcm_dtype=np.int16, frame_size: int = TICK_SAMPLES, similarity_threshold: float = 75.0,
                 pre_speech_silence: float = 0.8, speech_duration_min: Optional[float] = None,
                 speech_duration_max: Optional[float] = None, post_speech_silence: float = 0.4,
                 timeout: float = 30.0, max_push_seconds: float = 1.0, max_events: int = 0,
                 cuda_stream: Optional[int] = None, overlap: bool = False, preemphasis: float = 0.0, n_mfcc: int = 20):
        if n_streams < 1:
            raise ValueError("n_streams must be at least 1")
        if buffer_seconds <= 0:
            raise ValueError("buffer_seconds must be positive")
        if not templates:
            raise ValueError("at least one template is required")
        self.n_streams = n_streams
        self.pcm_dtype = np.dtype(pcm_dtype)
        if self.pcm_dtype not in (np.dtype(np.int16), np.dtype(np.float32)):
            raise ValueError("pcm_dtype must be int16 or float32")
        fmt = _lib.PCM_I16 if self.pcm_dtype == np.int16 else _lib.PCM_F32
        slack = int(max_push_seconds * 16000) + 2 * max(frame_size, TICK_SAMPLES)
        self.ctx = _lib.Context(device=device, n_streams=n_streams, ring_samples=buffer_seconds * 16000,
                                slack_samples=slack, pcm_format=fmt, max_templates=max(1, len(templates)),
                                max_events=max_events or max(4096, 2 * n_streams), preemphasis=preemphasis,
                                n_mfcc=n_mfcc)
        if cuda_stream is not None:
            self.ctx.set_cuda_stream(cuda_stream)
        if overlap:
            # level-2 matching on a second stream beside the next push (include/ewk.h: ewk_set_overlap); work that the
            # caller enqueues on `cuda_stream` after tick() and that reads the results must follow self.ctx.join()
            self.ctx.set_overlap(True)
        self.templates = []
        for slot, t in enumerate(templates):
            audio = load_wav_16k(t) if isinstance(t, (str, bytes)) or hasattr(t, "__fspath__") else \
                np.ascontiguousarray(t, dtype=np.float32)
            self.ctx.set_template(slot, audio)
            self.templates.append(audio)
        # voice-activity duration of every template on the device (K6); the bank's default timing follows the first
        # template, as the reference's single-template WakeWord does (tests/test_wakeword_simulated.py:687-775)
        self.template_vad = self.ctx.analyze_templates(self.templates)
        if speech_duration_min is None:
            d = float(self.template_vad["duration_s"][0]) if self.template_vad["voiced"][0] else None
            speech_duration_min = d if d is not None else 0.3
            if speech_duration_max is None:
                speech_duration_max = 2.0 * speech_duration_min if d is not None else 2.0
        elif speech_duration_max is None:
            speech_duration_max = 2.0 * speech_duration_min
        self.params = dict(frame_size=frame_size, similarity_threshold=similarity_threshold,
                           pre_speech_silence=pre_speech_silence, speech_duration_min=speech_duration_min,
                           speech_duration_max=speech_duration_max, post_speech_silence=post_speech_silence,
                           timeout=float(timeout), template_first=0, template_count=len(templates))
        self.ctx.set_stream_params(-1, **self.params)
        self.ticks = 0
        self.samples_pushed = 0

    # ---- configuration
    def set_stream(self, stream: int, **overrides):
        """Per-stream parameters (any field of ewk_stream_params)."""
        p = dict(self.params)
        p.update(overrides)
        self.ctx.set_stream_params(stream, **p)

    # ---- data plane
    def push(self, pcm, where=_lib.HOST):
        """pcm[n_streams, n] in the bank's dtype: n new samples for every stream (K1 / direct H2D)."""
        self.ctx.push(pcm, stream0=0, where=where)
        self.samples_pushed += pcm[2] if isinstance(pcm, tuple) else pcm.shape[-1]

    def push_g711(self, codes, law: str = "ulaw", where=_lib.HOST):
        """codes[n_streams, n] uint8: a G.711 mu-law / A-law feed, expanded to PCM16 on the device (int16 banks)."""
        self.ctx.push_g711(codes, stream0=0, where=where, law=law)
        self.samples_pushed += codes[2] if isinstance(codes, tuple) else codes.shape[-1]

    def tick(self, n_ticks: int = 1, trace: bool = False):
        """n_ticks x 100 ms of the reference's poll loop for every stream (K2 + K3)."""
        self.ticks += n_ticks
        return self.ctx.tick(n_ticks, trace=trace)

    def step(self, pcm, where=_lib.HOST):
        """push + as many ticks as the pushed audio covers."""
        n = pcm[2] if isinstance(pcm, tuple) else pcm.shape[-1]
        self.push(pcm, where)
        due = self.samples_pushed // TICK_SAMPLES - self.ticks
        if due > 0:
            self.tick(due)

    def poll(self) -> np.ndarray:
        """Structured array of events since the last poll, sorted by (tick, stream):
        kind 2 = level-2 evaluation (score, matched), kind 1 = timeout."""
        return self.ctx.poll()

    def results(self) -> np.ndarray:
        return self.ctx.results()

    def dense_scores(self, hop0: int, n_hops: int, template_first: int = 0, template_count: Optional[int] = None):
        """Per-hop scores [n_streams, n_hops, T] for hops [hop0, hop0 + n_hops) (hop h = 160*h samples
        pushed): calculate_similarity of the latest template-length window complete at each hop (K4)."""
        T = len(self.templates) - template_first if template_count is None else template_count
        return self.ctx.dense_scores(hop0, n_hops, template_first, T)

    def read_segment(self, event) -> np.ndarray:
        """word_audio of a level-2 event as float32 (what level 3 transcribes)."""
        return self.ctx.read_segment(int(event["stream"]), int(event["seg_start"]), int(event["seg_len"]))

    def dense_sweep(self, blocks: Iterable[np.ndarray], template_first: int = 0, template_count: Optional[int] = None):
        """Offline sweep (BASELINE config 5): push each [n_streams, n] block (n a multiple of 160) and yield
        (first_hop, scores[n_streams, n // 160, T]) for the hops it completes — every template scored at every hop."""
        self.ctx.set_stream_params(-1, **dict(self.params, live=1))     # no ticks here: pushes run free of the gate's guard
        for blk in blocks:
            n = blk.shape[-1]
            if n % 160:
                raise ValueError("dense_sweep blocks must hold a multiple of 160 samples")
            hop0 = self.samples_pushed // 160 + 1
            self.push(blk)
            yield hop0, self.dense_scores(hop0, n // 160, template_first, template_count)

    def prepare_for_transcription(self, events) -> list:
        """Batched level-3 pre-processing on the device (wakeword.py:1020-1025) of the word_audio of `events`
        (any structured rows with stream / seg_start / seg_len): list of float32 arrays ready for the STT backend."""
        ev = [e for e in events]
        return self.ctx.prepare_segments([int(e["stream"]) for e in ev], [int(e["seg_start"]) for e in ev],
                                         [int(e["seg_len"]) for e in ev])

    def status(self, stream: int):
        return self.ctx.status(stream)

    # ---- convenience: the reference's callback surface over the whole bank
    def run(self, blocks: Iterable[np.ndarray], on_match: Optional[Callable] = None, transcriber=None,
            textword: str = "", numberofwords: int = 1):
        """Feed an iterable of [n_streams, n] blocks; for every level-2 match call
        on_match(stream, tick, score, text) where text is the level-3 confirmation (None without a
        transcriber).  Returns the list of all events."""
        log = []
        for blk in blocks:
            self.step(blk)
            ev = self.poll()
            log.extend(e.copy() for e in ev)
            hits = [e for e in ev if e["kind"] == EV_SCORED and e["matched"]]
            if not hits or on_match is None:
                continue
            audio = self.prepare_for_transcription(hits) if transcriber is not None else [None] * len(hits)
            for e, a in zip(hits, audio):
                text = transcriber.transcribe(a) if transcriber is not None else None
                on_match(int(e["stream"]), int(e["tick"]), float(e["score"]), text)
        return log

    def close(self):
        self.ctx.close()

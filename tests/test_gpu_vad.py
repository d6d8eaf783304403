"""GPU parity: K6 template_vad (ewk_analyze_templates) vs the durations the reference's own
WakeWord._analyze_reference_audio_duration returned (tests/golden/vad.npz) and vs the oracle on random templates.
Durations and voiced spans are integers times 0.01 s: they must be identical; frame RMS within 1e-6 relative."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

RMS_RTOL = 1e-6


@pytest.fixture(scope="module")
def ctx():
    from easywakeword_b200 import _lib
    c = _lib.Context(device=0, n_streams=0, max_templates=1)
    yield c
    c.close()


def test_durations_match_reference_goldens(ctx, golden_vad):
    g = golden_vad
    names = [str(n) for n in g["names"]]
    audios = [g[f"pcm_{n}"].astype(np.float32) / np.float32(32768.0) for n in names]
    res, rms = ctx.analyze_templates(audios, want_rms=True)
    for i, n in enumerate(names):
        ref = float(g[f"duration_{n}"])
        assert res["n_frames"][i] == 1 + len(audios[i]) // 160 == len(rms[i]), n
        if np.isnan(ref):
            assert not res["voiced"][i] and res["first_frame"][i] == -1, n
        else:
            assert res["voiced"][i] and res["duration_s"][i] == ref, (n, res["duration_s"][i], ref)
        gr = g[f"rms_{n}"]
        assert np.abs(rms[i] - gr).max() <= RMS_RTOL * max(float(gr.max()), 1e-30), n
        assert abs(res["max_rms"][i] - gr.max()) <= RMS_RTOL * max(float(gr.max()), 1e-30), n


def test_random_templates_match_oracle(ctx):
    from oracle import ewk_oracle as O
    rng = np.random.default_rng(2024)
    audios = []
    for k in range(64):
        n = int(rng.integers(1, 40000))
        x = rng.standard_normal(n).astype(np.float32) * np.float32(0.003)
        a, b = sorted(rng.integers(0, n, 2))
        x[a:b] += rng.standard_normal(b - a).astype(np.float32) * np.float32(rng.uniform(0.02, 0.5))
        if k % 9 == 0:
            x[:] = 0
        audios.append(x)
    audios.append(np.zeros(0, np.float32))                    # empty template: one all-zero frame, not voiced
    res = ctx.analyze_templates(audios)
    for i, x in enumerate(audios):
        d = O.analyze_reference_audio_duration(x) if len(x) else None
        if d is None:
            assert not res["voiced"][i], i
        else:
            assert res["voiced"][i] and res["duration_s"][i] == d, (i, res["duration_s"][i], d)


def test_bank_default_durations_come_from_device_vad(word):
    """WakeWordBank without explicit durations: min = VAD duration of the first template, max = 2 * min
    (the _auto_calculate_speech_durations contract, reference tests/test_wakeword_simulated.py:687-775)."""
    from easywakeword_b200.bank import WakeWordBank
    from oracle import ewk_oracle as O
    bank = WakeWordBank(2, [word, word[:8000]], device=0)
    try:
        smin, smax = O.auto_speech_durations(word)
        assert bank.params["speech_duration_min"] == smin and bank.params["speech_duration_max"] == smax
        assert bank.template_vad["duration_s"][1] == O.analyze_reference_audio_duration(word[:8000])
    finally:
        bank.close()

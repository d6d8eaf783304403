"""CPU: the C-ABI library loads, exports every symbol include/ewk.h declares, its host-side tables
equal the oracle's, and it fails loudly without a GPU (no CPU fallback)."""
import os
import re

import numpy as np
import pytest

from conftest import REPO


@pytest.fixture(scope="module")
def lib():
    from easywakeword_b200 import _lib, build
    build.build()
    return _lib.load()


def test_header_symbols_exported(lib):
    hdr = open(os.path.join(REPO, "include", "ewk.h")).read()
    names = sorted(set(re.findall(r"\b(ewk_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    # and every symbol the binding declares is in the header
    assert set(lib._protos) <= set(names), set(lib._protos) - set(names)
    assert lib.ewk_abi_version() == 2


def test_struct_layouts_match_header():
    import ctypes as C
    from easywakeword_b200 import _lib
    assert C.sizeof(_lib.Config) == 40      # ABI 2: + preemphasis, n_mfcc, 2 reserved words
    assert C.sizeof(_lib.StreamParams) == 72
    assert C.sizeof(_lib.Event) == 40 and _lib.EVENT_DTYPE.itemsize == 40
    assert C.sizeof(_lib.StreamStatus) == 64
    assert _lib.RESULT_DTYPE.itemsize == 8


def test_host_tables_equal_oracle(lib):
    import scipy.fft
    from easywakeword_b200 import _lib
    from oracle import librosa_restated as L
    assert np.array_equal(_lib.host_table(0), L.hann_window().astype(np.float32))
    assert np.array_equal(_lib.host_table(1).reshape(128, 257), L.mel_filterbank())
    D = scipy.fft.dct(np.eye(128), axis=0, type=2, norm="ortho")[:20]
    assert np.abs(_lib.host_table(2).reshape(20, 128) - D).max() < 1e-7


def test_default_stream_params_are_the_references(lib):
    from easywakeword_b200 import _lib
    p = _lib.default_stream_params()
    assert (p.similarity_threshold, p.pre_speech_silence, p.speech_duration_min, p.speech_duration_max,
            p.post_speech_silence, p.min_threshold) == (75.0, 0.8, 0.3, 2.0, 0.4, 0.005)


def test_no_gpu_means_loud_failure(lib):
    import torch
    from easywakeword_b200 import _lib
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.EwkError, match="no CPU fallback"):
        _lib.Context()


@pytest.mark.gpu
def test_plain_c_client_on_gpu(lib, tmp_path):
    """The GPU leg of test_plain_c_client (that one runs in the CPU suite, where the client must be refused)."""
    test_plain_c_client(lib, tmp_path)


def test_plain_c_client(lib, tmp_path):
    """include/ewk.h is a C header and libewk.so a C library: a C program builds against them with gcc alone and, on this
    box, either scores the reference's self-similarity known answer (GPU) or is refused loudly (no GPU)."""
    import shutil
    import subprocess
    import torch
    from easywakeword_b200 import _lib
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    exe = str(tmp_path / "c_abi_client")
    libdir = os.path.dirname(_lib.LIB_PATH)
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(REPO, "include"),
                        os.path.join(REPO, "tests", "c_abi_client.c"), "-o", exe, "-L", libdir, "-lewk", "-lm",
                        f"-Wl,-rpath,{libdir}"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    if torch.cuda.is_available():
        assert r.returncode == 0, r.stdout + r.stderr
        assert "self-similarity 100.0000  matched 1" in r.stdout and "No reference word set" in r.stdout
    else:
        assert r.returncode == 3, r.stdout + r.stderr
        assert "no CPU fallback" in r.stdout


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under easywakeword_b200/ may reference it."""
    pkg = os.path.join(REPO, "easywakeword_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h")):
                src = open(os.path.join(root, f)).read()
                assert not re.search(r"^\s*(from|import)\s+\.*oracle\b", src, re.M), f
                assert not re.search(r"import_module\(.*oracle|__import__\(.*oracle", src), f
                if f.endswith(".py"):
                    assert "librosa_restated" not in src and "ewk_oracle" not in src, f

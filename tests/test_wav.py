"""WAV I/O mirror of the reference's wav_header.h (canonical 44-byte header)."""
import numpy as np
import pytest

from digital_signal_processsing_b200 import wav


def test_header_is_44_packed_bytes():
    assert wav.HEADER_BYTES == 44
    h = wav.make_header(1000, 2, np.int16)
    raw = h.pack()
    assert len(raw) == 44 and raw[:4] == b"RIFF" and raw[8:16] == b"WAVEfmt " and raw[36:40] == b"data"
    assert wav.WAVHeader.unpack(raw) == h
    assert h.total_samples == 1000 and h.blockAlign == 4 and h.byteRate == 44100 * 4


@pytest.mark.parametrize("dtype", [np.int16, np.float32])
def test_roundtrip(tmp_path, dtype):
    rng = np.random.default_rng(0)
    x = (rng.standard_normal(2 * 777) * 1000).astype(dtype)
    p = str(tmp_path / "a.wav")
    h = wav.make_header(x.size, 2, dtype)
    wav.write_samples(p, h, x)
    h2, y = wav.extract_samples(p)
    assert h2 == h and y.dtype == dtype and np.array_equal(x, y)
    assert h2.numChannels == 2


def test_scipy_int16_files_parse(tmp_path):
    """run_benchmarks.py of the reference writes its inputs with scipy (int16 PCM: canonical header)."""
    sciwav = pytest.importorskip("scipy.io.wavfile")
    x = np.arange(-500, 500, dtype=np.int16).reshape(-1, 2)
    p = str(tmp_path / "s.wav")
    sciwav.write(p, 44100, x)
    h, y = wav.extract_samples(p)
    assert h.numChannels == 2 and h.bitsPerSample == 16 and np.array_equal(y, x.reshape(-1))


def test_rejects_unsupported(tmp_path):
    h = wav.make_header(10, 1, np.int16)
    h.bitsPerSample = 24
    p = tmp_path / "b.wav"
    p.write_bytes(h.pack() + b"\0" * 30)
    with pytest.raises(ValueError):
        wav.extract_samples(str(p))
    with pytest.raises(ValueError):
        wav.make_header(2**31, 1, np.float32)   # dataBytes is a uint32

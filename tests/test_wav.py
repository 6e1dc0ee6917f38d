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
    big = wav.make_header(2**31, 1, np.float32)   # 8 GiB of samples: the 32-bit fields saturate, RF64 carries the sizes
    assert big.dataBytes == 0xFFFFFFFF and big.sizeOfFile == 0xFFFFFFFF


def test_non_canonical_layouts(tmp_path):
    """scipy writes float32 with an 18-byte fmt chunk and a `fact` chunk (58-byte header); other tools put a LIST
    chunk in front of the data or use WAVE_FORMAT_EXTENSIBLE.  The reference's 44-byte struct mis-parses all of
    them; the new reader walks the chunks and hands back a canonical header."""
    sciwav = pytest.importorskip("scipy.io.wavfile")
    x = (np.arange(600, dtype=np.float32) / 600.0).reshape(-1, 2)
    p = str(tmp_path / "f32.wav")
    sciwav.write(p, 48000, x)
    h, y = wav.extract_samples(p)
    assert (h.audioFormat, h.numChannels, h.bitsPerSample, h.sampleRate) == (3, 2, 32, 48000)
    assert h.fmtSize == 16 and h.dataBytes == x.size * 4 and np.array_equal(y, x.reshape(-1))
    # hand-made: 16-byte fmt, a LIST chunk of odd size (padded), then data
    import struct
    s = np.arange(-50, 50, dtype=np.int16)
    fmt = struct.pack("<HHIIHH", 1, 1, 8000, 16000, 2, 16)
    lst = b"INFOabcde"                                  # 9 bytes -> one pad byte
    body = b"WAVE" + b"fmt " + struct.pack("<I", 16) + fmt + b"LIST" + struct.pack("<I", len(lst)) + lst + b"\0" \
        + b"data" + struct.pack("<I", s.nbytes) + s.tobytes()
    q = tmp_path / "list.wav"
    q.write_bytes(b"RIFF" + struct.pack("<I", len(body)) + body)
    h, y = wav.extract_samples(str(q))
    assert h.numChannels == 1 and h.bitsPerSample == 16 and np.array_equal(y, s)
    # WAVE_FORMAT_EXTENSIBLE float32
    ext = struct.pack("<HHIIHH", 0xFFFE, 2, 44100, 44100 * 8, 8, 32) + struct.pack("<HHI", 22, 32, 3) \
        + struct.pack("<H", 3) + bytes(14)
    z = np.linspace(-1, 1, 64, dtype=np.float32)
    body = b"WAVE" + b"fmt " + struct.pack("<I", len(ext)) + ext + b"data" + struct.pack("<I", z.nbytes) + z.tobytes()
    r = tmp_path / "ext.wav"
    r.write_bytes(b"RIFF" + struct.pack("<I", len(body)) + body)
    h, y = wav.extract_samples(str(r))
    assert h.audioFormat == 3 and h.numChannels == 2 and np.array_equal(y, z)


def test_rf64_roundtrip_and_foreign_layout(tmp_path):
    """RF64 (payloads of 4 GiB and more): forced on a small payload so that the CPU suite can check the layout."""
    import struct
    x = np.arange(-400, 400, dtype=np.int16)
    h = wav.make_header(x.size, 2, np.int16)
    p = tmp_path / "r.wav"
    wav.write_samples(str(p), h, x, force_rf64=True)
    raw = p.read_bytes()
    assert raw[:4] == b"RF64" and raw[4:8] == b"\xff" * 4 and raw[12:16] == b"ds64"
    riff_size, data_size, frames = struct.unpack("<QQQ", raw[20:44])
    assert riff_size == len(raw) - 8 and data_size == x.nbytes and frames == x.size // 2
    h2, y = wav.extract_samples(str(p))
    assert np.array_equal(x, y) and (h2.numChannels, h2.bitsPerSample, h2.dataBytes) == (2, 16, x.nbytes)
    assert h2.riff == b"RIFF" and h2.pack() == h.pack()       # callers only ever see the canonical header
    # ds64 with a table and a JUNK chunk in between, float32, sizes only in ds64
    z = np.linspace(-1, 1, 50, dtype=np.float32)
    body = b"WAVE" + b"ds64" + struct.pack("<IQQQI", 28 + 12, 0, z.nbytes, z.size, 1) + bytes(12) \
        + b"JUNK" + struct.pack("<I", 6) + bytes(6) \
        + b"fmt " + struct.pack("<IHHIIHH", 16, 3, 1, 8000, 32000, 4, 32) + b"data" + struct.pack("<I", 0xFFFFFFFF) + z.tobytes()
    q = tmp_path / "f.wav"
    q.write_bytes(b"RF64" + struct.pack("<I", 0xFFFFFFFF) + body)
    h3, y3 = wav.extract_samples(str(q))
    assert h3.audioFormat == 3 and np.array_equal(y3, z)


def test_cpp_reader_matches_python_reader(tmp_path):
    """host/mavg_wav.h (the reader of the drop-in binaries) parses the same files the same way -- checked on the
    CPU with a tiny harness that needs neither libmavg nor a GPU."""
    import os
    import subprocess
    sciwav = pytest.importorskip("scipy.io.wavfile")
    root = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
    src = tmp_path / "wavdump.cpp"
    src.write_text('''
#include "mavg_wav.h"
int main(int argc, char** argv) {
    WAVHeader h{}; std::vector<unsigned char> b; std::string why;
    if (argc > 2) {   // writer check: canonical or RF64 int16 ramp
        std::vector<int16_t> v(600); for (int i = 0; i < 600; ++i) v[i] = (int16_t)(i - 300);
        return mavg_wav::write_file(argv[1], mavg_wav::make_header<int16_t>(v.size(), 2), v.data(), v.size(), argv[2][0] == 'r') ? 0 : 3;
    }
    if (!mavg_wav::read_file(argv[1], h, b, &why)) { printf("ERR %s\\n", why.c_str()); return 2; }
    unsigned long long sum = 0; for (unsigned char c : b) sum = sum * 131 + c;
    printf("%u %u %u %u %zu %llu\\n", h.audioFormat, h.numChannels, h.bitsPerSample, h.sampleRate, b.size(), sum);
    return 0;
}''')
    exe = tmp_path / "wavdump"
    subprocess.run(["g++", "-O1", "-std=c++17", "-I", os.path.join(root, "host"), str(src), "-o", str(exe)], check=True)

    def digest(raw: bytes) -> int:
        s = 0
        for c in raw:
            s = (s * 131 + c) % (1 << 64)
        return s

    files = []
    a = tmp_path / "canon.wav"
    xi = np.arange(-300, 300, dtype=np.int16)
    wav.write_samples(str(a), wav.make_header(xi.size, 2, np.int16), xi)
    files.append(a)
    b = tmp_path / "scipy_f32.wav"
    sciwav.write(str(b), 22050, (np.arange(500, dtype=np.float32) / 7).reshape(-1, 1))
    files.append(b)
    c = tmp_path / "rf64.wav"
    wav.write_samples(str(c), wav.make_header(xi.size, 2, np.int16), xi, force_rf64=True)
    files.append(c)
    for f in files:
        h, y = wav.extract_samples(str(f))
        out = subprocess.run([str(exe), str(f)], capture_output=True, text=True, check=True).stdout.split()
        assert [int(v) for v in out[:5]] == [h.audioFormat, h.numChannels, h.bitsPerSample, h.sampleRate, y.nbytes]
        assert int(out[5]) == digest(y.tobytes())
    for mode in ("c", "r"):                                   # files written by the C++ writer, read by the Python reader
        w = tmp_path / f"cpp_{mode}.wav"
        subprocess.run([str(exe), str(w), mode], check=True)
        assert w.read_bytes()[:4] == (b"RF64" if mode == "r" else b"RIFF")
        h, y = wav.extract_samples(str(w))
        assert h.numChannels == 2 and np.array_equal(y, xi)
    bad = tmp_path / "bad.wav"
    bad.write_bytes(b"RIFFxxxxWAVEjunk" + bytes(64))
    r = subprocess.run([str(exe), str(bad)], capture_output=True, text=True)
    assert r.returncode == 2 and "ERR" in r.stdout

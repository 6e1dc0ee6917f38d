"""Canonical 44-byte-header WAV I/O, the input/output format of the `averager` binaries.

Mirrors the reference's wav_header.h: the packed `WAVHeader` (:8-24), `extractSamples`
(:26-48) and `writeSamples` (:50-59, header verbatim followed by the raw samples), with
two changes the new library needs: float32 samples (audioFormat 3, 32 bit) are accepted
next to int16 (the reference rejects them, :34), and samples are read in one bulk read
instead of one `read()` per sample (:41-45).  The C++ twin used by the drop-in binaries is
host/wav_header.h.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass
from typing import Tuple

import numpy as np

_FMT = "<4sI4s4sIHHIIHH4sI"
HEADER_BYTES = struct.calcsize(_FMT)
assert HEADER_BYTES == 44

PCM, IEEE_FLOAT = 1, 3


@dataclass
class WAVHeader:
    riff: bytes = b"RIFF"
    sizeOfFile: int = 36
    wave: bytes = b"WAVE"
    fmt: bytes = b"fmt "
    fmtSize: int = 16
    audioFormat: int = PCM
    numChannels: int = 1
    sampleRate: int = 44100
    byteRate: int = 88200
    blockAlign: int = 2
    bitsPerSample: int = 16
    data: bytes = b"data"
    dataBytes: int = 0

    def pack(self) -> bytes:
        return struct.pack(_FMT, self.riff, self.sizeOfFile, self.wave, self.fmt, self.fmtSize, self.audioFormat,
                           self.numChannels, self.sampleRate, self.byteRate, self.blockAlign, self.bitsPerSample,
                           self.data, self.dataBytes)

    @classmethod
    def unpack(cls, raw: bytes) -> "WAVHeader":
        return cls(*struct.unpack(_FMT, raw))

    @property
    def total_samples(self) -> int:
        return self.dataBytes // (self.bitsPerSample // 8)


def make_header(num_samples: int, channels: int, dtype, sample_rate: int = 44100) -> WAVHeader:
    dtype = np.dtype(dtype)
    if dtype == np.int16:
        fmt, bits = PCM, 16
    elif dtype == np.float32:
        fmt, bits = IEEE_FLOAT, 32
    else:
        raise TypeError("WAV samples must be int16 or float32")
    data_bytes = num_samples * (bits // 8)
    if data_bytes >= 2**32:
        raise ValueError("canonical WAV dataBytes is 32 bit: at most 2^32-1 bytes of samples")
    return WAVHeader(sizeOfFile=36 + data_bytes, audioFormat=fmt, numChannels=channels, sampleRate=sample_rate,
                     byteRate=sample_rate * channels * bits // 8, blockAlign=channels * bits // 8,
                     bitsPerSample=bits, dataBytes=data_bytes)


def _walk_chunks(f) -> WAVHeader:
    """General RIFF walk for files that are not in the canonical 44-byte form (18/40-byte fmt chunks,
    WAVE_FORMAT_EXTENSIBLE, `fact`/`LIST` chunks before the data -- what scipy writes for float32).  Leaves the
    file positioned at the payload and returns a synthesised canonical header."""
    f.seek(12)
    fmt = None
    while True:
        head = f.read(8)
        if len(head) < 8:
            raise ValueError("no fmt/data chunk found")
        cid, size = head[:4], struct.unpack("<I", head[4:])[0]
        if cid == b"fmt ":
            body = f.read(size + (size & 1))
            if size < 16:
                raise ValueError("truncated fmt chunk")
            tag, ch, rate, brate, align, bits = struct.unpack("<HHIIHH", body[:16])
            if tag == 0xFFFE and size >= 26:
                tag = struct.unpack("<H", body[24:26])[0]     # sub-format GUID starts with the format tag
            fmt = (tag, ch, rate, brate, align, bits)
        elif cid == b"data":
            if fmt is None:
                raise ValueError("data chunk before fmt chunk")
            tag, ch, rate, brate, align, bits = fmt
            return WAVHeader(sizeOfFile=36 + size, audioFormat=tag, numChannels=ch, sampleRate=rate, byteRate=brate,
                             blockAlign=align, bitsPerSample=bits, dataBytes=size)
        else:
            f.seek(size + (size & 1), 1)


def extract_samples(path: str) -> Tuple[WAVHeader, np.ndarray]:
    """(header, interleaved samples).  Raises on unreadable/unsupported files.  Canonical files are read as they
    are; other RIFF/WAVE layouts are walked chunk by chunk and come back with a canonical header."""
    with open(path, "rb") as f:
        raw = f.read(HEADER_BYTES)
        if len(raw) < HEADER_BYTES:
            raise ValueError("file shorter than a canonical WAV header")
        h = WAVHeader.unpack(raw)
        if h.riff != b"RIFF" or h.wave != b"WAVE":
            raise ValueError("not a RIFF/WAVE file")
        if not (h.fmt == b"fmt " and h.fmtSize == 16 and h.data == b"data"):
            h = _walk_chunks(f)
        if h.bitsPerSample == 16:
            dtype = np.int16
        elif h.bitsPerSample == 32 and h.audioFormat == IEEE_FLOAT:
            dtype = np.float32
        else:
            raise ValueError(f"unsupported bits per sample: {h.bitsPerSample}")
        samples = np.fromfile(f, dtype=dtype, count=h.total_samples)
    return h, samples


def write_samples(path: str, header: WAVHeader, samples: np.ndarray) -> None:
    """Input header verbatim + raw samples, as the reference's writeSamples does."""
    with open(path, "wb") as f:
        f.write(header.pack())
        np.ascontiguousarray(samples).tofile(f)

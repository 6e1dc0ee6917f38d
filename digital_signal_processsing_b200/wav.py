"""Canonical 44-byte-header WAV I/O, the input/output format of the `averager` binaries.

Mirrors the reference's wav_header.h: the packed `WAVHeader` (:8-24), `extractSamples`
(:26-48) and `writeSamples` (:50-59, header verbatim followed by the raw samples), with
the changes the new library needs: float32 samples (audioFormat 3, 32 bit) are accepted
next to int16 (the reference rejects them, :34), samples are read in one bulk read
instead of one `read()` per sample (:41-45), and payloads of 4 GiB and more are read and
written as RF64 (EBU Tech 3306: 32-bit sizes 0xFFFFFFFF, 64-bit sizes in a `ds64` chunk;
the reference's uint32 counts stop at 2^32 bytes).  The C++ twin used by the drop-in
binaries is host/mavg_wav.h.
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
_U32 = 0xFFFFFFFF


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
    dataBytes: int = 0        # 0xFFFFFFFF when the payload needs RF64; the sample array carries the true count

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
    return WAVHeader(sizeOfFile=min(36 + data_bytes, _U32), audioFormat=fmt, numChannels=channels,
                     sampleRate=sample_rate, byteRate=sample_rate * channels * bits // 8,
                     blockAlign=channels * bits // 8, bitsPerSample=bits, dataBytes=min(data_bytes, _U32))


def _walk_chunks(f, rf64: bool = False) -> Tuple[WAVHeader, int]:
    """General RIFF walk for files that are not in the canonical 44-byte form (18/40-byte fmt chunks,
    WAVE_FORMAT_EXTENSIBLE, `fact`/`LIST` chunks before the data -- what scipy writes for float32 -- and RF64 with
    its `ds64` chunk).  Leaves the file positioned at the payload and returns a synthesised canonical header plus
    the payload size in bytes (which may exceed the header's 32-bit field)."""
    f.seek(12)
    fmt = None
    ds64_data = 0
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
        elif rf64 and cid == b"ds64":
            body = f.read(size + (size & 1))
            if size < 28:
                raise ValueError("truncated ds64 chunk")
            ds64_data = struct.unpack("<QQQ", body[:24])[1]        # riffSize, dataSize, sampleCount
        elif cid == b"data":
            if fmt is None:
                raise ValueError("data chunk before fmt chunk")
            if rf64 and size == _U32:
                size = ds64_data
            tag, ch, rate, brate, align, bits = fmt
            return WAVHeader(sizeOfFile=min(36 + size, _U32), audioFormat=tag, numChannels=ch, sampleRate=rate,
                             byteRate=brate, blockAlign=align, bitsPerSample=bits, dataBytes=min(size, _U32)), size
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
        rf64 = h.riff == b"RF64"
        if (h.riff != b"RIFF" and not rf64) or h.wave != b"WAVE":
            raise ValueError("not a RIFF/WAVE file")
        data_bytes = h.dataBytes
        if rf64 or not (h.fmt == b"fmt " and h.fmtSize == 16 and h.data == b"data"):
            h, data_bytes = _walk_chunks(f, rf64)
        if h.bitsPerSample == 16:
            dtype = np.int16
        elif h.bitsPerSample == 32 and h.audioFormat == IEEE_FLOAT:
            dtype = np.float32
        else:
            raise ValueError(f"unsupported bits per sample: {h.bitsPerSample}")
        samples = np.fromfile(f, dtype=dtype, count=data_bytes // (h.bitsPerSample // 8))
    return h, samples


def write_samples(path: str, header: WAVHeader, samples: np.ndarray, force_rf64: bool = False) -> None:
    """Input header verbatim + raw samples, as the reference's writeSamples does; RF64 (ds64 chunk with the 64-bit
    sizes) once the payload no longer fits the 32-bit fields, or when forced."""
    samples = np.ascontiguousarray(samples)
    payload = samples.nbytes
    with open(path, "wb") as f:
        if not force_rf64 and payload <= _U32 - 36:
            f.write(header.pack())
        else:
            riff_size = 4 + (8 + 28) + (8 + 16) + 8 + payload + (payload & 1)
            frames = payload // header.blockAlign if header.blockAlign else 0
            f.write(b"RF64" + struct.pack("<I", _U32) + b"WAVE")
            f.write(b"ds64" + struct.pack("<IQQQI", 28, riff_size, payload, frames, 0))
            f.write(b"fmt " + struct.pack("<IHHIIHH", 16, header.audioFormat, header.numChannels, header.sampleRate,
                                          header.byteRate, header.blockAlign, header.bitsPerSample))
            f.write(b"data" + struct.pack("<I", _U32))
        samples.tofile(f)

// mavg_wav.h -- canonical-header WAV reader/writer for the drop-in `averager` binaries.
//
// Interface twin of the reference's wav_header.h (WAVHeader :8-24, extractSamples :26-48,
// writeSamples :50-59, extractSamples64 :62-84) so that code written against those names keeps
// compiling, re-implemented for the new library:
//   * one bulk read instead of one ifstream::read per sample;
//   * float32 files (audioFormat 3, 32 bit) are accepted next to int16 -- the north-star extension;
//   * RIFF/WAVE magic is validated, non-canonical layouts (fact/LIST chunks, 18/40-byte fmt, extensible format)
//     are walked chunk by chunk, and the sample count is clamped to the bytes present;
//   * payloads of 4 GiB and more use RF64 (EBU Tech 3306: "RF64" magic, 32-bit sizes set to 0xFFFFFFFF, the real
//     64-bit sizes in a `ds64` chunk) on input and on output -- the reference's uint32 counts stop at 2^32 bytes;
//   * failures are reported through the return value, nothing is printed from library code.
#pragma once

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <exception>
#include <string>
#include <utility>
#include <vector>

#pragma pack(push, 1)
struct WAVHeader {            // 44 bytes, little endian, exactly the on-disk layout
    char     riff[4];         // "RIFF"
    uint32_t sizeOfFile;      // file size - 8
    char     wave[4];         // "WAVE"
    char     fmt[4];          // "fmt "
    uint32_t fmtSize;         // 16 for the canonical header
    uint16_t audioFormat;     // 1 = integer PCM, 3 = IEEE float
    uint16_t numChannels;
    uint32_t sampleRate;
    uint32_t byteRate;
    uint16_t blockAlign;
    uint16_t bitsPerSample;
    char     data[4];         // "data"
    uint32_t dataBytes;       // payload size in bytes
};
#pragma pack(pop)
static_assert(sizeof(WAVHeader) == 44, "canonical WAV header is 44 bytes");

namespace mavg_wav {

enum class SampleKind { Unsupported, Int16, Float32 };

inline SampleKind kind_of(const WAVHeader& h)
{
    if (h.bitsPerSample == 16) return SampleKind::Int16;  // the reference reads any 16-bit file as PCM
    if (h.bitsPerSample == 32 && h.audioFormat == 3) return SampleKind::Float32;
    return SampleKind::Unsupported;
}

inline uint32_t clamp32(uint64_t v) { return v > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)v; }

inline bool magic_ok(const WAVHeader& h)
{
    return !memcmp(h.riff, "RIFF", 4) && !memcmp(h.wave, "WAVE", 4) && !memcmp(h.data, "data", 4);
}

// Reads header + payload.  Returns false (and leaves `bytes` empty) when the file cannot be read.
// Canonical files (44-byte header, the only thing the reference parses) are taken as they are.  Other RIFF/WAVE
// files -- an 18- or 40-byte fmt chunk, WAVE_FORMAT_EXTENSIBLE, `fact` / `LIST` chunks in front of the data (what
// scipy writes for float32) -- are walked chunk by chunk and returned with a synthesised canonical header, so the
// rest of the program (and the output file) only ever sees the 44-byte form.
inline bool read_file(const std::string& path, WAVHeader& h, std::vector<unsigned char>& bytes, std::string* why = nullptr)
{
    bytes.clear();
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) { if (why) *why = "could not open file"; return false; }
    auto fail = [&](const char* msg) { if (why) *why = msg; fclose(f); return false; };
    if (fread(&h, sizeof h, 1, f) != 1) return fail("file shorter than a WAV header");
    const bool rf64 = !memcmp(h.riff, "RF64", 4);
    if ((memcmp(h.riff, "RIFF", 4) && !rf64) || memcmp(h.wave, "WAVE", 4)) return fail("not a RIFF/WAVE file");
    uint64_t data_bytes = 0, ds64_data = 0;
    if (!rf64 && magic_ok(h) && h.fmtSize == 16) {
        data_bytes = h.dataBytes;                              // canonical: payload follows directly
    } else {
        // general RIFF walk from the first chunk (offset 12)
        if (fseek(f, 12, SEEK_SET) != 0) return fail("seek failed");
        bool have_fmt = false, have_data = false;
        uint16_t tag = 0, channels = 0, bits = 0, align = 0;
        uint32_t rate = 0, brate = 0;
        for (;;) {
            char id[4];
            uint32_t size = 0;
            if (fread(id, 4, 1, f) != 1 || fread(&size, 4, 1, f) != 1) break;
            if (!memcmp(id, "fmt ", 4)) {
                unsigned char buf[40] = {0};
                const uint32_t take = size < 40 ? size : 40;
                if (size < 16 || fread(buf, 1, take, f) != take) return fail("truncated fmt chunk");
                memcpy(&tag, buf, 2); memcpy(&channels, buf + 2, 2); memcpy(&rate, buf + 4, 4);
                memcpy(&brate, buf + 8, 4); memcpy(&align, buf + 12, 2); memcpy(&bits, buf + 14, 2);
                if (tag == 0xFFFE && take >= 26) memcpy(&tag, buf + 24, 2);   // extensible: sub-format GUID starts with the tag
                have_fmt = true;
                if (fseek(f, (long)(size - take) + (size & 1), SEEK_CUR) != 0) break;
            } else if (rf64 && !memcmp(id, "ds64", 4)) {
                unsigned char buf[28];
                if (size < 28 || fread(buf, 1, 28, f) != 28) return fail("truncated ds64 chunk");
                memcpy(&ds64_data, buf + 8, 8);                  // riffSize, dataSize, sampleCount, tableLength
                if (fseek(f, (long)(size - 28) + (size & 1), SEEK_CUR) != 0) break;
            } else if (!memcmp(id, "data", 4)) {
                data_bytes = (rf64 && size == 0xFFFFFFFFu) ? ds64_data : size;
                have_data = true;
                break;                                           // payload starts here
            } else if (fseek(f, (long)size + (size & 1), SEEK_CUR) != 0) {
                break;
            }
        }
        if (!have_fmt || !have_data) return fail("no fmt/data chunk found");
        memcpy(h.riff, "RIFF", 4); memcpy(h.fmt, "fmt ", 4); memcpy(h.data, "data", 4);
        h.fmtSize = 16; h.audioFormat = tag; h.numChannels = channels; h.sampleRate = rate; h.byteRate = brate;
        h.blockAlign = align; h.bitsPerSample = bits;
        h.dataBytes = clamp32(data_bytes);                       // callers take the true count from bytes.size()
        h.sizeOfFile = clamp32(36 + data_bytes);
    }
    if (kind_of(h) == SampleKind::Unsupported) {
        if (why) *why = "unsupported bits per sample: " + std::to_string(h.bitsPerSample);
        fclose(f);
        return false;
    }
    // the header's payload size is a claim (0xFFFFFFFF in streamed WAVs, anything in a damaged ds64 chunk): never
    // allocate more than the file holds behind the payload offset
    {
        const long here = ftell(f);
        if (here >= 0 && fseek(f, 0, SEEK_END) == 0) {
            const long end = ftell(f);
            if (end >= here && (uint64_t)(end - here) < data_bytes) data_bytes = (uint64_t)(end - here);
            fseek(f, here, SEEK_SET);
        }
    }
    try {
        bytes.resize(data_bytes);
    } catch (const std::exception&) {
        if (why) *why = "cannot allocate " + std::to_string(data_bytes) + " bytes for the samples";
        fclose(f);
        return false;
    }
    const size_t got = data_bytes ? fread(bytes.data(), 1, data_bytes, f) : 0;
    const size_t step = h.bitsPerSample / 8;
    bytes.resize(got / step * step);  // clamp to whole samples actually present
    if (bytes.size() != data_bytes) { h.dataBytes = clamp32(bytes.size()); h.sizeOfFile = clamp32(36 + (uint64_t)bytes.size()); }
    fclose(f);
    return true;
}

// Header verbatim + samples (wav_header.h:50-59) while the payload fits the 32-bit size fields; RF64 beyond
// (or when forced, for tests): "RF64" 0xFFFFFFFF "WAVE", ds64 {riff size, data size, sample count, 0}, the header's
// fmt fields, "data" 0xFFFFFFFF, samples.
template <typename T>
inline bool write_file(const std::string& path, const WAVHeader& h, const T* samples, size_t count, bool force_rf64 = false)
{
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) return false;
    const uint64_t payload = (uint64_t)count * sizeof(T);
    bool ok;
    if (!force_rf64 && payload <= 0xFFFFFFFFull - 36) {
        ok = fwrite(&h, sizeof h, 1, f) == 1;
    } else {
        const uint32_t ff = 0xFFFFFFFFu, ds_size = 28, fmt_size = 16, table = 0;
        const uint64_t riff_size = 4 + (8 + 28) + (8 + 16) + 8 + payload + (payload & 1);
        const uint64_t frames = h.blockAlign ? payload / h.blockAlign : 0;
        ok = fwrite("RF64", 4, 1, f) == 1 && fwrite(&ff, 4, 1, f) == 1 && fwrite("WAVE", 4, 1, f) == 1 &&
             fwrite("ds64", 4, 1, f) == 1 && fwrite(&ds_size, 4, 1, f) == 1 && fwrite(&riff_size, 8, 1, f) == 1 &&
             fwrite(&payload, 8, 1, f) == 1 && fwrite(&frames, 8, 1, f) == 1 && fwrite(&table, 4, 1, f) == 1 &&
             fwrite("fmt ", 4, 1, f) == 1 && fwrite(&fmt_size, 4, 1, f) == 1 && fwrite(&h.audioFormat, 16, 1, f) == 1 &&
             fwrite("data", 4, 1, f) == 1 && fwrite(&ff, 4, 1, f) == 1;
    }
    ok = ok && (count == 0 || fwrite(samples, sizeof(T), count, f) == count);
    return fclose(f) == 0 && ok;
}

template <typename T>
inline WAVHeader make_header(size_t samples, uint16_t channels, uint32_t rate = 44100)
{
    WAVHeader h;
    memcpy(h.riff, "RIFF", 4); memcpy(h.wave, "WAVE", 4); memcpy(h.fmt, "fmt ", 4); memcpy(h.data, "data", 4);
    h.fmtSize = 16;
    h.audioFormat = sizeof(T) == 4 ? 3 : 1;
    h.numChannels = channels;
    h.sampleRate = rate;
    h.bitsPerSample = (uint16_t)(8 * sizeof(T));
    h.blockAlign = (uint16_t)(channels * sizeof(T));
    h.byteRate = rate * h.blockAlign;
    h.dataBytes = clamp32((uint64_t)samples * sizeof(T));     // 0xFFFFFFFF = "see ds64" (write_file switches to RF64)
    h.sizeOfFile = clamp32(36 + (uint64_t)samples * sizeof(T));
    return h;
}

}  // namespace mavg_wav

// ---- reference-compatible names (wav_header.h:26-84)
inline std::pair<WAVHeader, std::vector<int16_t>> extractSamples(const std::string& pathName)
{
    WAVHeader h{};
    std::vector<unsigned char> raw;
    std::string why;
    if (!mavg_wav::read_file(pathName, h, raw, &why) || mavg_wav::kind_of(h) != mavg_wav::SampleKind::Int16) {
        printf("%s\n", why.empty() ? "unsupported bits per sample" : why.c_str());
        return {};
    }
    std::vector<int16_t> s(raw.size() / 2);
    if (!s.empty()) memcpy(s.data(), raw.data(), s.size() * 2);
    return {h, std::move(s)};
}

inline std::pair<WAVHeader, std::vector<int64_t>> extractSamples64(const std::string& pathName)
{
    auto hs = extractSamples(pathName);
    return {hs.first, std::vector<int64_t>(hs.second.begin(), hs.second.end())};
}

inline std::pair<WAVHeader, std::vector<float>> extractSamplesF32(const std::string& pathName)
{
    WAVHeader h{};
    std::vector<unsigned char> raw;
    std::string why;
    if (!mavg_wav::read_file(pathName, h, raw, &why) || mavg_wav::kind_of(h) != mavg_wav::SampleKind::Float32) {
        printf("%s\n", why.empty() ? "not a float32 WAV" : why.c_str());
        return {};
    }
    std::vector<float> s(raw.size() / 4);
    if (!s.empty()) memcpy(s.data(), raw.data(), s.size() * 4);
    return {h, std::move(s)};
}

template <typename T>
inline void writeSamples(const std::string& name, const WAVHeader header, std::vector<T>& samples)
{
    if (!mavg_wav::write_file(name, header, samples.data(), samples.size())) printf("could not open output file\n");
}

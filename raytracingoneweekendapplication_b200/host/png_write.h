// png_write.h — minimal RGB8 PNG writer (stored/uncompressed deflate blocks).
// Stands in for the reference's stbi_write_png call (Camera.txt:118); the vendored
// stb_image_write.h is outside the hot path and is not part of this repository.
#ifndef RTB200_PNG_WRITE_H
#define RTB200_PNG_WRITE_H
#include <cstdint>
#include <cstdio>
#include <vector>

namespace rtb200 {

inline uint32_t png_crc(const uint8_t* p, size_t n, uint32_t crc = 0xffffffffu) {
    static uint32_t table[256];
    static bool ready = false;
    if (!ready) {
        for (uint32_t i = 0; i < 256; i++) {
            uint32_t c = i;
            for (int k = 0; k < 8; k++) c = (c & 1) ? 0xedb88320u ^ (c >> 1) : c >> 1;
            table[i] = c;
        }
        ready = true;
    }
    for (size_t i = 0; i < n; i++) crc = table[(crc ^ p[i]) & 0xff] ^ (crc >> 8);
    return crc;
}

inline void png_chunk(std::FILE* f, const char* tag, const std::vector<uint8_t>& body) {
    uint8_t len[4] = {(uint8_t)(body.size() >> 24), (uint8_t)(body.size() >> 16), (uint8_t)(body.size() >> 8), (uint8_t)body.size()};
    std::fwrite(len, 1, 4, f);
    std::fwrite(tag, 1, 4, f);
    if (!body.empty()) std::fwrite(body.data(), 1, body.size(), f);
    uint32_t crc = png_crc((const uint8_t*)tag, 4);
    if (!body.empty()) crc = png_crc(body.data(), body.size(), crc);
    crc ^= 0xffffffffu;
    uint8_t c[4] = {(uint8_t)(crc >> 24), (uint8_t)(crc >> 16), (uint8_t)(crc >> 8), (uint8_t)crc};
    std::fwrite(c, 1, 4, f);
}

inline bool write_png_rgb8(const char* path, int w, int h, const uint8_t* rgb) {
    std::FILE* f = std::fopen(path, "wb");
    if (!f) return false;
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', '\r', '\n', 0x1a, '\n'};
    std::fwrite(sig, 1, 8, f);
    std::vector<uint8_t> ihdr = {(uint8_t)(w >> 24), (uint8_t)(w >> 16), (uint8_t)(w >> 8), (uint8_t)w,
                                 (uint8_t)(h >> 24), (uint8_t)(h >> 16), (uint8_t)(h >> 8), (uint8_t)h,
                                 8, 2, 0, 0, 0};
    png_chunk(f, "IHDR", ihdr);
    // raw scanlines, filter byte 0 in front of each
    std::vector<uint8_t> raw;
    raw.reserve((size_t)h * (3 * (size_t)w + 1));
    for (int y = 0; y < h; y++) {
        raw.push_back(0);
        raw.insert(raw.end(), rgb + (size_t)y * w * 3, rgb + (size_t)(y + 1) * w * 3);
    }
    std::vector<uint8_t> z;
    z.push_back(0x78); z.push_back(0x01);
    uint32_t a = 1, b = 0;
    for (size_t pos = 0; pos < raw.size() || pos == 0;) {
        size_t n = raw.size() - pos < 65535 ? raw.size() - pos : 65535;
        bool last = pos + n >= raw.size();
        z.push_back(last ? 1 : 0);
        z.push_back((uint8_t)n); z.push_back((uint8_t)(n >> 8));
        z.push_back((uint8_t)~n); z.push_back((uint8_t)(~n >> 8));
        for (size_t i = 0; i < n; i++) {
            uint8_t v = raw[pos + i];
            z.push_back(v);
            a = (a + v) % 65521u;
            b = (b + a) % 65521u;
        }
        pos += n;
        if (last) break;
    }
    uint32_t adler = (b << 16) | a;
    z.push_back((uint8_t)(adler >> 24)); z.push_back((uint8_t)(adler >> 16)); z.push_back((uint8_t)(adler >> 8)); z.push_back((uint8_t)adler);
    png_chunk(f, "IDAT", z);
    png_chunk(f, "IEND", {});
    std::fclose(f);
    return true;
}

}  // namespace rtb200
#endif

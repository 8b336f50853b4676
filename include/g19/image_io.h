// g19/image_io.h -- PNG and PPM writers for tightly packed or strided RGB888 rows (row 0 = top).
// The reference saves through Qt: Viewer::getImage().save(file, "PNG") (reference gui.h:39-45);
// QImage is not available headless, so the stand-in QImage::save (g19/qimage_min.h) and the
// headless driver use this: a valid PNG made of stored (uncompressed) deflate blocks -- no zlib.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace g19 {
namespace io {

inline uint32_t crc32(const uint8_t* p, size_t n, uint32_t crc = 0) {
    static uint32_t table[256];
    static bool init = false;
    if (!init) {
        for (uint32_t i = 0; i < 256; ++i) {
            uint32_t c = i;
            for (int k = 0; k < 8; ++k) c = (c & 1u) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
            table[i] = c;
        }
        init = true;
    }
    crc = ~crc;
    for (size_t i = 0; i < n; ++i) crc = table[(crc ^ p[i]) & 0xffu] ^ (crc >> 8);
    return ~crc;
}

inline void be32(std::vector<uint8_t>& v, uint32_t x) {
    v.push_back(uint8_t(x >> 24)); v.push_back(uint8_t(x >> 16)); v.push_back(uint8_t(x >> 8)); v.push_back(uint8_t(x));
}

inline void chunk(std::vector<uint8_t>& out, const char type[4], const std::vector<uint8_t>& data) {
    be32(out, uint32_t(data.size()));
    size_t at = out.size();
    out.insert(out.end(), type, type + 4);
    out.insert(out.end(), data.begin(), data.end());
    be32(out, crc32(out.data() + at, out.size() - at));
}

// rows: h scanlines of w*3 bytes, `stride` bytes apart
inline bool write_png(const std::string& path, const uint8_t* rows, int w, int h, size_t stride) {
    if (w < 0 || h < 0) return false;
    std::vector<uint8_t> raw; // filter byte 0 + pixels, per scanline
    raw.reserve((size_t(w) * 3 + 1) * size_t(h));
    for (int y = 0; y < h; ++y) {
        raw.push_back(0);
        raw.insert(raw.end(), rows + size_t(y) * stride, rows + size_t(y) * stride + size_t(w) * 3);
    }
    std::vector<uint8_t> z; // zlib stream of stored blocks
    z.push_back(0x78); z.push_back(0x01);
    uint32_t a = 1, b = 0; // adler32
    size_t pos = 0;
    do {
        size_t n = raw.size() - pos;
        if (n > 65535) n = 65535;
        z.push_back(pos + n == raw.size() ? 1 : 0);
        z.push_back(uint8_t(n)); z.push_back(uint8_t(n >> 8));
        z.push_back(uint8_t(~n)); z.push_back(uint8_t((~n) >> 8));
        z.insert(z.end(), raw.begin() + pos, raw.begin() + pos + n);
        for (size_t i = 0; i < n; ++i) {
            a = (a + raw[pos + i]) % 65521u;
            b = (b + a) % 65521u;
        }
        pos += n;
    } while (pos < raw.size());
    be32(z, (b << 16) | a);
    std::vector<uint8_t> out = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    std::vector<uint8_t> ihdr;
    be32(ihdr, uint32_t(w));
    be32(ihdr, uint32_t(h));
    ihdr.push_back(8); ihdr.push_back(2); ihdr.push_back(0); ihdr.push_back(0); ihdr.push_back(0); // 8-bit RGB
    chunk(out, "IHDR", ihdr);
    chunk(out, "IDAT", z);
    chunk(out, "IEND", {});
    FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) return false;
    bool ok = std::fwrite(out.data(), 1, out.size(), f) == out.size();
    return std::fclose(f) == 0 && ok;
}

inline bool write_ppm(const std::string& path, const uint8_t* rows, int w, int h, size_t stride) {
    FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) return false;
    std::fprintf(f, "P6\n%d %d\n255\n", w, h);
    bool ok = true;
    for (int y = 0; y < h; ++y) ok = ok && std::fwrite(rows + size_t(y) * stride, 1, size_t(w) * 3, f) == size_t(w) * 3;
    return std::fclose(f) == 0 && ok;
}

// format: "PNG" / "PPM" (case-insensitive) or null = by file extension (default PNG)
inline bool save_rgb888(const std::string& path, const char* format, const uint8_t* rows, int w, int h, size_t stride) {
    std::string f = format ? format : "";
    if (f.empty()) {
        size_t dot = path.rfind('.');
        f = dot == std::string::npos ? "png" : path.substr(dot + 1);
    }
    for (char& c : f) c = char(c >= 'A' && c <= 'Z' ? c + 32 : c);
    if (f == "ppm" || f == "pnm") return write_ppm(path, rows, w, h, stride);
    if (f == "png") return write_png(path, rows, w, h, stride);
    return false;
}

} // namespace io
} // namespace g19

// g19/qimage_min.h -- headless stand-in for the slice of QImage that Image
// (include/image.h) needs: RGB888 storage, setPixel/pixel/fill/bits. Used only
// when Qt is not installed (this container, the GPU box); with Qt present
// image.h includes the real <QImage> and the Qt viewer paints it unchanged.
#pragma once
#include <algorithm>
#include <cstdint>
#include <string>
#include <vector>

#include "g19/image_io.h"

typedef uint32_t QRgb;
inline int qRed(QRgb c) { return (c >> 16) & 0xff; }
inline int qGreen(QRgb c) { return (c >> 8) & 0xff; }
inline int qBlue(QRgb c) { return c & 0xff; }
inline QRgb qRgb(int r, int g, int b) { return 0xff000000u | (uint32_t(r & 0xff) << 16) | (uint32_t(g & 0xff) << 8) | uint32_t(b & 0xff); }
namespace Qt { enum GlobalColor { black = 2 }; }

class QImage {
  public:
    enum Format { Format_RGB888 = 13 };
    QImage() : _w(0), _h(0), _stride(0) {}
    QImage(int w, int h, Format) : _w(w), _h(h), _stride((w * 3 + 3) & ~3), _px(size_t(_stride) * size_t(h), 0) {}
    int width() const { return _w; }
    int height() const { return _h; }
    int bytesPerLine() const { return _stride; } // 32-bit aligned scanlines, like Qt
    uint8_t* scanLine(int y) { return _px.data() + size_t(y) * _stride; }
    const uint8_t* constScanLine(int y) const { return _px.data() + size_t(y) * _stride; }
    void setPixel(int x, int y, QRgb c) {
        uint8_t* p = scanLine(y) + 3 * x;
        p[0] = uint8_t(qRed(c)); p[1] = uint8_t(qGreen(c)); p[2] = uint8_t(qBlue(c));
    }
    QRgb pixel(int x, int y) const {
        const uint8_t* p = constScanLine(y) + 3 * x;
        return qRgb(p[0], p[1], p[2]);
    }
    void fill(Qt::GlobalColor) { std::fill(_px.begin(), _px.end(), uint8_t(0)); }
    // QImage::save(fileName, format): "PNG" (what gui.h:41-44 asks for) or "PPM"; null = by extension
    bool save(const std::string& file, const char* format = nullptr) const {
        return g19::io::save_rgb888(file, format, _px.data(), _w, _h, size_t(_stride));
    }
  private:
    int _w, _h, _stride;
    std::vector<uint8_t> _px;
};

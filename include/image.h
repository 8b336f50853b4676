// image.h -- interface-compatible Image (reference include/image.h:7-29): a
// QImage(Format_RGB888) wrapper with `friend class Viewer`, so the reference's
// Qt viewer paints it unchanged. setPixel keeps the reference's truncating
// store; RayTracer::run fills whole scanlines from the engine's RGB888 output.
#pragma once
#include <cstring>
#include <string>
#include "g19/compat.h"

struct Image {
    Image() = delete;
    Image(int width, int height) : _image(width, height, QImage::Format_RGB888) { clear(); }

    int width() const { return _image.width(); }
    int height() const { return _image.height(); }

    void setPixel(int x, int y, glm::dvec3 c) {
        int r = int(255 * c.x), g = int(255 * c.y), b = int(255 * c.z);
        bool ok = r >= 0 && r <= 255 && g >= 0 && g <= 255 && b >= 0 && b <= 255; // QColor validity
        _image.setPixel(x, y, ok ? qRgb(r, g, b) : qRgb(0, 0, 0));
    }
    glm::dvec3 getPixel(int x, int y) const {
        QRgb p = _image.pixel(x, y);
        return glm::dvec3(qRed(p) / 255., qGreen(p) / 255., qBlue(p) / 255.);
    }
    void clear() { _image.fill(Qt::black); }

    // (addition) bulk store of tightly packed RGB888 rows, row 0 = top
    void setRows(const uint8_t* rgb888) {
        const int w = width(), h = height();
        for (int y = 0; y < h; ++y) std::memcpy(_image.scanLine(y), rgb888 + size_t(y) * size_t(w) * 3, size_t(w) * 3);
    }
    const uint8_t* row(int y) const { return _image.constScanLine(y); }
    // (addition) what the reference's "Save as..." does through the viewer: getImage().save(file, "PNG") (gui.h:39-45)
    bool save(const std::string& file, const char* format = nullptr) const { return _image.save(file.c_str(), format); }

  private:
    QImage _image;
    friend class Viewer;
};

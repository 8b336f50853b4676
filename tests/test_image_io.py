"""PNG / PPM writers behind Image::save (include/g19/image_io.h): what the reference's "Save as..."
does through QImage::save(file, "PNG") (reference gui.h:39-45), reproduced headless. CPU only."""
import os
import struct
import subprocess
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SRC = r"""
#include "image.h"
int main(int argc, char** argv) {
    int w = 37, h = 19;                     // odd width: scanlines are padded to 4 bytes, the files are not
    Image img(w, h);
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) img.setPixel(x, y, glm::dvec3(x / 37.0, y / 19.0, ((x * 7 + y * 13) % 256) / 255.0));
    img.setPixel(3, 4, glm::dvec3(2.0, 0.5, 0.5));   // out of range: QColor invalid -> black
    return (img.save(argv[1], "PNG") && img.save(argv[2]) && !img.save(argv[3], "BMP")) ? 0 : 1;
}
"""


def _decode_png(raw):
    assert raw[:8] == b"\x89PNG\r\n\x1a\n"
    pos, idat, w, h = 8, b"", 0, 0
    while pos < len(raw):
        n, typ = struct.unpack(">I4s", raw[pos:pos + 8])
        body = raw[pos + 8:pos + 8 + n]
        crc, = struct.unpack(">I", raw[pos + 8 + n:pos + 12 + n])
        assert zlib.crc32(typ + body) & 0xffffffff == crc, typ
        if typ == b"IHDR":
            w, h, depth, ctype = struct.unpack(">IIBB", body[:10])
            assert (depth, ctype) == (8, 2)
        if typ == b"IDAT":
            idat += body
        pos += 12 + n
    rows = np.frombuffer(zlib.decompress(idat), np.uint8).reshape(h, 1 + 3 * w)
    assert (rows[:, 0] == 0).all()
    return rows[:, 1:].reshape(h, w, 3)


def test_png_and_ppm_round_trip(tmp_path):
    src = tmp_path / "w.cpp"
    src.write_text(SRC)
    exe = str(tmp_path / "w")
    subprocess.run(["g++", "-std=c++14", "-O1", "-DG19_NO_QT", "-I", os.path.join(ROOT, "include"), str(src), "-o", exe,
                    "-L", os.path.join(ROOT, "2019global_b200"), "-l2019global_b200",
                    "-Wl,-rpath," + os.path.join(ROOT, "2019global_b200")], check=True)
    png, ppm, bmp = (str(tmp_path / n) for n in ("a.png", "a.ppm", "a.bmp"))
    subprocess.run([exe, png, ppm, bmp], check=True)
    x, y = np.meshgrid(np.arange(37), np.arange(19))
    exp = np.stack([(255 * (x / 37.0)).astype(int), (255 * (y / 19.0)).astype(int),
                    (255 * (((x * 7 + y * 13) % 256) / 255.0)).astype(int)], -1).astype(np.uint8)
    exp[4, 3] = 0
    got = _decode_png(open(png, "rb").read())
    assert np.array_equal(got, exp)
    raw = open(ppm, "rb").read()
    assert raw.startswith(b"P6\n37 19\n255\n")
    assert np.array_equal(np.frombuffer(raw[len(b"P6\n37 19\n255\n"):], np.uint8).reshape(19, 37, 3), exp)
    try:
        from PIL import Image as PILImage
        assert np.array_equal(np.asarray(PILImage.open(png).convert("RGB")), exp)
    except ImportError:
        pass

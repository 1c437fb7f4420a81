// Image file I/O for the host side (texture decode, PNG output).
// Stands in for the reference's stb wrappers, src/utils/imageio.h:6-11 / imageio.cpp:11-20:
//   load_image(path, w, h, bpp)  -> decoded 8-bit texels, row-major, top row first
//   save_image(path, w, h, bpp, data) -> PNG
// Formats: PNG (8-bit gray / gray+alpha / RGB / RGBA / palette, non-interlaced; inflate via zlib) and
// binary PPM/PGM (P6/P5). JPEG is NOT decoded here: the reference's two JPEG assets ship in assets/
// as PNGs holding the texels exactly as the reference's decoder produced them (tools/make_assets.py),
// which keeps texture lookups bit-identical to the oracle's.
#ifndef ARTB_HOST_IMAGEIO_H
#define ARTB_HOST_IMAGEIO_H

#include <string>
#include <vector>

namespace artb {

class imageio {
public:
    static bool load_image(const std::string& path, std::vector<unsigned char>& data, int& width, int& height, int& bytes_per_pixel);
    static bool save_image(const std::string& path, int width, int height, int bytes_per_pixel, const void* data);
};

} // namespace artb

#endif

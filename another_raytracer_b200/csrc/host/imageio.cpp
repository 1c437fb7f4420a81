#include "imageio.h"

#include <zlib.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <iterator>

namespace artb {

namespace {

std::uint32_t be32(const unsigned char* p) { return (std::uint32_t(p[0]) << 24) | (std::uint32_t(p[1]) << 16) | (std::uint32_t(p[2]) << 8) | p[3]; }

void put_be32(std::vector<unsigned char>& v, std::uint32_t x)
{
    v.push_back((unsigned char)(x >> 24)); v.push_back((unsigned char)(x >> 16));
    v.push_back((unsigned char)(x >> 8)); v.push_back((unsigned char)x);
}

int paeth(int a, int b, int c)
{
    const int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
    if (pa <= pb && pa <= pc) return a;
    return pb <= pc ? b : c;
}

bool load_png(const std::vector<unsigned char>& file, std::vector<unsigned char>& out, int& w, int& h, int& bpp)
{
    static const unsigned char sig[8] = { 0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a };
    if (file.size() < 8 || std::memcmp(file.data(), sig, 8) != 0) return false;
    size_t pos = 8;
    int depth = 0, ctype = 0, interlace = 0;
    std::vector<unsigned char> idat, palette;
    bool seen_ihdr = false;
    while (pos + 12 <= file.size()) {
        const std::uint32_t len = be32(&file[pos]);
        const char* type = reinterpret_cast<const char*>(&file[pos + 4]);
        const unsigned char* body = &file[pos + 8];
        if (pos + 12 + len > file.size()) return false;
        if (!std::memcmp(type, "IHDR", 4)) {
            w = int(be32(body)); h = int(be32(body + 4));
            depth = body[8]; ctype = body[9]; interlace = body[12];
            seen_ihdr = true;
        } else if (!std::memcmp(type, "PLTE", 4)) {
            palette.assign(body, body + len);
        } else if (!std::memcmp(type, "IDAT", 4)) {
            idat.insert(idat.end(), body, body + len);
        } else if (!std::memcmp(type, "IEND", 4)) {
            break;
        }
        pos += 12 + len;
    }
    if (!seen_ihdr || depth != 8 || interlace != 0 || w <= 0 || h <= 0) return false;
    int channels = 0;
    switch (ctype) { case 0: channels = 1; break; case 2: channels = 3; break; case 3: channels = 1; break;
                     case 4: channels = 2; break; case 6: channels = 4; break; default: return false; }
    const size_t stride = size_t(w) * size_t(channels);
    std::vector<unsigned char> raw((stride + 1) * size_t(h));
    uLongf raw_len = uLongf(raw.size());
    if (uncompress(raw.data(), &raw_len, idat.data(), uLong(idat.size())) != Z_OK || raw_len != raw.size()) return false;

    std::vector<unsigned char> img(stride * size_t(h));
    for (int y = 0; y < h; ++y) {
        const unsigned char* src = &raw[(stride + 1) * size_t(y)];
        const int filter = src[0];
        ++src;
        unsigned char* dst = &img[stride * size_t(y)];
        const unsigned char* up = y > 0 ? dst - stride : nullptr;
        for (size_t i = 0; i < stride; ++i) {
            const int a = i >= size_t(channels) ? dst[i - size_t(channels)] : 0;
            const int b = up ? up[i] : 0;
            const int c = (up && i >= size_t(channels)) ? up[i - size_t(channels)] : 0;
            int v = src[i];
            switch (filter) {
                case 0: break;
                case 1: v += a; break;
                case 2: v += b; break;
                case 3: v += (a + b) / 2; break;
                case 4: v += paeth(a, b, c); break;
                default: return false;
            }
            dst[i] = (unsigned char)v;
        }
    }
    if (ctype == 3) {   // expand palette to RGB
        out.resize(size_t(w) * size_t(h) * 3);
        for (size_t i = 0; i < size_t(w) * size_t(h); ++i) {
            const size_t idx = size_t(img[i]) * 3;
            for (int k = 0; k < 3; ++k) out[i * 3 + size_t(k)] = idx + size_t(k) < palette.size() ? palette[idx + size_t(k)] : 0;
        }
        bpp = 3;
    } else {
        out.swap(img);
        bpp = channels;
    }
    return true;
}

bool load_pnm(const std::vector<unsigned char>& file, std::vector<unsigned char>& out, int& w, int& h, int& bpp)
{
    if (file.size() < 3 || file[0] != 'P' || (file[1] != '6' && file[1] != '5')) return false;
    bpp = file[1] == '6' ? 3 : 1;
    size_t pos = 2;
    int vals[3], got = 0;
    while (got < 3 && pos < file.size()) {
        while (pos < file.size() && std::isspace(file[pos])) ++pos;
        if (pos < file.size() && file[pos] == '#') { while (pos < file.size() && file[pos] != '\n') ++pos; continue; }
        int v = 0; bool any = false;
        while (pos < file.size() && std::isdigit(file[pos])) { v = v * 10 + (file[pos] - '0'); ++pos; any = true; }
        if (!any) return false;
        vals[got++] = v;
    }
    if (got < 3 || vals[2] != 255) return false;
    ++pos;   // single whitespace after maxval
    w = vals[0]; h = vals[1];
    const size_t n = size_t(w) * size_t(h) * size_t(bpp);
    if (pos + n > file.size()) return false;
    out.assign(file.begin() + std::ptrdiff_t(pos), file.begin() + std::ptrdiff_t(pos + n));
    return true;
}

void put_chunk(std::vector<unsigned char>& png, const char* type, const std::vector<unsigned char>& body)
{
    put_be32(png, std::uint32_t(body.size()));
    const size_t start = png.size();
    png.insert(png.end(), type, type + 4);
    png.insert(png.end(), body.begin(), body.end());
    const uLong crc = crc32(0L, &png[start], uInt(png.size() - start));
    put_be32(png, std::uint32_t(crc));
}

} // namespace

bool imageio::load_image(const std::string& path, std::vector<unsigned char>& data, int& width, int& height, int& bytes_per_pixel)
{
    std::ifstream f(path, std::ios::binary);
    if (!f) return false;
    std::vector<unsigned char> file((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    if (load_png(file, data, width, height, bytes_per_pixel)) return true;
    if (load_pnm(file, data, width, height, bytes_per_pixel)) return true;
    if (file.size() > 2 && file[0] == 0xff && file[1] == 0xd8)
        std::cerr << "imageio: '" << path << "' is a JPEG; convert it to PNG first (see tools/make_assets.py)\n";
    return false;
}

bool imageio::save_image(const std::string& path, int width, int height, int bytes_per_pixel, const void* data)
{
    if (width <= 0 || height <= 0 || bytes_per_pixel < 1 || bytes_per_pixel > 4 || !data) return false;
    static const int ctype_of[5] = { 0, 0, 4, 2, 6 };
    const size_t stride = size_t(width) * size_t(bytes_per_pixel);
    std::vector<unsigned char> raw((stride + 1) * size_t(height));
    const unsigned char* src = static_cast<const unsigned char*>(data);
    for (int y = 0; y < height; ++y) {
        raw[(stride + 1) * size_t(y)] = 0;   // filter: none
        std::memcpy(&raw[(stride + 1) * size_t(y) + 1], src + stride * size_t(y), stride);
    }
    uLongf zlen = compressBound(uLong(raw.size()));
    std::vector<unsigned char> z(zlen);
    if (compress2(z.data(), &zlen, raw.data(), uLong(raw.size()), 6) != Z_OK) return false;
    z.resize(zlen);

    std::vector<unsigned char> png = { 0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a };
    std::vector<unsigned char> ihdr;
    put_be32(ihdr, std::uint32_t(width)); put_be32(ihdr, std::uint32_t(height));
    ihdr.push_back(8); ihdr.push_back((unsigned char)ctype_of[bytes_per_pixel]);
    ihdr.push_back(0); ihdr.push_back(0); ihdr.push_back(0);
    put_chunk(png, "IHDR", ihdr);
    put_chunk(png, "IDAT", z);
    put_chunk(png, "IEND", {});
    std::ofstream f(path, std::ios::binary);
    f.write(reinterpret_cast<const char*>(png.data()), std::streamsize(png.size()));
    return f.good();
}

} // namespace artb

/*
 * EncodePNGToBuffer / WritePNGFile of the reference's tests/common/write_png_file.hpp, which its
 * GMs call (lots_of_images encodes its textures on the fly) and which the reference implements
 * with libpng. Here: 8-bit RGBA, filter 0 on every row, one zlib stream -- any PNG reader
 * accepts it, the backend's platformDecodeImageTexture included.
 */
#include "common/write_png_file.hpp"

#include <zlib.h>

#include <cstdio>
#include <cstring>

static void put_be32(std::vector<uint8_t>& out, uint32_t v)
{
    out.push_back(static_cast<uint8_t>(v >> 24));
    out.push_back(static_cast<uint8_t>(v >> 16));
    out.push_back(static_cast<uint8_t>(v >> 8));
    out.push_back(static_cast<uint8_t>(v));
}

static void put_chunk(std::vector<uint8_t>& out, const char type[4], const uint8_t* body, size_t len)
{
    put_be32(out, static_cast<uint32_t>(len));
    const size_t start = out.size();
    out.insert(out.end(), type, type + 4);
    out.insert(out.end(), body, body + len);
    put_be32(out, static_cast<uint32_t>(crc32(0, out.data() + start, static_cast<uInt>(4 + len))));
}

std::vector<uint8_t> EncodePNGToBuffer(uint32_t width, uint32_t height, uint8_t* imageDataRGBA, PNGCompression compression)
{
    std::vector<uint8_t> raw;
    raw.reserve((static_cast<size_t>(width) * 4 + 1) * height);
    for (uint32_t y = 0; y < height; ++y)
    {
        raw.push_back(0); // filter: none
        raw.insert(raw.end(), imageDataRGBA + static_cast<size_t>(y) * width * 4, imageDataRGBA + static_cast<size_t>(y + 1) * width * 4);
    }
    uLongf packedLen = compressBound(static_cast<uLong>(raw.size()));
    std::vector<uint8_t> packed(packedLen);
    compress2(packed.data(), &packedLen, raw.data(), static_cast<uLong>(raw.size()), compression == PNGCompression::compact ? 9 : 1);
    std::vector<uint8_t> out = {0x89, 'P', 'N', 'G', '\r', '\n', 0x1a, '\n'};
    uint8_t ihdr[13];
    const uint32_t dims[2] = {width, height};
    for (int k = 0; k < 2; ++k)
        for (int b = 0; b < 4; ++b)
            ihdr[k * 4 + b] = static_cast<uint8_t>(dims[k] >> (24 - 8 * b));
    ihdr[8] = 8;  // bit depth
    ihdr[9] = 6;  // RGBA
    ihdr[10] = ihdr[11] = ihdr[12] = 0;
    put_chunk(out, "IHDR", ihdr, sizeof(ihdr));
    put_chunk(out, "IDAT", packed.data(), packedLen);
    put_chunk(out, "IEND", nullptr, 0);
    return out;
}

void WritePNGFile(uint8_t* pixels, int width, int height, bool flipY, const char* file_name, PNGCompression compression)
{
    std::vector<uint8_t> rows(static_cast<size_t>(width) * height * 4);
    for (int y = 0; y < height; ++y)
        memcpy(rows.data() + static_cast<size_t>(y) * width * 4, pixels + static_cast<size_t>(flipY ? height - 1 - y : y) * width * 4, static_cast<size_t>(width) * 4);
    const std::vector<uint8_t> png = EncodePNGToBuffer(static_cast<uint32_t>(width), static_cast<uint32_t>(height), rows.data(), compression);
    if (FILE* f = fopen(file_name, "wb"))
    {
        fwrite(png.data(), 1, png.size(), f);
        fclose(f);
    }
}

/*
 * png_decode.hpp -- RenderContextCUDAImpl::platformDecodeImageTexture: the backend's own decoder
 * for the encoded images an application hands to Factory::decodeImage (render_context_impl.hpp:43;
 * the reference falls back to its rive_decoders library, which needs libpng / libjpeg / libwebp).
 * PNG only (what the reference's image GMs and most .riv assets embed): 8 / 16-bit grey, grey +
 * alpha, RGB, RGBA and 1..8-bit palette, non-interlaced, inflated with zlib. Output: RGBA8,
 * premultiplied exactly as Bitmap::pixelFormat(RGBAPremul) does (decoders/src/bitmap_decoder.cpp
 * :66-90: (c * a + 128 + ((c * a + 128) >> 8)) >> 8, opaque pixels untouched).
 */
#pragma once

#include <zlib.h>

#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace rivecuda_host
{
inline bool decode_png_rgba_premul(const uint8_t* data, size_t size, uint32_t* outWidth, uint32_t* outHeight, std::vector<uint8_t>* outPixels)
{
    static const uint8_t kMagic[8] = {0x89, 'P', 'N', 'G', '\r', '\n', 0x1a, '\n'};
    if (size < 8 + 25 || memcmp(data, kMagic, 8) != 0)
        return false;
    auto be32 = [](const uint8_t* p) { return (uint32_t(p[0]) << 24) | (uint32_t(p[1]) << 16) | (uint32_t(p[2]) << 8) | uint32_t(p[3]); };
    uint32_t width = 0, height = 0;
    int depth = 0, colorType = 0, interlace = 0;
    std::vector<uint8_t> idat, palette, paletteAlpha;
    int greyKey = -1, keyR = -1, keyG = -1, keyB = -1;
    size_t pos = 8;
    bool sawEnd = false;
    while (pos + 12 <= size && !sawEnd)
    {
        const uint32_t len = be32(data + pos);
        const uint8_t* type = data + pos + 4;
        const uint8_t* body = data + pos + 8;
        if (pos + 12 + static_cast<size_t>(len) > size)
            return false;
        if (memcmp(type, "IHDR", 4) == 0 && len >= 13)
        {
            width = be32(body);
            height = be32(body + 4);
            depth = body[8];
            colorType = body[9];
            interlace = body[12];
        }
        else if (memcmp(type, "PLTE", 4) == 0)
            palette.assign(body, body + len);
        else if (memcmp(type, "tRNS", 4) == 0)
        {
            if (colorType == 3)
                paletteAlpha.assign(body, body + len);
            else if (colorType == 0 && len >= 2)
                greyKey = (body[0] << 8) | body[1];
            else if (colorType == 2 && len >= 6)
            {
                keyR = (body[0] << 8) | body[1];
                keyG = (body[2] << 8) | body[3];
                keyB = (body[4] << 8) | body[5];
            }
        }
        else if (memcmp(type, "IDAT", 4) == 0)
            idat.insert(idat.end(), body, body + len);
        else if (memcmp(type, "IEND", 4) == 0)
            sawEnd = true;
        pos += 12 + static_cast<size_t>(len);
    }
    if (width == 0 || height == 0 || width > 16384 || height > 16384 || interlace != 0)
        return false;
    int channels;
    switch (colorType)
    {
        case 0: channels = 1; break;
        case 2: channels = 3; break;
        case 3: channels = 1; break;
        case 4: channels = 2; break;
        case 6: channels = 4; break;
        default: return false;
    }
    if (!(depth == 8 || depth == 16 || ((colorType == 0 || colorType == 3) && (depth == 1 || depth == 2 || depth == 4))))
        return false;
    if (colorType == 3 && (depth == 16 || palette.size() < 3))
        return false;
    const size_t bitsPerPixel = static_cast<size_t>(channels) * depth;
    const size_t bytesPerPixel = (bitsPerPixel + 7) / 8; // the filters' "previous pixel" distance
    const size_t stride = (static_cast<size_t>(width) * bitsPerPixel + 7) / 8;
    std::vector<uint8_t> raw((stride + 1) * height);
    uLongf rawLen = static_cast<uLongf>(raw.size());
    if (uncompress(raw.data(), &rawLen, idat.data(), static_cast<uLong>(idat.size())) != Z_OK || rawLen != raw.size())
        return false;
    // Undo the row filters in place.
    std::vector<uint8_t> zeros(stride, 0);
    for (uint32_t y = 0; y < height; ++y)
    {
        uint8_t* row = raw.data() + (stride + 1) * y + 1;
        const uint8_t* up = y > 0 ? row - (stride + 1) : zeros.data();
        const int filter = row[-1];
        for (size_t i = 0; i < stride; ++i)
        {
            const int a = i >= bytesPerPixel ? row[i - bytesPerPixel] : 0;
            const int b = up[i];
            const int c = i >= bytesPerPixel ? up[i - bytesPerPixel] : 0;
            int predictor = 0;
            switch (filter)
            {
                case 0: break;
                case 1: predictor = a; break;
                case 2: predictor = b; break;
                case 3: predictor = (a + b) >> 1; break;
                case 4:
                {
                    const int p = a + b - c, pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
                    predictor = (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
                    break;
                }
                default: return false;
            }
            row[i] = static_cast<uint8_t>(row[i] + predictor);
        }
    }
    outPixels->assign(static_cast<size_t>(width) * height * 4, 255);
    for (uint32_t y = 0; y < height; ++y)
    {
        const uint8_t* row = raw.data() + (stride + 1) * y + 1;
        uint8_t* dst = outPixels->data() + static_cast<size_t>(y) * width * 4;
        for (uint32_t x = 0; x < width; ++x, dst += 4)
        {
            auto sample = [&](uint32_t index) -> int { // index-th sample of the row, scaled to 8 bits (16-bit: the value itself)
                if (depth == 8)
                    return row[index];
                if (depth == 16)
                    return (row[index * 2] << 8) | row[index * 2 + 1];
                const uint32_t bit = index * depth;
                return (row[bit >> 3] >> (8 - depth - (bit & 7))) & ((1 << depth) - 1);
            };
            auto to8 = [&](int v) -> uint8_t { return depth == 16 ? static_cast<uint8_t>(v >> 8) : (depth == 8 ? static_cast<uint8_t>(v) : static_cast<uint8_t>(v * 255 / ((1 << depth) - 1))); };
            switch (colorType)
            {
                case 0:
                {
                    const int g = sample(x);
                    dst[0] = dst[1] = dst[2] = to8(g);
                    dst[3] = g == greyKey ? 0 : 255;
                    break;
                }
                case 2:
                {
                    const int r = sample(x * 3), g = sample(x * 3 + 1), b = sample(x * 3 + 2);
                    dst[0] = to8(r);
                    dst[1] = to8(g);
                    dst[2] = to8(b);
                    dst[3] = (r == keyR && g == keyG && b == keyB) ? 0 : 255;
                    break;
                }
                case 3:
                {
                    const size_t idx = static_cast<size_t>(sample(x));
                    if (idx * 3 + 2 < palette.size())
                    {
                        dst[0] = palette[idx * 3];
                        dst[1] = palette[idx * 3 + 1];
                        dst[2] = palette[idx * 3 + 2];
                    }
                    dst[3] = idx < paletteAlpha.size() ? paletteAlpha[idx] : 255;
                    break;
                }
                case 4:
                    dst[0] = dst[1] = dst[2] = to8(sample(x * 2));
                    dst[3] = to8(sample(x * 2 + 1));
                    break;
                default:
                    dst[0] = to8(sample(x * 4));
                    dst[1] = to8(sample(x * 4 + 1));
                    dst[2] = to8(sample(x * 4 + 2));
                    dst[3] = to8(sample(x * 4 + 3));
                    break;
            }
            const uint32_t alpha = dst[3];
            if (alpha != 255)
            {
                for (int k = 0; k < 3; ++k)
                {
                    const uint32_t v = dst[k] * alpha + 128;
                    dst[k] = static_cast<uint8_t>((v + (v >> 8)) >> 8);
                }
            }
        }
    }
    *outWidth = width;
    *outHeight = height;
    return true;
}
} // namespace rivecuda_host

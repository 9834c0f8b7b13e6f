// image.cpp — host image I/O of the compat layer (include/compat/image.h): `Image` with the reference's interface
// (reference include/image.h:16-75, src/image.cpp:14-56) on an own PNG codec (zlib for inflate/deflate/crc32) plus
// NumPy .npy planes.  Cold path: fixture and render-dump I/O either side of the denoise path, never per frame.
//
// PNG support: 8- and 16-bit (high byte kept) gray, gray+alpha, RGB, RGBA and 1/2/4/8-bit palette images,
// non-interlaced — what Blender/Cycles and the reference's stb_image_write produce (render/cornell/1/*.png are 8-bit
// RGB).  Channel conversion follows the convention the reference relies on (src/image.cpp:36: stbi_load with a forced
// channel count): RGB -> RGBA adds A = 255, gray replicates, colour -> gray uses (77 r + 150 g + 29 b) >> 8.
#include "image.h"
#include "extended_math.h"

#include <zlib.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>

namespace {

std::vector<unsigned char> read_file(const std::string& path) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) throw std::runtime_error("Failed to load image '" + path + "': cannot open file");
    std::vector<unsigned char> buf;
    unsigned char chunk[65536];
    size_t n;
    while ((n = fread(chunk, 1, sizeof(chunk), f)) > 0) buf.insert(buf.end(), chunk, chunk + n);
    fclose(f);
    return buf;
}

uint32_t be32(const unsigned char* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

int paeth(int a, int b, int c) {
    const int p = a + b - c, pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

// decodes to 8-bit samples, `channels_out` = channels of the file (1, 2, 3 or 4; palette -> 3 or 4 with tRNS)
std::vector<unsigned char> decode_png(const std::string& path, int& W, int& H, int& channels_out) {
    const std::vector<unsigned char> file = read_file(path);
    static const unsigned char sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    auto fail = [&](const char* why) -> std::runtime_error { return std::runtime_error("Failed to load image '" + path + "': " + why); };
    if (file.size() < 33 || memcmp(file.data(), sig, 8) != 0) throw fail("not a PNG file");
    int depth = 0, ctype = 0, interlace = 0;
    std::vector<unsigned char> idat, plte, trns;
    size_t pos = 8;
    bool have_ihdr = false, done = false;
    while (!done && pos + 12 <= file.size()) {
        const uint32_t len = be32(&file[pos]);
        const unsigned char* type = &file[pos + 4];
        const unsigned char* data = &file[pos + 8];
        if (pos + 12 + (size_t)len > file.size()) throw fail("truncated chunk");
        if (!memcmp(type, "IHDR", 4)) {
            if (len != 13) throw fail("bad IHDR");
            W = (int)be32(data); H = (int)be32(data + 4);
            depth = data[8]; ctype = data[9]; interlace = data[12];
            have_ihdr = true;
        } else if (!memcmp(type, "PLTE", 4)) plte.assign(data, data + len);
        else if (!memcmp(type, "tRNS", 4)) trns.assign(data, data + len);
        else if (!memcmp(type, "IDAT", 4)) idat.insert(idat.end(), data, data + len);
        else if (!memcmp(type, "IEND", 4)) done = true;
        pos += 12 + (size_t)len;
    }
    if (!have_ihdr || W <= 0 || H <= 0) throw fail("missing IHDR");
    if (interlace) throw fail("interlaced PNG not supported");
    int samples;
    switch (ctype) {
        case 0: samples = 1; break;
        case 2: samples = 3; break;
        case 3: samples = 1; break;
        case 4: samples = 2; break;
        case 6: samples = 4; break;
        default: throw fail("bad colour type");
    }
    if (!((depth == 8 || depth == 16) || (ctype == 3 && (depth == 1 || depth == 2 || depth == 4)) ||
          (ctype == 0 && (depth == 1 || depth == 2 || depth == 4))))
        throw fail("unsupported bit depth");
    if (ctype == 3 && depth == 16) throw fail("bad palette depth");
    const size_t bits_pp = (size_t)samples * depth;
    const size_t stride = ((size_t)W * bits_pp + 7) / 8, bpp = bits_pp >= 8 ? bits_pp / 8 : 1;
    std::vector<unsigned char> raw((stride + 1) * (size_t)H);
    uLongf raw_len = (uLongf)raw.size();
    if (uncompress(raw.data(), &raw_len, idat.data(), (uLong)idat.size()) != Z_OK || raw_len != raw.size())
        throw fail("corrupt image data");
    // un-filter in place (PNG spec 9.2)
    std::vector<unsigned char> prev(stride, 0);
    for (int y = 0; y < H; ++y) {
        unsigned char* line = &raw[(stride + 1) * (size_t)y];
        const int ft = line[0];
        unsigned char* cur = line + 1;
        for (size_t i = 0; i < stride; ++i) {
            const int a = i >= bpp ? cur[i - bpp] : 0, b = prev[i], c = i >= bpp ? prev[i - bpp] : 0;
            int v = cur[i];
            switch (ft) {
                case 0: break;
                case 1: v += a; break;
                case 2: v += b; break;
                case 3: v += (a + b) >> 1; break;
                case 4: v += paeth(a, b, c); break;
                default: throw fail("bad filter type");
            }
            cur[i] = (unsigned char)v;
        }
        memcpy(prev.data(), cur, stride);
    }
    // expand to 8-bit samples
    channels_out = ctype == 3 ? (trns.empty() ? 3 : 4) : samples;
    std::vector<unsigned char> out((size_t)W * H * channels_out);
    for (int y = 0; y < H; ++y) {
        const unsigned char* cur = &raw[(stride + 1) * (size_t)y + 1];
        unsigned char* o = &out[(size_t)y * W * channels_out];
        for (int x = 0; x < W; ++x) {
            if (ctype == 3 || (ctype == 0 && depth < 8)) {
                int idx;
                if (depth == 8) idx = cur[x];
                else {
                    const int per = 8 / depth, shift = (per - 1 - x % per) * depth;
                    idx = (cur[x / per] >> shift) & ((1 << depth) - 1);
                }
                if (ctype == 3) {
                    if ((size_t)idx * 3 + 2 >= plte.size()) throw fail("palette index out of range");
                    o[0] = plte[idx * 3]; o[1] = plte[idx * 3 + 1]; o[2] = plte[idx * 3 + 2];
                    if (channels_out == 4) o[3] = (size_t)idx < trns.size() ? trns[idx] : 255;
                } else {
                    o[0] = (unsigned char)(idx * 255 / ((1 << depth) - 1));
                }
            } else {
                for (int s = 0; s < samples; ++s) o[s] = depth == 8 ? cur[(size_t)x * samples + s] : cur[((size_t)x * samples + s) * 2];
            }
            o += channels_out;
        }
    }
    return out;
}

unsigned char luma(int r, int g, int b) { return (unsigned char)((r * 77 + g * 150 + b * 29) >> 8); }

// n-channel samples -> req channels (the conversions a forced channel count implies)
std::vector<unsigned char> convert_channels(const std::vector<unsigned char>& in, size_t px, int n, int req) {
    if (n == req) return in;
    std::vector<unsigned char> out(px * req);
    for (size_t i = 0; i < px; ++i) {
        const unsigned char* s = &in[i * n];
        unsigned char* d = &out[i * req];
        const int r = s[0], g = n >= 3 ? s[1] : s[0], b = n >= 3 ? s[2] : s[0];
        const int a = (n == 2) ? s[1] : (n == 4 ? s[3] : 255);
        switch (req) {
            case 1: d[0] = n >= 3 ? luma(r, g, b) : s[0]; break;
            case 2: d[0] = n >= 3 ? luma(r, g, b) : s[0]; d[1] = (unsigned char)a; break;
            case 3: d[0] = (unsigned char)r; d[1] = (unsigned char)g; d[2] = (unsigned char)b; break;
            default: d[0] = (unsigned char)r; d[1] = (unsigned char)g; d[2] = (unsigned char)b; d[3] = (unsigned char)a; break;
        }
    }
    return out;
}

void put_be32(std::vector<unsigned char>& v, uint32_t x) {
    v.push_back((unsigned char)(x >> 24)); v.push_back((unsigned char)(x >> 16));
    v.push_back((unsigned char)(x >> 8)); v.push_back((unsigned char)x);
}
void put_chunk(std::vector<unsigned char>& file, const char* type, const std::vector<unsigned char>& data) {
    put_be32(file, (uint32_t)data.size());
    const size_t start = file.size();
    file.insert(file.end(), type, type + 4);
    file.insert(file.end(), data.begin(), data.end());
    put_be32(file, (uint32_t)crc32(0L, &file[start], (uInt)(file.size() - start)));
}

void encode_png(const std::string& path, const unsigned char* data, int W, int H, int ch) {
    auto fail = [&](const char* why) { return std::runtime_error("Failed to save image '" + path + "': " + why); };
    if (!data || W <= 0 || H <= 0 || ch < 1 || ch > 4) throw fail("bad image");
    static const int ctype_of[5] = {0, 0, 4, 2, 6};
    std::vector<unsigned char> file = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A}, ihdr;
    put_be32(ihdr, (uint32_t)W); put_be32(ihdr, (uint32_t)H);
    ihdr.push_back(8); ihdr.push_back((unsigned char)ctype_of[ch]); ihdr.push_back(0); ihdr.push_back(0); ihdr.push_back(0);
    put_chunk(file, "IHDR", ihdr);
    const size_t stride = (size_t)W * ch;
    std::vector<unsigned char> raw((stride + 1) * (size_t)H);
    for (int y = 0; y < H; ++y) {
        raw[(stride + 1) * (size_t)y] = 0;  // filter type None
        memcpy(&raw[(stride + 1) * (size_t)y + 1], data + stride * (size_t)y, stride);
    }
    uLongf clen = compressBound((uLong)raw.size());
    std::vector<unsigned char> comp(clen);
    if (compress2(comp.data(), &clen, raw.data(), (uLong)raw.size(), 6) != Z_OK) throw fail("deflate failed");
    comp.resize(clen);
    put_chunk(file, "IDAT", comp);
    put_chunk(file, "IEND", {});
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) throw fail("cannot open file for writing");
    const bool ok = fwrite(file.data(), 1, file.size(), f) == file.size();
    fclose(f);
    if (!ok) throw fail("short write");
}

// ---- NumPy .npy (format 1.0 / 2.0, C order, little endian) --------------------------------------------------------
struct Npy {
    std::string descr;
    std::vector<long> shape;
    size_t offset;
};
Npy parse_npy(const std::vector<unsigned char>& f, const std::string& path) {
    auto fail = [&](const char* why) { return std::runtime_error("Failed to load '" + path + "': " + why); };
    if (f.size() < 12 || memcmp(f.data(), "\x93NUMPY", 6) != 0) throw fail("not a .npy file");
    const int major = f[6];
    size_t hlen, hoff;
    if (major == 1) { hlen = f[8] | (f[9] << 8); hoff = 10; }
    else { hlen = f[8] | (f[9] << 8) | (f[10] << 16) | ((size_t)f[11] << 24); hoff = 12; }
    if (hoff + hlen > f.size()) throw fail("truncated header");
    const std::string h((const char*)&f[hoff], hlen);
    Npy n;
    n.offset = hoff + hlen;
    size_t p = h.find("'descr'");
    if (p == std::string::npos) throw fail("no descr");
    p = h.find('\'', h.find(':', p));
    n.descr = h.substr(p + 1, h.find('\'', p + 1) - p - 1);
    if (h.find("'fortran_order': False") == std::string::npos) throw fail("fortran order not supported");
    p = h.find('(', h.find("'shape'"));
    const size_t e = h.find(')', p);
    const std::string dims = h.substr(p + 1, e - p - 1);
    size_t i = 0;
    while (i < dims.size()) {
        while (i < dims.size() && (dims[i] == ' ' || dims[i] == ',')) ++i;
        if (i >= dims.size()) break;
        n.shape.push_back(strtol(&dims[i], nullptr, 10));
        while (i < dims.size() && dims[i] != ',') ++i;
    }
    return n;
}

}  // namespace

Image::Image() : shape{0, 0, 0}, data(nullptr), owns(false) {}

Image::Image(int3 shape) : shape(shape), data((byte*)malloc((size_t)totalSize(shape))), owns(true) {
    if (!data && totalSize(shape) > 0) throw std::runtime_error("Image: out of memory");
}

Image::Image(byte* data, int3 shape) : shape(shape), data(data), owns(false) {}

Image::Image(std::string filename, int channels) : shape{0, 0, channels}, data(nullptr), owns(true) {
    if (channels < 1 || channels > 4) throw std::runtime_error("Failed to load image '" + filename + "': bad channel count");
    std::vector<unsigned char> px;
    int n = 0;
    if (filename.size() > 4 && filename.compare(filename.size() - 4, 4, ".npy") == 0) {
        const std::vector<unsigned char> f = read_file(filename);
        const Npy h = parse_npy(f, filename);
        if (h.descr != "|u1" && h.descr != "u1") throw std::runtime_error("Failed to load image '" + filename + "': dtype must be uint8");
        if (h.shape.size() < 2 || h.shape.size() > 3) throw std::runtime_error("Failed to load image '" + filename + "': shape must be (H,W[,C])");
        shape.y = (int)h.shape[0]; shape.x = (int)h.shape[1];
        n = h.shape.size() == 3 ? (int)h.shape[2] : 1;
        if (n < 1 || n > 4 || f.size() - h.offset < (size_t)shape.x * shape.y * n)
            throw std::runtime_error("Failed to load image '" + filename + "': truncated data");
        px.assign(f.begin() + h.offset, f.begin() + h.offset + (size_t)shape.x * shape.y * n);
    } else {
        px = decode_png(filename, shape.x, shape.y, n);
    }
    const std::vector<unsigned char> conv = convert_channels(px, (size_t)shape.x * shape.y, n, channels);
    data = (byte*)malloc(conv.size());
    if (!data) throw std::runtime_error("Image: out of memory");
    memcpy(data, conv.data(), conv.size());
}

Image::Image(Image&& o) noexcept : shape(o.shape), data(o.data), owns(o.owns) { o.data = nullptr; o.owns = false; }
Image& Image::operator=(Image&& o) noexcept {
    if (this != &o) {
        if (owns) free(data);
        shape = o.shape; data = o.data; owns = o.owns;
        o.data = nullptr; o.owns = false;
    }
    return *this;
}

Image::~Image() {
    if (owns) free(data);
}

void Image::save(std::string filename) { encode_png(filename, data, shape.x, shape.y, shape.z); }
void Image::save(std::string filename, byte* data, int3 shape) { encode_png(filename, data, shape.x, shape.y, shape.z); }

int3 rmdLoadNpyFloat(const std::string& filename, std::vector<float>& out) {
    const std::vector<unsigned char> f = read_file(filename);
    const Npy h = parse_npy(f, filename);
    if (h.descr != "<f4") throw std::runtime_error("Failed to load '" + filename + "': dtype must be little-endian float32");
    if (h.shape.size() < 2 || h.shape.size() > 3) throw std::runtime_error("Failed to load '" + filename + "': shape must be (H,W[,C])");
    const int H = (int)h.shape[0], W = (int)h.shape[1], C = h.shape.size() == 3 ? (int)h.shape[2] : 1;
    const size_t n = (size_t)W * H * C;
    if (f.size() - h.offset < n * 4) throw std::runtime_error("Failed to load '" + filename + "': truncated data");
    out.resize(n);
    memcpy(out.data(), &f[h.offset], n * 4);
    return make_int3(W, H, C);
}

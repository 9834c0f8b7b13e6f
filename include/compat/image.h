// include/compat/image.h — source-compatible `Image` (reference include/image.h:16-75): a host image, `int3 shape`
// {width, height, channels} + `byte* data`, loaded from / saved to PNG.  The reference decodes with stb_image
// (src/image.cpp:33-40); this implementation (csrc/compat/image.cpp) has its own PNG codec on zlib and additionally
// reads NumPy `.npy` planes (uint8 / float32), which is how depth and motion reach CudaGBuffer::openImages.
// Same conventions: `channels` forces the channel count (an RGB file asked for 4 channels gets A = 255), rows are
// tightly packed, failures throw std::runtime_error (src/image.cpp:38-39).  Unlike the reference (latent double free,
// src/image.cpp:27-31, 54-56) the non-owning constructor does not free.
#pragma once
#ifndef RMD_COMPAT_IMAGE_H
#define RMD_COMPAT_IMAGE_H

#include "utils.h"
#include "vector.h"

#include <cuda_runtime.h>
#include <string>
#include <vector>

struct Image {
    int3 shape;  // {width, height, channels}
    byte* data;

    Image();
    Image(int3 shape);                         // owning, uninitialised
    Image(byte* data, int3 shape);             // non-owning view
    Image(std::string filename, int channels); // .png (8-bit gray / gray+alpha / RGB / RGBA / palette) or .npy (uint8)
    Image(const Image&) = delete;
    Image& operator=(const Image&) = delete;
    Image(Image&& o) noexcept;
    Image& operator=(Image&& o) noexcept;
    ~Image();

    void save(std::string filename);
    static void save(std::string filename, byte* data, int3 shape);

   private:
    bool owns = false;
};

// float32 plane(s) from a NumPy .npy file (C order, little endian, shape (H, W) or (H, W, C)); returns {W, H, C}
int3 rmdLoadNpyFloat(const std::string& filename, std::vector<float>& out);

#endif

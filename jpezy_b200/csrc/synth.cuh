// synth.cuh -- deterministic synthetic planar RGB generator, integer arithmetic only.
// The same function is implemented with numpy in jpezy_b200/synth.py; tests compare the two.
#pragma once
#include "common.cuh"

namespace jz {

__host__ __device__ inline uint32_t synth_hash(uint32_t frame, uint32_t c, uint32_t idx)
{
    uint32_t h = 0x6a70657au ^ (frame * 0x9E3779B1u) ^ (c * 0x85EBCA77u);
    h += idx * 0xC2B2AE3Du;
    h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
    return h;
}
__host__ __device__ inline int synth_tri(uint32_t v, uint32_t P)
{
    const int d = int(v % P) - int(P / 2);
    return d < 0 ? -d : d;
}
__host__ __device__ inline uint8_t synth_pixel(int family, uint32_t W, uint32_t x, uint32_t y, uint32_t frame, uint32_t c)
{
    const uint32_t idx = y * W + x;
    int v;
    if (family == 0) {            // S-photo: low-frequency triangles + +-8 noise
        const uint32_t P1 = 256u + 64u * c, P2 = 192u + 32u * c;
        const int base = 20 + (synth_tri(x + 37u * c + 5u * frame, P1) * 220) / int(P1) +
                         (synth_tri(y + 91u * c + 3u * frame, P2) * 180) / int(P2);
        v = base + int((synth_hash(frame, c, idx) >> 24) % 17u) - 8;
    } else if (family == 1) {     // S-noise: uniform noise
        v = 128 + int((synth_hash(frame, c, idx) >> 24) % 255u) - 127;
    } else {                      // adversarial: flat gray tiles (left half) and integer ramps (right half)
        if (x < W / 2) v = int((((x >> 4) + 3u * (y >> 4) + frame) * 7u) & 255u);
        else v = int((x * (c + 1u) + y) & 255u);
    }
    return uint8_t(v < 0 ? 0 : (v > 255 ? 255 : v));
}

__global__ void __launch_bounds__(256) k_synth(uint8_t* r, uint8_t* g, uint8_t* b, uint32_t W, uint32_t H, uint32_t first_frame,
                                               int family, uint32_t y0)
{
    pdl_wait();
    const size_t npx = size_t(W) * H;
    const size_t img = blockIdx.y;
    const uint32_t frame = first_frame + uint32_t(img);
    for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < npx; i += size_t(gridDim.x) * blockDim.x) {
        const uint32_t yl = uint32_t(i / W), x = uint32_t(i - size_t(yl) * W), y = y0 + yl;   // H rows starting at image row y0
        r[img * npx + i] = synth_pixel(family, W, x, y, frame, 0);
        g[img * npx + i] = synth_pixel(family, W, x, y, frame, 1);
        b[img * npx + i] = synth_pixel(family, W, x, y, frame, 2);
    }
}

}  // namespace jz

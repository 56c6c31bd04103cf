/* synth.h -- deterministic synthetic 4:2:0 clip generator (TEST INFRASTRUCTURE, integer only).
 *
 * Shared by the oracle, the reference shim (oracle/ref_shim.cpp), the tests and bench.py (through
 * liboracle*.so) so that every party sees byte-identical input without shipping YUV files.
 * Follows the recipe of SURVEY.md §8(d): 8x8-blocky random texture under per-band translation
 * (sub-pel motion at half resolution), a scrolling diagonal ramp, per-pixel noise, one hard scene
 * cut at mid clip and a luminance fade after it (so that weighted prediction has work to do).
 * Pure 32-bit integer hashing: no libm, no RNG state, any frame can be generated on its own.
 */
#ifndef X265LA_SYNTH_H
#define X265LA_SYNTH_H

#include <stdint.h>

static inline uint32_t synth_mix(uint32_t a, uint32_t b, uint32_t c, uint32_t d)
{
    uint32_t h = a * 0x9E3779B1u ^ (b + 0x7F4A7C15u) * 0x85EBCA77u ^ (c + 0x165667B1u) * 0xC2B2AE3Du ^ (d * 0x27D4EB2Fu);
    h ^= h >> 15; h *= 0x2C1B3C6Du;
    h ^= h >> 12; h *= 0x297A2D39u;
    h ^= h >> 15;
    return h;
}

/* luma sample in 8-bit range */
static inline int synth_luma8(int x, int y, int t, int w, int h, int nframes, uint32_t seed)
{
    int cut = nframes / 2;
    int scene = (t >= cut && nframes > 3) ? 1 : 0;
    int band = (4 * y) / h;                       /* 4 horizontal bands with their own motion */
    int dx = 3 + band, dy = 2 - band;
    if (scene) { dx = -2 - band; dy = 1 + band; }
    int tx = x + dx * t + 4096, ty = y + dy * t + 4096;
    int tex = (int)(synth_mix((uint32_t)(tx >> 3), (uint32_t)(ty >> 3), (uint32_t)scene, seed) & 255u);
    int r = (x + y + 4 * t) & 511;
    int ramp = r > 255 ? 511 - r : r;
    int noise = (int)(synth_mix((uint32_t)x, (uint32_t)y, (uint32_t)t, seed ^ 0xA5A5A5A5u) & 15u);
    int v = (154 * tex + 77 * ramp) >> 8;
    v += noise;
    if (scene)
    {
        /* fade: gain rises from 0.55 by 1/32 per frame after the cut, capped at 1.0 */
        int g = 70 + 4 * (t - cut);
        if (g > 128) g = 128;
        v = (v * g + 64) >> 7;
    }
    (void)w;
    return v < 0 ? 0 : (v > 255 ? 255 : v);
}

static inline int synth_chroma8(int x, int y, int t, int plane, int nframes, uint32_t seed)
{
    int scene = (t >= nframes / 2 && nframes > 3) ? 1 : 0;
    int tx = x + t + 2048, ty = y + 2048;
    int tex = (int)(synth_mix((uint32_t)(tx >> 2), (uint32_t)(ty >> 2), (uint32_t)(scene * 2 + plane), seed ^ 0x3C3C3C3Cu) & 255u);
    int v = 128 + ((tex - 128) >> 2);
    if (scene && (seed & 0x40000000u))
    {
        /* seeds with bit 30 set: the chroma planes fade with the luma (the explicit weight analysis then has chroma work:
         * mcChroma + the chroma sweeps of weightPrediction.cpp); every other seed keeps the static chroma */
        int g = 70 + 4 * (t - nframes / 2);
        if (g > 128) g = 128;
        v = (v * g + 64) >> 7;
    }
    return v;
}

/* depth 8: planes are uint8_t; depth > 8: planes are uint16_t holding value << (depth - 8).
 * strides are in samples.  u/v may be NULL. */
static inline void synth_frame(int w, int h, int t, int nframes, uint32_t seed, int depth,
                               void* yp, int ystride, void* up, void* vp, int cstride)
{
    int sh = depth - 8;
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++)
        {
            int v = synth_luma8(x, y, t, w, h, nframes, seed);
            if (depth == 8) ((uint8_t*)yp)[(intptr_t)y * ystride + x] = (uint8_t)v;
            else ((uint16_t*)yp)[(intptr_t)y * ystride + x] = (uint16_t)(v << sh);
        }
    if (!up || !vp)
        return;
    for (int y = 0; y < h / 2; y++)
        for (int x = 0; x < w / 2; x++)
        {
            int a = synth_chroma8(x, y, t, 0, nframes, seed);
            int b = synth_chroma8(x, y, t, 1, nframes, seed);
            if (depth == 8)
            {
                ((uint8_t*)up)[(intptr_t)y * cstride + x] = (uint8_t)a;
                ((uint8_t*)vp)[(intptr_t)y * cstride + x] = (uint8_t)b;
            }
            else
            {
                ((uint16_t*)up)[(intptr_t)y * cstride + x] = (uint16_t)(a << sh);
                ((uint16_t*)vp)[(intptr_t)y * cstride + x] = (uint16_t)(b << sh);
            }
        }
}

#endif

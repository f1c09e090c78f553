// pixel_core.cuh -- the per-block and per-pixel arithmetic shared by K1 (idct.cu), K2 (color.cu) and the fused K1+K2 kernel
// (fused.cu): the reference's binary32 AAN network and its YCbCr -> RGB matrix, op for op, never contracted.
#pragma once

#include "common.cuh"

namespace {

// decode.swift:4042-4093  idct8 -- one lane of the reference's SIMD8 network
__device__ __forceinline__ void idct8(float &h0, float &h1, float &h2, float &h3, float &h4, float &h5, float &h6,
                                      float &h7, const float shift)
{
    const float sh0 = fadd(shift, h0);
    const float a0 = fadd(sh0, h4);
    const float a1 = fsub(sh0, h4);
    const float b = fadd(h2, h6);
    const float c = fsub(fmul(1.414213562f, fsub(h2, h6)), b);
    const float r0 = fadd(a0, b), r1 = fadd(a1, c), r2 = fsub(a1, c), r3 = fsub(a0, b);
    const float d0 = fsub(h5, h3), d1 = fadd(h1, h7), d2 = fsub(h1, h7), d3 = fadd(h5, h3);
    const float f = fmul(1.414213562f, fsub(d1, d3));
    const float l = fmul(1.847759065f, fadd(d0, d2));
    const float m0 = fsub(l, fmul(d2, 1.082392200f));
    const float m1 = fsub(l, fmul(d0, 2.613125930f));
    const float s0 = fadd(d1, d3);
    const float s1 = fsub(m1, s0);
    const float s2 = fsub(f, s1);
    const float s3 = fsub(m0, s2);
    h0 = fadd(r0, s0);
    h1 = fadd(r1, s1);
    h2 = fadd(r2, s2);
    h3 = fadd(r3, s3);
    h4 = fsub(r3, s3);
    h5 = fsub(r2, s2);
    h6 = fsub(r1, s1);
    h7 = fsub(r0, s0);
}

// idct8 without the level shift (first pass: shift = 0 in the reference, `0 + h0` is exact)
__device__ __forceinline__ void idct8_noshift(float &h0, float &h1, float &h2, float &h3, float &h4, float &h5,
                                              float &h6, float &h7)
{
    // (0 + h0) == h0 for every h0 except -0.0 -> +0.0, which cannot change any later sum's value
    const float a0 = fadd(h0, h4);
    const float a1 = fsub(h0, h4);
    const float b = fadd(h2, h6);
    const float c = fsub(fmul(1.414213562f, fsub(h2, h6)), b);
    const float r0 = fadd(a0, b), r1 = fadd(a1, c), r2 = fsub(a1, c), r3 = fsub(a0, b);
    const float d0 = fsub(h5, h3), d1 = fadd(h1, h7), d2 = fsub(h1, h7), d3 = fadd(h5, h3);
    const float f = fmul(1.414213562f, fsub(d1, d3));
    const float l = fmul(1.847759065f, fadd(d0, d2));
    const float m0 = fsub(l, fmul(d2, 1.082392200f));
    const float m1 = fsub(l, fmul(d0, 2.613125930f));
    const float s0 = fadd(d1, d3);
    const float s1 = fsub(m1, s0);
    const float s2 = fsub(f, s1);
    const float s3 = fsub(m0, s2);
    h0 = fadd(r0, s0);
    h1 = fadd(r1, s1);
    h2 = fadd(r2, s2);
    h3 = fadd(r3, s3);
    h4 = fsub(r3, s3);
    h5 = fsub(r2, s2);
    h6 = fsub(r1, s1);
    h7 = fsub(r0, s0);
}

// ---- fast paths: 8-bit planes -> RGB8 --------------------------------------------------------------------------------
// jpeg.swift:441-453 with the two zero matrix entries dropped: Y + 0 * d == Y exactly (d finite), so
//   r = Y + 1.402 dr,   g = (Y + -0.34414 db) + -0.71414 dr,   b = Y + 1.772 db     -- same binary32 values.
// clamp(0..255) then truncate == truncate then saturate (F2I.TRUNC + saturating byte pack, as in the IDCT kernel).
__device__ __forceinline__ void ycc_to_rgb_fast(float Y, float db, float dr, float &r, float &g, float &b)
{
    r = fadd(Y, fmul(1.40200f, dr));
    g = fadd(fadd(Y, fmul(-0.34414f, db)), fmul(-0.71414f, dr));
    b = fadd(Y, fmul(1.77200f, db));
}
__device__ __forceinline__ uint32_t byte_of(uint32_t w, int i) { return (w >> (8 * i)) & 0xffu; }

__device__ __forceinline__ float s8_to_float(uint32_t v, int byte) { return (float) (int) (int8_t) (v >> (8 * byte)); }

}  // namespace

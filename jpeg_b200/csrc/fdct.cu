// fdct.cu -- K5: clamp + level shift + 8x8 AAN FDCT + quantise (true division, round half away) + zig-zag store.
//
// Replaces Spectral.Plane.fdct(_:quanta:precision:) (reference encode.swift:199-248; Planar.Plane.load 80-99,
// fdct8 123-188, fdct8x8 191-196).  Same thread-per-block register design as the IDCT kernel: both transposes of
// the reference are register renaming, the zig-zag permutation is resolved at compile time.  The quantised
// integers must be bit-exact, so the arithmetic is the reference's binary32 sequence, never contracted, with an
// IEEE division and roundf (ties away from zero).
#include "common.cuh"

namespace {

constexpr int TILE = 128;

struct FdctParams {
    float    q[64];  // (r[k] * r[h]) * (8 * Q[zz(k,h)]), index h*8+k
    uint32_t total_blocks, units_x, blocks_per_image;
    uint64_t in_image_stride;   // samples
    uint64_t out_image_stride;  // int16 elements
    float    level;             // 2^(P-1) * 8
    float    limit;             // 2^P - 1
    const void *samples;
    int16_t    *coef;
};

// encode.swift:123-188 fdct8, one lane
__device__ __forceinline__ void fdct8(float &g0, float &g1, float &g2, float &g3, float &g4, float &g5, float &g6,
                                      float &g7, const float shift, const bool shifted)
{
    const float a0 = fadd(g0, g7), a1 = fadd(g1, g6), a2 = fadd(g2, g5), a3 = fadd(g3, g4);
    const float b0 = fadd(a0, a3), b1 = fadd(a1, a2), b2 = fsub(a1, a2), b3 = fsub(a0, a3);
    const float c = fmul(0.707106781f, fadd(b2, b3));
    float       r0 = fadd(b0, b1);
    if (shifted) r0 = fsub(r0, shift);  // `- 0` in the second pass is exact
    const float r1 = fadd(b3, c), r2 = fsub(b0, b1), r3 = fsub(b3, c);
    const float d0 = fsub(g3, g4), d1 = fsub(g2, g5), d2 = fsub(g1, g6), d3 = fsub(g0, g7);
    const float f0 = fadd(d0, d1), f1 = fadd(d1, d2), f2 = fadd(d2, d3);
    const float k = fmul(0.707106781f, f1);
    const float l = fmul(0.382683433f, fsub(f0, f2));
    const float m0 = fadd(l, fmul(f0, 0.541196100f));
    const float m1 = fadd(l, fmul(f2, 1.306562965f));
    const float n0 = fadd(d3, k), n1 = fsub(d3, k);
    const float s0 = fadd(n0, m1), s1 = fsub(n1, m0), s2 = fadd(n1, m0), s3 = fsub(n0, m1);
    g0 = r0;
    g1 = s0;
    g2 = r1;
    g3 = s1;
    g4 = r2;
    g5 = s2;
    g6 = r3;
    g7 = s3;
}

template <typename InT>
__global__ void __launch_bounds__(TILE) k_fdct(const __grid_constant__ FdctParams P)
{
    for (uint32_t b = blockIdx.x * TILE + threadIdx.x; b < P.total_blocks; b += gridDim.x * TILE) {
        const uint32_t img = b / P.blocks_per_image;
        const uint32_t rem = b - img * P.blocks_per_image;
        const uint32_t by = rem / P.units_x, bx = rem - by * P.units_x;
        const size_t   w = (size_t) 8 * P.units_x;
        const InT     *src = reinterpret_cast<const InT *>(P.samples) + (size_t) img * P.in_image_stride +
                         (size_t) (8 * by) * w + (size_t) 8 * bx;
        float g[8][8];  // g[y][x]
#pragma unroll
        for (int y = 0; y < 8; ++y) {
            if constexpr (sizeof(InT) == 1) {
                const uint2 v = *reinterpret_cast<const uint2 *>(src + (size_t) y * w);
#pragma unroll
                for (int x = 0; x < 8; ++x) {
                    const uint32_t s = ((x < 4 ? v.x : v.y) >> (8 * (x & 3))) & 0xffu;
                    g[y][x] = fminf(P.limit, (float) s);  // pointwiseMin(limit, ...) encode.swift:86
                }
            } else {
                const uint4    v = *reinterpret_cast<const uint4 *>(src + (size_t) y * w);
                const uint32_t ww[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int x = 0; x < 8; ++x) {
                    const uint32_t s = (ww[x >> 1] >> (16 * (x & 1))) & 0xffffu;
                    g[y][x] = fminf(P.limit, (float) s);
                }
            }
        }
        // pass 1 (horizontal, level shift on the DC output), pass 2 (vertical)
#pragma unroll
        for (int y = 0; y < 8; ++y) fdct8(g[y][0], g[y][1], g[y][2], g[y][3], g[y][4], g[y][5], g[y][6], g[y][7], P.level, true);
#pragma unroll
        for (int u = 0; u < 8; ++u) fdct8(g[0][u], g[1][u], g[2][u], g[3][u], g[4][u], g[5][u], g[6][u], g[7][u], 0.0f, false);
        // quantise + zig-zag: coef[zz(k,h)] = Int16(h[h][k] / q[h][k], rounding: .toNearestOrAwayFromZero)
        uint32_t out[32];
#pragma unroll
        for (int wd = 0; wd < 32; ++wd) out[wd] = 0u;
#pragma unroll
        for (int h = 0; h < 8; ++h)
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int      z = zigzag_index(k, h);  // folds to a constant after unrolling
                const float    v = roundf(__fdiv_rn(g[h][k], P.q[h * 8 + k]));
                const uint32_t c = (uint32_t) __float2int_rz(v) & 0xffffu;
                out[z >> 1] |= (z & 1) ? (c << 16) : c;
            }
        uint4 *dst = reinterpret_cast<uint4 *>(P.coef + (size_t) img * P.out_image_stride + (size_t) rem * 64);
#pragma unroll
        for (int j = 0; j < 8; ++j) dst[j] = make_uint4(out[4 * j], out[4 * j + 1], out[4 * j + 2], out[4 * j + 3]);
    }
}

}  // namespace

int jpeg_fdct_launch(jpeg_sm100_ctx *ctx, const void *d_samples, int sample_bytes, uint64_t in_image_stride,
                     uint32_t n_images, uint32_t ux, uint32_t uy, const float q[64], int precision, int16_t *d_coef,
                     uint64_t out_image_stride)
{
    const uint64_t total = (uint64_t) n_images * ux * uy;
    if (total == 0) return JPEG_SM100_OK;
    if (total > 0x7fffffffull) return JPEG_SM100_ERR_UNSUPPORTED;
    if ((reinterpret_cast<uintptr_t>(d_samples) & 15) || (reinterpret_cast<uintptr_t>(d_coef) & 15) ||
        ((in_image_stride * sample_bytes) & 15) || ((out_image_stride * 2) & 15))
        return JPEG_SM100_ERR_INVALID_ARGUMENT;
    FdctParams P;
    memcpy(P.q, q, sizeof P.q);
    P.total_blocks = (uint32_t) total;
    P.units_x = ux;
    P.blocks_per_image = ux * uy;
    P.in_image_stride = in_image_stride;
    P.out_image_stride = out_image_stride;
    P.level = ldexpf(1.0f, precision - 1) * 8.0f;
    P.limit = ldexpf(1.0f, precision) - 1.0f;
    P.samples = d_samples;
    P.coef = d_coef;
    uint64_t g = (total + TILE - 1) / TILE;
    if (g > (uint64_t) ctx->sm_count * 8) g = (uint64_t) ctx->sm_count * 8;
    if (sample_bytes == 1) k_fdct<uint8_t><<<(uint32_t) g, TILE, 0, ctx->stream>>>(P);
    else k_fdct<uint16_t><<<(uint32_t) g, TILE, 0, ctx->stream>>>(P);
    LAUNCH_CHECK(ctx);
    return JPEG_SM100_OK;
}

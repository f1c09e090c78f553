// spectral_ops.cu -- N3: operations in the coefficient domain, on Spectral planes that stay on the device between the entropy
// decoder (K3) and the entropy encoder (K6/K7).  No IDCT / FDCT round trip, so they are lossless up to the arithmetic the
// reference's own examples spell out:
//   * requantise      examples/recompress/main.swift:35-58   c' = Int16(Double(Int16(q[z]) * c) / Double(q'[z]) + 0.3 * sign)
//   * block transform examples/rotate/main.swift:164-190     dst[offset + M s][z] = src[s][zmap[z]] * mul[z]
//     (Block.transform, main.swift:13-99, builds zmap / mul for the three rotations; any signed permutation works here)
// Both are pure streaming kernels: 2 bytes in + 2 bytes out per coefficient, bound by HBM.
#include "common.cuh"

namespace {

struct RequantParams {
    const int16_t *src;
    int16_t       *dst;
    uint64_t       src_stride, dst_stride;  // int16 elements between consecutive images
    uint64_t       chunks_per_image;        // 8-coefficient chunks per image (8 * blocks)
    uint32_t       n_images;
    uint16_t       q_old[64], q_new[64];
};

// thread = 8 consecutive coefficients (one 16-byte vector) of one block
__global__ void __launch_bounds__(256) k_requantize(const __grid_constant__ RequantParams P)
{
    const uint64_t total = P.chunks_per_image * P.n_images;
    for (uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t) gridDim.x * blockDim.x) {
        const uint32_t img = (uint32_t) (i / P.chunks_per_image);
        const uint64_t c = i - (uint64_t) img * P.chunks_per_image;
        const uint4    v = __ldg(reinterpret_cast<const uint4 *>(P.src + (size_t) img * P.src_stride) + c);
        const int      z0 = (int) (c & 7u) * 8;
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
        uint32_t       o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            uint32_t packed = 0;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int     z = z0 + 2 * k + h;
                const int16_t coef = (int16_t) (w[k] >> (16 * h));
                const int16_t scaled = (int16_t) ((int) (int16_t) P.q_old[z] * (int) coef);  // Int16 * Int16 (main.swift:48)
                const double  r = (double) scaled / (double) P.q_new[z];                      // IEEE double division
                const double  b = r + 0.3 * (r < 0.0 ? -1.0 : 1.0);
                packed |= (uint32_t) (uint16_t) (int16_t) __double2int_rz(b) << (16 * h);       // Int16(_: Double) truncates
            }
            o[k] = packed;
        }
        reinterpret_cast<uint4 *>(P.dst + (size_t) img * P.dst_stride)[c] = make_uint4(o[0], o[1], o[2], o[3]);
    }
}

struct TransformParams {
    const int16_t *src;
    int16_t       *dst;
    uint64_t       src_stride, dst_stride;
    int32_t        sux, suy, dux, duy;  // units of the source / destination plane
    int32_t        xx, xy, yx, yy, ox, oy;
    uint32_t       n_images;
    uint8_t        zmap[64];
    int8_t         mul[64];
};

// one warp per source block; lane l produces destination coefficients 2l and 2l + 1 (one 4-byte store: the warp writes the
// 128-byte destination block in one transaction; its 64 two-byte gathers all hit the one source line)
__global__ void __launch_bounds__(256) k_transform_blocks(const __grid_constant__ TransformParams P)
{
    const uint32_t lane = threadIdx.x & 31u;
    const uint64_t per_image = (uint64_t) P.sux * P.suy;
    const uint64_t total = per_image * P.n_images;
    const uint64_t warps = ((uint64_t) gridDim.x * blockDim.x) >> 5;
    const int      z0 = 2 * (int) lane, z1 = z0 + 1;
    const int      s0 = P.zmap[z0], s1 = P.zmap[z1], m0 = P.mul[z0], m1 = P.mul[z1];
    for (uint64_t b = ((uint64_t) blockIdx.x * blockDim.x + threadIdx.x) >> 5; b < total; b += warps) {
        const uint32_t img = (uint32_t) (b / per_image);
        const uint32_t rem = (uint32_t) (b - (uint64_t) img * per_image);
        const int      sy = (int) (rem / (uint32_t) P.sux), sx = (int) (rem - (uint32_t) sy * (uint32_t) P.sux);
        const int      dx = P.ox + P.xx * sx + P.xy * sy, dy = P.oy + P.yx * sx + P.yy * sy;
        if (dx < 0 || dx >= P.dux || dy < 0 || dy >= P.duy) continue;  // Spectral.Plane's setter drops it (decode.swift:1470)
        const int16_t *s = P.src + (size_t) img * P.src_stride + 64 * (size_t) rem;
        const int16_t  a = (int16_t) ((int) __ldg(s + s0) * m0), c = (int16_t) ((int) __ldg(s + s1) * m1);
        uint32_t      *d = reinterpret_cast<uint32_t *>(P.dst + (size_t) img * P.dst_stride + 64 * ((size_t) P.dux * dy + dx));
        d[lane] = (uint32_t) (uint16_t) a | ((uint32_t) (uint16_t) c << 16);
    }
}

inline uint32_t grid_cap(jpeg_sm100_ctx *ctx, uint64_t threads, int block, int per_sm)
{
    uint64_t g = (threads + block - 1) / block, cap = (uint64_t) ctx->sm_count * per_sm;
    if (g > cap) g = cap;
    return g ? (uint32_t) g : 1u;
}

}  // namespace

// ---- layer B ------------------------------------------------------------------------------------------------------------
JPEG_API int jpeg_sm100_dev_requantize(jpeg_sm100_ctx *ctx, const jpeg_sm100_dev_spectral *in, const uint16_t *q_old,
                                       const uint16_t *q_new, const jpeg_sm100_dev_spectral *out)
{
    if (!ctx || !in || !out || !q_old || !q_new) return JPEG_SM100_ERR_INVALID_ARGUMENT;
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    if (in->n_planes != out->n_planes || in->n_images != out->n_images || in->n_planes > 4) return JPEG_SM100_ERR_INVALID_ARGUMENT;
    for (uint32_t p = 0; p < in->n_planes; ++p) {
        const auto &a = in->plane[p];
        const auto &b = out->plane[p];
        if (a.units_x != b.units_x || a.units_y != b.units_y || a.units_x < 0 || a.units_y < 0) return JPEG_SM100_ERR_INVALID_ARGUMENT;
        const uint64_t blocks = (uint64_t) a.units_x * a.units_y;
        if (blocks == 0 || in->n_images == 0) continue;
        if ((reinterpret_cast<uintptr_t>(a.coef) & 15) || (reinterpret_cast<uintptr_t>(b.coef) & 15) || (a.image_stride & 7) ||
            (b.image_stride & 7))
            return JPEG_SM100_ERR_INVALID_ARGUMENT;
        RequantParams P;
        P.src = a.coef, P.dst = b.coef, P.src_stride = a.image_stride, P.dst_stride = b.image_stride;
        P.chunks_per_image = blocks * 8, P.n_images = in->n_images;
        for (int z = 0; z < 64; ++z) {
            if (q_new[64 * p + z] == 0) return JPEG_SM100_ERR_INVALID_ARGUMENT;
            P.q_old[z] = q_old[64 * p + z], P.q_new[z] = q_new[64 * p + z];
        }
        k_requantize<<<grid_cap(ctx, P.chunks_per_image * P.n_images, 256, 8), 256, 0, ctx->stream>>>(P);
        LAUNCH_CHECK(ctx);
    }
    return JPEG_SM100_OK;
}

JPEG_API int jpeg_sm100_dev_transform_blocks(jpeg_sm100_ctx *ctx, const jpeg_sm100_dev_spectral *in, const int32_t matrix[4],
                                             const uint8_t zmap[64], const int8_t mul[64], const jpeg_sm100_dev_spectral *out)
{
    if (!ctx || !in || !out || !matrix || !zmap || !mul) return JPEG_SM100_ERR_INVALID_ARGUMENT;
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    if (in->n_planes != out->n_planes || in->n_images != out->n_images || in->n_planes > 4) return JPEG_SM100_ERR_INVALID_ARGUMENT;
    for (int z = 0; z < 64; ++z)
        if (zmap[z] > 63) return JPEG_SM100_ERR_INVALID_ARGUMENT;
    for (uint32_t p = 0; p < in->n_planes; ++p) {
        const auto &a = in->plane[p];
        const auto &b = out->plane[p];
        if (a.units_x < 0 || a.units_y < 0 || b.units_x < 0 || b.units_y < 0) return JPEG_SM100_ERR_INVALID_ARGUMENT;
        if (in->n_images == 0 || (uint64_t) b.units_x * b.units_y == 0) continue;
        if ((reinterpret_cast<uintptr_t>(b.coef) & 3) || (b.image_stride & 1)) return JPEG_SM100_ERR_INVALID_ARGUMENT;
        // destination blocks nothing lands on are zero, as in a newly created Spectral (decode.swift:2241-2256)
        CU_TRY(ctx, cudaMemset2DAsync(b.coef, b.image_stride * 2, 0, (size_t) 128 * b.units_x * b.units_y, in->n_images, ctx->stream));
        if ((uint64_t) a.units_x * a.units_y == 0) continue;
        TransformParams P;
        P.src = a.coef, P.dst = b.coef, P.src_stride = a.image_stride, P.dst_stride = b.image_stride;
        P.sux = a.units_x, P.suy = a.units_y, P.dux = b.units_x, P.duy = b.units_y;
        P.xx = matrix[0], P.xy = matrix[1], P.yx = matrix[2], P.yy = matrix[3];
        // examples/rotate/main.swift:168-172: mirrored axes start from the far end of the SOURCE plane
        P.ox = (P.xx < 0 ? P.sux - 1 : 0) + (P.xy < 0 ? P.suy - 1 : 0);
        P.oy = (P.yx < 0 ? P.sux - 1 : 0) + (P.yy < 0 ? P.suy - 1 : 0);
        P.n_images = in->n_images;
        memcpy(P.zmap, zmap, 64);
        memcpy(P.mul, mul, 64);
        const uint64_t threads = (uint64_t) a.units_x * a.units_y * in->n_images * 32;
        k_transform_blocks<<<grid_cap(ctx, threads, 256, 8), 256, 0, ctx->stream>>>(P);
        LAUNCH_CHECK(ctx);
    }
    return JPEG_SM100_OK;
}

// ---- layer A: one plane, host buffers -------------------------------------------------------------------------------------
namespace {
int one_plane(jpeg_sm100_dev_spectral &sp, void *d, uint32_t ux, uint32_t uy)
{
    memset(&sp, 0, sizeof sp);
    sp.n_images = 1, sp.n_planes = 1;
    sp.plane[0].coef = reinterpret_cast<int16_t *>(d);
    sp.plane[0].image_stride = (uint64_t) 64 * ux * uy;
    sp.plane[0].units_x = (int32_t) ux, sp.plane[0].units_y = (int32_t) uy;
    sp.plane[0].factor_x = sp.plane[0].factor_y = 1;
    return JPEG_SM100_OK;
}
}  // namespace

JPEG_API int jpeg_sm100_requantize(jpeg_sm100_ctx *ctx, const int16_t *coef, uint32_t ux, uint32_t uy, const uint16_t q_old[64],
                                   const uint16_t q_new[64], int16_t *out)
{
    if (!ctx || !q_old || !q_new) return JPEG_SM100_ERR_INVALID_ARGUMENT;
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    const size_t bytes = (size_t) 128 * ux * uy;
    if (!bytes) return JPEG_SM100_OK;
    if (!coef || !out) return JPEG_SM100_ERR_INVALID_ARGUMENT;
    void *d_in = nullptr, *d_out = nullptr;
    J_TRY(scratch_reserve(ctx, 0, bytes, &d_in));
    J_TRY(scratch_reserve(ctx, 1, bytes, &d_out));
    CU_TRY(ctx, cudaMemcpyAsync(d_in, coef, bytes, cudaMemcpyHostToDevice, ctx->stream));
    jpeg_sm100_dev_spectral a, b;
    one_plane(a, d_in, ux, uy);
    one_plane(b, d_out, ux, uy);
    J_TRY(jpeg_sm100_dev_requantize(ctx, &a, q_old, q_new, &b));
    CU_TRY(ctx, cudaMemcpyAsync(out, d_out, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return JPEG_SM100_OK;
}

JPEG_API int jpeg_sm100_transform_blocks(jpeg_sm100_ctx *ctx, const int16_t *coef, uint32_t ux, uint32_t uy, const int32_t matrix[4],
                                         const uint8_t zmap[64], const int8_t mul[64], int16_t *out, uint32_t out_ux, uint32_t out_uy)
{
    if (!ctx || !matrix || !zmap || !mul) return JPEG_SM100_ERR_INVALID_ARGUMENT;
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    const size_t in_bytes = (size_t) 128 * ux * uy, out_bytes = (size_t) 128 * out_ux * out_uy;
    if (!out_bytes) return JPEG_SM100_OK;
    if (!out || (in_bytes && !coef)) return JPEG_SM100_ERR_INVALID_ARGUMENT;
    void *d_in = nullptr, *d_out = nullptr;
    J_TRY(scratch_reserve(ctx, 0, in_bytes + 16, &d_in));
    J_TRY(scratch_reserve(ctx, 1, out_bytes, &d_out));
    if (in_bytes) CU_TRY(ctx, cudaMemcpyAsync(d_in, coef, in_bytes, cudaMemcpyHostToDevice, ctx->stream));
    jpeg_sm100_dev_spectral a, b;
    one_plane(a, d_in, ux, uy);
    one_plane(b, d_out, out_ux, out_uy);
    J_TRY(jpeg_sm100_dev_transform_blocks(ctx, &a, matrix, zmap, mul, &b));
    CU_TRY(ctx, cudaMemcpyAsync(out, d_out, out_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return JPEG_SM100_OK;
}

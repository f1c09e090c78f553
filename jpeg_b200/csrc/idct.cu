// idct.cu -- K1: fused de-zigzag + dequantise + 8x8 AAN IDCT + level shift + clamp + store.
//
// Replaces Spectral.Plane.idct(quanta:precision:) (reference decode.swift:4101-4133; modulate 3984-4017,
// load 4020-4039, idct8 4042-4093, idct8x8 4095-4100).
//
// B200 design
//   * HBM-bound: 128 B of int16 coefficients in, 64 B (8-bit) or 128 B (16-bit) of samples out per block.
//   * One THREAD owns one 8x8 block: all 64 values live in registers, so the de-zigzag permutation and both
//     transposes of the reference are pure register renaming -- no shuffles, no shared-memory round trips.
//     (An 8-point butterfly has no use for tensor cores.)
//   * Coefficient tiles (128 blocks = 16 KB, contiguous in the reference layout) are pulled from HBM by TMA
//     (cp.async.bulk.tensor.2d, 128B-swizzled) into a 3-stage mbarrier ring; the swizzle makes the per-thread
//     128-byte row reads bank-conflict free (chunk j of row t lives at chunk j ^ (t & 7)).
//   * Lanes of a warp own horizontally adjacent blocks, so every row store is a contiguous 256 B (u8) /
//     512 B (u16) segment.
//   * Dequantisation multipliers come straight from the constant bank (kernel parameters): one FMUL per
//     coefficient, no loads.
//   * Arithmetic is the reference's binary32 AAN network, op for op, never contracted (-fmad=false +
//     __f*_rn) => results are bit-identical to the Swift path, not just within +-1.
#include "common.cuh"
#include "pixel_core.cuh"

namespace {

constexpr int TILE   = 128;  // blocks (= threads) per tile
constexpr int STAGES = 3;    // TMA ring depth
constexpr int TILE_BYTES = TILE * 128;

struct IdctParams {
    float    q[64];           // modulated quanta, q[h*8+k]
    uint32_t total_blocks;    // n_images * blocks_per_image
    uint32_t units_x;         // blocks per plane row
    uint32_t blocks_per_image;
    uint32_t n_tiles;
    uint64_t out_image_stride;  // samples
    float    level;             // 2^(P-1) + 0.5
    float    limit;             // 2^P - 1
    const int16_t *coef;        // only used by the non-TMA variant
    void    *out;
};

// w[32]: the block's 64 int16 coefficients in zig-zag order, two per word.
template <typename OutT>
__device__ __forceinline__ void transform_and_store(const uint32_t (&w)[32], const IdctParams &P, OutT *dst,
                                                    const size_t row_stride)
{
    float v[8][8];  // v[h][k]
#pragma unroll
    for (int h = 0; h < 8; ++h)
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            constexpr int dummy = 0;
            (void) dummy;
            const int      z = zigzag_index(k, h);
            const uint32_t word = w[z >> 1];
            const short    s = (z & 1) ? (short) (word >> 16) : (short) (word & 0xffffu);
            v[h][k] = fmul(P.q[h * 8 + k], (float) s);  // load(): quanta * row   decode.swift:4037
        }
    // pass 1: for every horizontal frequency k, transform over the vertical frequency h (shift 0)
#pragma unroll
    for (int k = 0; k < 8; ++k)
        idct8_noshift(v[0][k], v[1][k], v[2][k], v[3][k], v[4][k], v[5][k], v[6][k], v[7][k]);
    // pass 2: for every image row y, transform over k with the level shift folded into the DC term
#pragma unroll
    for (int y = 0; y < 8; ++y) {
        idct8(v[y][0], v[y][1], v[y][2], v[y][3], v[y][4], v[y][5], v[y][6], v[y][7], P.level);
        if constexpr (sizeof(OutT) == 1) {
            // trunc(clamp(g, 0, 255)) == clamp(trunc(g), 0, 255): F2I.TRUNC + saturating pack
            uint2 o;
            o.x = pack4_u8_trunc(v[y][0], v[y][1], v[y][2], v[y][3]);
            o.y = pack4_u8_trunc(v[y][4], v[y][5], v[y][6], v[y][7]);
            *reinterpret_cast<uint2 *>(dst + (size_t) y * row_stride) = o;
        } else {
            uint32_t o[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float a = fminf(fmaxf(v[y][2 * j], 0.0f), P.limit);
                const float b = fminf(fmaxf(v[y][2 * j + 1], 0.0f), P.limit);
                o[j] = (uint32_t) __float2int_rz(a) | ((uint32_t) __float2int_rz(b) << 16);
            }
            *reinterpret_cast<uint4 *>(dst + (size_t) y * row_stride) = make_uint4(o[0], o[1], o[2], o[3]);
        }
    }
}

template <typename OutT>
__device__ __forceinline__ OutT *block_destination(const IdctParams &P, uint32_t b, size_t &row_stride)
{
    const uint32_t img = b / P.blocks_per_image;
    const uint32_t rem = b - img * P.blocks_per_image;
    const uint32_t by = rem / P.units_x;
    const uint32_t bx = rem - by * P.units_x;
    row_stride = (size_t) 8 * P.units_x;
    return reinterpret_cast<OutT *>(P.out) + (size_t) img * P.out_image_stride + (size_t) (8 * by) * row_stride +
           (size_t) 8 * bx;
}

// ---- TMA-staged persistent kernel --------------------------------------------------------------------------
template <typename OutT>
__global__ void __launch_bounds__(TILE, 4)
k_idct_tma(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ IdctParams P)
{
    extern __shared__ uint8_t smem_raw[];
    // 128B-swizzled TMA destinations must be 1024-byte aligned
    uint8_t  *tiles = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t *full = reinterpret_cast<uint64_t *>(tiles + STAGES * TILE_BYTES);

    const uint32_t t = threadIdx.x;
    const uint32_t first = blockIdx.x, step = gridDim.x;

    if (t == 0) {
        prefetch_tmap(&tmap);
#pragma unroll
        for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1);
        fence_mbar_init();
        fence_proxy_async();
#pragma unroll
        for (int s = 0; s < STAGES; ++s) {
            const uint32_t tile = first + s * step;
            if (tile < P.n_tiles) {
                mbar_arrive_expect_tx(&full[s], TILE_BYTES);
                tma_load_2d(tiles + s * TILE_BYTES, &tmap, 0, (int) (tile * TILE), &full[s]);
            }
        }
    }
    __syncthreads();

    uint32_t it = 0;
    for (uint32_t tile = first; tile < P.n_tiles; tile += step, ++it) {
        const uint32_t s = it % STAGES;
        const uint32_t parity = (it / STAGES) & 1u;
        mbar_wait(&full[s], parity);

        // this thread's 128-byte row, un-swizzled: logical chunk j sits at physical chunk j ^ (t & 7)
        const uint4 *row = reinterpret_cast<const uint4 *>(tiles + s * TILE_BYTES + t * 128);
        uint32_t     w[32];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const uint4 c = row[j ^ (t & 7)];
            w[4 * j + 0] = c.x;
            w[4 * j + 1] = c.y;
            w[4 * j + 2] = c.z;
            w[4 * j + 3] = c.w;
        }
        __syncthreads();  // everyone has drained stage s: refill it while we compute
        if (t == 0) {
            const uint32_t next = tile + STAGES * step;
            if (next < P.n_tiles) {
                mbar_arrive_expect_tx(&full[s], TILE_BYTES);
                tma_load_2d(tiles + s * TILE_BYTES, &tmap, 0, (int) (next * TILE), &full[s]);
            }
        }
        const uint32_t b = tile * TILE + t;
        if (b < P.total_blocks) {
            size_t row_stride;
            OutT  *dst = block_destination<OutT>(P, b, row_stride);
            transform_and_store<OutT>(w, P, dst, row_stride);
        }
    }
}

// ---- plain-load variant (no TMA): validation / bring-up path, selected with JPEG_SM100_IDCT=ldg -------------
template <typename OutT>
__global__ void __launch_bounds__(TILE, 4) k_idct_ldg(const __grid_constant__ IdctParams P)
{
    for (uint32_t tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x) {
        const uint32_t b = tile * TILE + threadIdx.x;
        if (b >= P.total_blocks) continue;
        const uint4 *src = reinterpret_cast<const uint4 *>(P.coef + (size_t) b * 64);
        uint32_t     w[32];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const uint4 c = __ldg(src + j);
            w[4 * j + 0] = c.x;
            w[4 * j + 1] = c.y;
            w[4 * j + 2] = c.z;
            w[4 * j + 3] = c.w;
        }
        size_t row_stride;
        OutT  *dst = block_destination<OutT>(P, b, row_stride);
        transform_and_store<OutT>(w, P, dst, row_stride);
    }
}

typedef CUresult (*encode_tiled_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

}  // namespace

// cuTensorMapEncodeTiled through the runtime's driver entry point query (also used by fused.cu)
int resolve_encode_tiled(jpeg_sm100_ctx *ctx)
{
    if (ctx->encode_tiled) return JPEG_SM100_OK;
    void                           *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CU_TRY(ctx, cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (qres != cudaDriverEntryPointSuccess || !fn) {
        ctx->last_error = "cuTensorMapEncodeTiled not available from the driver";
        return JPEG_SM100_ERR_CUDA;
    }
    ctx->encode_tiled = fn;
    return JPEG_SM100_OK;
}

namespace {

template <typename OutT>
int launch_idct(jpeg_sm100_ctx *ctx, const int16_t *d_coef, uint32_t n_images, uint32_t ux, uint32_t uy,
                const float q[64], int precision, OutT *d_out, uint64_t out_image_stride)
{
    const uint64_t total = (uint64_t) n_images * ux * uy;
    if (total == 0) return JPEG_SM100_OK;
    if (total > 0x7fffffffu / TILE * (uint64_t) TILE) return JPEG_SM100_ERR_UNSUPPORTED;
    IdctParams P;
    memcpy(P.q, q, sizeof P.q);
    P.total_blocks = (uint32_t) total;
    P.units_x = ux;
    P.blocks_per_image = ux * uy;
    P.n_tiles = (uint32_t) ((total + TILE - 1) / TILE);
    P.out_image_stride = out_image_stride;
    P.level = ldexpf(1.0f, precision - 1) + 0.5f;
    P.limit = ldexpf(1.0f, precision) - 1.0f;
    P.coef = d_coef;
    P.out = d_out;

    static const bool use_ldg = [] {
        const char *e = getenv("JPEG_SM100_IDCT");
        return e && strcmp(e, "ldg") == 0;
    }();
    const uint32_t grid = P.n_tiles < (uint32_t) ctx->sm_count * 4 ? P.n_tiles : (uint32_t) ctx->sm_count * 4;
    if (use_ldg || (reinterpret_cast<uintptr_t>(d_coef) & 15)) {
        k_idct_ldg<OutT><<<grid, TILE, 0, ctx->stream>>>(P);
        LAUNCH_CHECK(ctx);
        return JPEG_SM100_OK;
    }
    J_TRY(resolve_encode_tiled(ctx));
    CUtensorMap      tmap;
    const cuuint64_t gdim[2] = {64, total};
    const cuuint64_t gstride[1] = {128};
    const cuuint32_t box[2] = {64, TILE};
    const cuuint32_t estride[2] = {1, 1};
    CUresult         r = reinterpret_cast<encode_tiled_fn>(ctx->encode_tiled)(
        &tmap, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, const_cast<int16_t *>(d_coef), gdim, gstride, box, estride,
        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        ctx->last_error = "cuTensorMapEncodeTiled failed (" + std::to_string((int) r) + ")";
        return JPEG_SM100_ERR_CUDA;
    }
    const size_t smem = (size_t) STAGES * TILE_BYTES + 1024 + 64;
    // the opt-in is a per-device function attribute: remembered per ctx (one ctx = one device), never in a process-wide static
    if (!ctx->idct_smem_set[sizeof(OutT) - 1]) {
        CU_TRY(ctx, cudaFuncSetAttribute(k_idct_tma<OutT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
        ctx->idct_smem_set[sizeof(OutT) - 1] = true;
    }
    k_idct_tma<OutT><<<grid, TILE, smem, ctx->stream>>>(tmap, P);
    LAUNCH_CHECK(ctx);
    return JPEG_SM100_OK;
}

}  // namespace

// internal entry points used by api.cu
int jpeg_idct_launch_u8(jpeg_sm100_ctx *ctx, const int16_t *d_coef, uint32_t n_images, uint32_t ux, uint32_t uy,
                        const float q[64], uint8_t *d_out, uint64_t out_image_stride)
{
    return launch_idct<uint8_t>(ctx, d_coef, n_images, ux, uy, q, 8, d_out, out_image_stride);
}
int jpeg_idct_launch_u16(jpeg_sm100_ctx *ctx, const int16_t *d_coef, uint32_t n_images, uint32_t ux, uint32_t uy,
                         const float q[64], int precision, uint16_t *d_out, uint64_t out_image_stride)
{
    return launch_idct<uint16_t>(ctx, d_coef, n_images, ux, uy, q, precision, d_out, out_image_stride);
}

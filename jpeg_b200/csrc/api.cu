// api.cu -- the C-ABI of libjpeg_sm100.so (include/jpeg_sm100.h): lifecycle, layer B (device pointers) and
// layer A (host buffers; what the Swift shim binds).  There is no CPU fallback anywhere in this library.
#include <sched.h>

#include <algorithm>

#include "common.cuh"

// kernels' launchers (idct.cu, color.cu, huffman_decode.cu, fdct.cu, encode.cu)
int jpeg_idct_launch_u8(jpeg_sm100_ctx *, const int16_t *, uint32_t, uint32_t, uint32_t, const float[64], uint8_t *, uint64_t);
int jpeg_idct_launch_u16(jpeg_sm100_ctx *, const int16_t *, uint32_t, uint32_t, uint32_t, const float[64], int, uint16_t *, uint64_t);
int jpeg_color_planar_to_rgb8(jpeg_sm100_ctx *, const jpeg_sm100_dev_planar *, uint32_t, uint32_t, int, uint8_t *);
int jpeg_fused_spectral_to_rgb8(jpeg_sm100_ctx *, const jpeg_sm100_dev_spectral *, const uint16_t *, uint32_t, uint32_t, int, uint8_t *);
int jpeg_color_interleave(jpeg_sm100_ctx *, const jpeg_sm100_dev_planar *, uint32_t, uint32_t, int, uint16_t *);
int jpeg_color_unpack(jpeg_sm100_ctx *, const uint16_t *, uint64_t, int, uint8_t *, bool);
int jpeg_color_pack_rgb(jpeg_sm100_ctx *, const uint8_t *, uint64_t, int, uint16_t *);
int jpeg_color_decompose(jpeg_sm100_ctx *, const void *, bool, uint32_t, uint32_t, const jpeg_sm100_dev_planar *);
int jpeg_huffman_decode_scan(jpeg_sm100_ctx *, const jpeg_sm100_scan_desc *, const uint8_t *, const uint64_t *, uint32_t,
                             uint64_t, int, const jpeg_sm100_huff_table *, int, const jpeg_sm100_dev_spectral *, int32_t *);
int jpeg_huffman_validate_tables(const jpeg_sm100_scan_desc *, const jpeg_sm100_huff_table *);
struct LexPlan {
    uint32_t  n_images, tiles_max;
    uint64_t  n_tiles;
    void     *d_images;
    uint32_t *d_counts, *d_split_base, *d_foreign, *d_bad_phase, *d_n_splits;
    uint64_t *d_emit_base;
};
int jpeg_lex_count(jpeg_sm100_ctx *, const uint8_t *, const uint64_t *, const uint64_t *, uint32_t, LexPlan *);
int jpeg_lex_scatter(jpeg_sm100_ctx *, const uint8_t *, const LexPlan *, uint32_t, uint8_t *, uint64_t *, int32_t *);
int jpeg_fdct_launch(jpeg_sm100_ctx *, const void *, int, uint64_t, uint32_t, uint32_t, uint32_t, const float[64], int,
                     int16_t *, uint64_t);
int jpeg_huffman_encode_scan(jpeg_sm100_ctx *, const jpeg_sm100_scan_desc *, const jpeg_sm100_dev_spectral *, uint64_t,
                             jpeg_sm100_huff_table *, uint8_t *, uint64_t, uint64_t *, uint64_t *h_needed);

// =====================================================================================================================
// lifecycle
// =====================================================================================================================
JPEG_API int jpeg_sm100_abi_version(void) { return JPEG_SM100_ABI_VERSION; }

static int create_common(int device, cudaStream_t borrowed, bool borrow, jpeg_sm100_ctx **out)
{
    if (!out) return JPEG_SM100_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) return JPEG_SM100_ERR_CUDA;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return JPEG_SM100_ERR_CUDA;
    if (prop.major != 10) return JPEG_SM100_ERR_CUDA;  // kernels are built for sm_100a only
    if (cudaSetDevice(device) != cudaSuccess) return JPEG_SM100_ERR_CUDA;
    jpeg_sm100_ctx *ctx = new jpeg_sm100_ctx();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    if (borrow) {
        ctx->stream = borrowed;
        ctx->owns_stream = false;
    } else {
        if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
            delete ctx;
            return JPEG_SM100_ERR_CUDA;
        }
        ctx->owns_stream = true;
    }
    *out = ctx;
    return JPEG_SM100_OK;
}
JPEG_API int jpeg_sm100_create(int device, jpeg_sm100_ctx **out) { return create_common(device, nullptr, false, out); }
JPEG_API int jpeg_sm100_create_on_stream(int device, void *stream, jpeg_sm100_ctx **out)
{
    return create_common(device, reinterpret_cast<cudaStream_t>(stream), true, out);
}
JPEG_API void jpeg_sm100_destroy(jpeg_sm100_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (auto &b : ctx->scratch)
        if (b.ptr) cudaFree(b.ptr);
    for (auto &p : ctx->pinned) {
        if (p.ptr) cudaFreeHost(p.ptr);
        if (p.done) cudaEventDestroy(p.done);
    }
    for (auto &e : ctx->events) cudaEventDestroy(e);
    if (ctx->wait_event) cudaEventDestroy(ctx->wait_event);
    if (ctx->copy_in) cudaStreamDestroy(ctx->copy_in);
    if (ctx->copy_out) cudaStreamDestroy(ctx->copy_out);
    if (ctx->owns_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}
JPEG_API int jpeg_sm100_sync(jpeg_sm100_ctx *ctx)
{
    if (!ctx) return JPEG_SM100_ERR_INVALID_ARGUMENT;
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return JPEG_SM100_OK;
}
JPEG_API const char *jpeg_sm100_error_string(int code)
{
    switch (code) {
    case JPEG_SM100_OK: return "ok";
    case JPEG_SM100_ERR_TRUNCATED_ECS: return "truncatedEntropyCodedSegment";
    case JPEG_SM100_ERR_INVALID_COMPOSITE_VALUE: return "invalidCompositeValue";
    case JPEG_SM100_ERR_INVALID_BLOCK_RUN: return "invalidCompositeBlockRun";
    case JPEG_SM100_ERR_UNDEFINED_DC: return "undefinedScanHuffmanDCReference";
    case JPEG_SM100_ERR_UNDEFINED_AC: return "undefinedScanHuffmanACReference";
    case JPEG_SM100_ERR_PRECONDITION: return "precondition failure in the reference";
    case JPEG_SM100_ERR_INVALID_HUFFMAN: return "invalidHuffmanTable";
    case JPEG_SM100_ERR_RESTART_PHASE: return "invalidRestartPhase";
    case JPEG_SM100_ERR_ECS_COUNT: return "restart marker count differs from the batch geometry";
    case JPEG_SM100_ERR_MISSING_INTERVAL: return "missingRestartIntervalSegment";
    case JPEG_SM100_ERR_INVALID_ARGUMENT: return "invalid argument";
    case JPEG_SM100_ERR_UNSUPPORTED: return "unsupported";
    case JPEG_SM100_ERR_NO_MEMORY: return "out of memory / buffer too small";
    case JPEG_SM100_ERR_CUDA: return "CUDA error (no sm_100 device, or a runtime failure)";
    default: return "unknown";
    }
}
JPEG_API const char *jpeg_sm100_last_cuda_error(jpeg_sm100_ctx *ctx) { return ctx ? ctx->last_error.c_str() : ""; }
JPEG_API uint64_t    jpeg_sm100_launch_count(jpeg_sm100_ctx *ctx) { return ctx ? ctx->launches : 0; }

JPEG_API int jpeg_sm100_malloc(jpeg_sm100_ctx *ctx, size_t bytes, void **p)
{
    if (!ctx || !p) return JPEG_SM100_ERR_INVALID_ARGUMENT;
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    CU_TRY(ctx, cudaMalloc(p, bytes ? bytes : 1));
    return JPEG_SM100_OK;
}
JPEG_API int jpeg_sm100_free(jpeg_sm100_ctx *ctx, void *p)
{
    if (!ctx) return JPEG_SM100_ERR_INVALID_ARGUMENT;
    CU_TRY(ctx, cudaFree(p));
    return JPEG_SM100_OK;
}
JPEG_API int jpeg_sm100_malloc_host(jpeg_sm100_ctx *ctx, size_t bytes, void **p)
{
    if (!ctx || !p) return JPEG_SM100_ERR_INVALID_ARGUMENT;
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    CU_TRY(ctx, cudaMallocHost(p, bytes ? bytes : 1));
    return JPEG_SM100_OK;
}
JPEG_API int jpeg_sm100_free_host(jpeg_sm100_ctx *ctx, void *p)
{
    if (!ctx) return JPEG_SM100_ERR_INVALID_ARGUMENT;
    CU_TRY(ctx, cudaFreeHost(p));
    return JPEG_SM100_OK;
}
JPEG_API int jpeg_sm100_upload(jpeg_sm100_ctx *ctx, void *d, const void *h, size_t n)
{
    if (!ctx) return JPEG_SM100_ERR_INVALID_ARGUMENT;
    if (n) CU_TRY(ctx, copy_h2d(ctx, d, h, n, ctx->stream));
    return JPEG_SM100_OK;
}
JPEG_API int jpeg_sm100_download(jpeg_sm100_ctx *ctx, void *h, const void *d, size_t n)
{
    if (!ctx) return JPEG_SM100_ERR_INVALID_ARGUMENT;
    if (n) CU_TRY(ctx, copy_d2h(ctx, h, d, n, ctx->stream));
    return JPEG_SM100_OK;
}
JPEG_API int jpeg_sm100_memset(jpeg_sm100_ctx *ctx, void *d, int v, size_t n)
{
    if (!ctx) return JPEG_SM100_ERR_INVALID_ARGUMENT;
    if (n) CU_TRY(ctx, cudaMemsetAsync(d, v, n, ctx->stream));
    return JPEG_SM100_OK;
}

// =====================================================================================================================
// layer B
// =====================================================================================================================
#define REQUIRE_CTX(ctx)                                                                                             \
    do {                                                                                                             \
        if (!(ctx)) return JPEG_SM100_ERR_INVALID_ARGUMENT;                                                          \
        CU_TRY((ctx), cudaSetDevice((ctx)->device));                                                                 \
    } while (0)

JPEG_API int jpeg_sm100_dev_decode_scan(jpeg_sm100_ctx *ctx, const jpeg_sm100_scan_desc *scan, const uint8_t *d_ecs,
                                        const uint64_t *d_ecs_offsets, uint32_t n_ecs, uint64_t interval, int extend,
                                        const jpeg_sm100_huff_table *tables, int tables_shared,
                                        const jpeg_sm100_dev_spectral *spectral, int32_t *d_status)
{
    REQUIRE_CTX(ctx);
    return jpeg_huffman_decode_scan(ctx, scan, d_ecs, d_ecs_offsets, n_ecs, interval, extend, tables, tables_shared,
                                    spectral, d_status);
}

JPEG_API int jpeg_sm100_dev_idct(jpeg_sm100_ctx *ctx, const jpeg_sm100_dev_spectral *sp, const uint16_t *quanta,
                                 int precision, const jpeg_sm100_dev_planar *pl)
{
    REQUIRE_CTX(ctx);
    if (!sp || !pl || !quanta || sp->n_planes != pl->n_planes || sp->n_images != pl->n_images || sp->n_planes > 4)
        return JPEG_SM100_ERR_INVALID_ARGUMENT;
    if (precision < 1 || precision > 16 || (pl->sample_bytes == 1 && precision != 8)) return JPEG_SM100_ERR_INVALID_ARGUMENT;
    for (uint32_t p = 0; p < sp->n_planes; ++p) {
        const auto &s = sp->plane[p];
        const auto &d = pl->plane[p];
        if (s.units_x != d.units_x || s.units_y != d.units_y || s.units_x < 0 || s.units_y < 0) return JPEG_SM100_ERR_INVALID_ARGUMENT;
        float q[64];
        modulate_quanta(quanta + 64 * p, 0.125f, q);
        const uint64_t blocks = (uint64_t) s.units_x * s.units_y;
        // images contiguous in both buffers => one launch for the whole batch; otherwise one launch per image
        const bool contiguous = sp->n_images == 1 || s.image_stride == blocks * 64;
        const uint32_t launches = contiguous ? 1 : sp->n_images;
        for (uint32_t i = 0; i < launches; ++i) {
            const int16_t *src = s.coef + (size_t) i * s.image_stride;
            const uint32_t n = contiguous ? sp->n_images : 1;
            if (pl->sample_bytes == 1)
                J_TRY(jpeg_idct_launch_u8(ctx, src, n, s.units_x, s.units_y, q,
                                          reinterpret_cast<uint8_t *>(d.samples) + (size_t) i * d.image_stride, d.image_stride));
            else
                J_TRY(jpeg_idct_launch_u16(ctx, src, n, s.units_x, s.units_y, q, precision,
                                           reinterpret_cast<uint16_t *>(d.samples) + (size_t) i * d.image_stride, d.image_stride));
        }
    }
    return JPEG_SM100_OK;
}

JPEG_API int jpeg_sm100_dev_planar_to_rgb8(jpeg_sm100_ctx *ctx, const jpeg_sm100_dev_planar *pl, uint32_t sx, uint32_t sy,
                                           int cosited, uint8_t *d_rgb)
{
    REQUIRE_CTX(ctx);
    return jpeg_color_planar_to_rgb8(ctx, pl, sx, sy, cosited, d_rgb);
}
JPEG_API int jpeg_sm100_dev_interleave(jpeg_sm100_ctx *ctx, const jpeg_sm100_dev_planar *pl, uint32_t sx, uint32_t sy,
                                       int cosited, uint16_t *d_out)
{
    REQUIRE_CTX(ctx);
    return jpeg_color_interleave(ctx, pl, sx, sy, cosited, d_out);
}
JPEG_API int jpeg_sm100_dev_unpack_rgb8(jpeg_sm100_ctx *ctx, const uint16_t *d_il, uint64_t n, int arity, uint8_t *d_rgb)
{
    REQUIRE_CTX(ctx);
    return jpeg_color_unpack(ctx, d_il, n, arity, d_rgb, true);
}
JPEG_API int jpeg_sm100_dev_unpack_ycc8(jpeg_sm100_ctx *ctx, const uint16_t *d_il, uint64_t n, int arity, uint8_t *d_ycc)
{
    REQUIRE_CTX(ctx);
    return jpeg_color_unpack(ctx, d_il, n, arity, d_ycc, false);
}
JPEG_API int jpeg_sm100_dev_rgb8_to_planar(jpeg_sm100_ctx *ctx, const uint8_t *d_rgb, uint32_t sx, uint32_t sy,
                                           const jpeg_sm100_dev_planar *pl)
{
    REQUIRE_CTX(ctx);
    return jpeg_color_decompose(ctx, d_rgb, true, sx, sy, pl);
}
JPEG_API int jpeg_sm100_dev_pack_rgb8(jpeg_sm100_ctx *ctx, const uint8_t *d_rgb, uint64_t n, int arity, uint16_t *d_il)
{
    REQUIRE_CTX(ctx);
    return jpeg_color_pack_rgb(ctx, d_rgb, n, arity, d_il);
}
JPEG_API int jpeg_sm100_dev_decompose(jpeg_sm100_ctx *ctx, const uint16_t *d_il, uint32_t sx, uint32_t sy,
                                      const jpeg_sm100_dev_planar *pl)
{
    REQUIRE_CTX(ctx);
    return jpeg_color_decompose(ctx, d_il, false, sx, sy, pl);
}
JPEG_API int jpeg_sm100_dev_fdct(jpeg_sm100_ctx *ctx, const jpeg_sm100_dev_planar *pl, const uint16_t *quanta, int precision,
                                 const jpeg_sm100_dev_spectral *sp)
{
    REQUIRE_CTX(ctx);
    if (!sp || !pl || !quanta || sp->n_planes != pl->n_planes || sp->n_images != pl->n_images || sp->n_planes > 4)
        return JPEG_SM100_ERR_INVALID_ARGUMENT;
    if (precision < 1 || precision > 16 || (pl->sample_bytes == 1 && precision > 8)) return JPEG_SM100_ERR_INVALID_ARGUMENT;
    for (uint32_t p = 0; p < sp->n_planes; ++p) {
        const auto &d = sp->plane[p];
        const auto &s = pl->plane[p];
        if (s.units_x != d.units_x || s.units_y != d.units_y) return JPEG_SM100_ERR_INVALID_ARGUMENT;
        float q[64];
        modulate_quanta(quanta + 64 * p, 8.0f, q);
        J_TRY(jpeg_fdct_launch(ctx, s.samples, pl->sample_bytes, s.image_stride, sp->n_images, s.units_x, s.units_y, q,
                               precision, d.coef, d.image_stride));
    }
    return JPEG_SM100_OK;
}
JPEG_API int jpeg_sm100_dev_encode_scan(jpeg_sm100_ctx *ctx, const jpeg_sm100_scan_desc *scan,
                                        const jpeg_sm100_dev_spectral *sp, uint64_t interval_mcus,
                                        jpeg_sm100_huff_table *tables_out, uint8_t *d_ecs, uint64_t ecs_image_stride,
                                        uint64_t *d_ecs_len)
{
    REQUIRE_CTX(ctx);
    return jpeg_huffman_encode_scan(ctx, scan, sp, interval_mcus, tables_out, d_ecs, ecs_image_stride, d_ecs_len, nullptr);
}

// =====================================================================================================================
// layer A: host buffers, synchronous.  scratch slots 0..7 belong to this file.
// =====================================================================================================================
namespace {

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct PlaneSet {
    jpeg_sm100_dev_spectral sp;
    size_t                  bytes[4];
};

// uploads n planes of coefficients into one scratch allocation
int upload_spectral(jpeg_sm100_ctx *ctx, int slot, const jpeg_sm100_plane_i16 *planes, uint32_t n, const int32_t *factors,
                    PlaneSet &out)
{
    if (n < 1 || n > 4) return JPEG_SM100_ERR_INVALID_ARGUMENT;
    memset(&out, 0, sizeof out);
    size_t total = 0, off[4];
    for (uint32_t p = 0; p < n; ++p) {
        if (planes[p].units_x < 0 || planes[p].units_y < 0) return JPEG_SM100_ERR_INVALID_ARGUMENT;
        out.bytes[p] = (size_t) 128 * planes[p].units_x * planes[p].units_y;
        off[p] = total;
        total += align_up(out.bytes[p], 1024);
    }
    void *base = nullptr;
    J_TRY(scratch_reserve(ctx, slot, total + 1024, &base));
    out.sp.n_images = 1;
    out.sp.n_planes = n;
    for (uint32_t p = 0; p < n; ++p) {
        out.sp.plane[p].coef = reinterpret_cast<int16_t *>(reinterpret_cast<uint8_t *>(base) + off[p]);
        out.sp.plane[p].image_stride = out.bytes[p] / 2;
        out.sp.plane[p].units_x = planes[p].units_x;
        out.sp.plane[p].units_y = planes[p].units_y;
        out.sp.plane[p].factor_x = factors ? factors[2 * p] : 1;
        out.sp.plane[p].factor_y = factors ? factors[2 * p + 1] : 1;
        if (out.bytes[p]) {
            if (!planes[p].coef) return JPEG_SM100_ERR_INVALID_ARGUMENT;
            CU_TRY(ctx, copy_h2d(ctx, out.sp.plane[p].coef, planes[p].coef, out.bytes[p], ctx->stream));
        }
    }
    return JPEG_SM100_OK;
}

int alloc_planar(jpeg_sm100_ctx *ctx, int slot, const jpeg_sm100_dev_spectral &sp, int sample_bytes,
                 jpeg_sm100_dev_planar &pl)
{
    memset(&pl, 0, sizeof pl);
    pl.n_images = sp.n_images;
    pl.n_planes = sp.n_planes;
    pl.sample_bytes = sample_bytes;
    size_t total = 0, off[4];
    for (uint32_t p = 0; p < sp.n_planes; ++p) {
        off[p] = total;
        total += align_up((size_t) 64 * sp.plane[p].units_x * sp.plane[p].units_y * sample_bytes * sp.n_images, 1024);
    }
    void *base = nullptr;
    J_TRY(scratch_reserve(ctx, slot, total + 1024, &base));
    for (uint32_t p = 0; p < sp.n_planes; ++p) {
        pl.plane[p].samples = reinterpret_cast<uint8_t *>(base) + off[p];
        pl.plane[p].image_stride = (uint64_t) 64 * sp.plane[p].units_x * sp.plane[p].units_y;
        pl.plane[p].units_x = sp.plane[p].units_x;
        pl.plane[p].units_y = sp.plane[p].units_y;
        pl.plane[p].factor_x = sp.plane[p].factor_x;
        pl.plane[p].factor_y = sp.plane[p].factor_y;
    }
    return JPEG_SM100_OK;
}

// coefficients -> RGB8 on the device: the fused kernel where the geometry allows, else K1 into scratch planes (slot 4) and K2
int spectral_to_rgb8_dev(jpeg_sm100_ctx *ctx, const jpeg_sm100_dev_spectral *sp, const uint16_t *quanta, uint32_t sx, uint32_t sy,
                         int cosited, uint8_t *d_rgb)
{
    const int f = jpeg_fused_spectral_to_rgb8(ctx, sp, quanta, sx, sy, cosited, d_rgb);
    if (f <= 0) return f;
    jpeg_sm100_dev_planar pl;
    J_TRY(alloc_planar(ctx, 4, *sp, 1, pl));
    J_TRY(jpeg_sm100_dev_idct(ctx, sp, quanta, 8, &pl));
    return jpeg_color_planar_to_rgb8(ctx, &pl, sx, sy, cosited, d_rgb);
}

}  // namespace

JPEG_API int jpeg_sm100_dev_spectral_to_rgb8(jpeg_sm100_ctx *ctx, const jpeg_sm100_dev_spectral *sp, const uint16_t *quanta,
                                             uint32_t sx, uint32_t sy, int cosited, uint8_t *d_rgb)
{
    REQUIRE_CTX(ctx);
    if (!sp || !quanta || !d_rgb || (sp->n_planes != 1 && sp->n_planes != 3)) return JPEG_SM100_ERR_INVALID_ARGUMENT;
    return spectral_to_rgb8_dev(ctx, sp, quanta, sx, sy, cosited, d_rgb);
}

JPEG_API int jpeg_sm100_decode_scan(jpeg_sm100_ctx *ctx, const jpeg_sm100_scan_desc *scan, const uint8_t *ecs_concat,
                                    const uint64_t *ecs_offsets, uint32_t n_ecs, uint64_t interval, int extend,
                                    const jpeg_sm100_huff_table dc[4], const jpeg_sm100_huff_table ac[4],
                                    jpeg_sm100_plane_i16 *planes, uint32_t n_planes)
{
    REQUIRE_CTX(ctx);
    if (!scan || !planes || !dc || !ac || (n_ecs && (!ecs_offsets))) return JPEG_SM100_ERR_INVALID_ARGUMENT;
    if (n_ecs == 0) return JPEG_SM100_OK;
    // without DRI the reference's stride yields one element: only the first ECS is consumed (decode.swift:3500)
    if (interval == JPEG_SM100_INTERVAL_NONE) n_ecs = 1;
    PlaneSet ps;
    J_TRY(upload_spectral(ctx, 0, planes, n_planes, nullptr, ps));
    const uint64_t total = ecs_offsets[n_ecs];
    void          *d_ecs = nullptr, *d_off = nullptr, *d_status = nullptr;
    J_TRY(scratch_reserve(ctx, 1, total + 64, &d_ecs));
    J_TRY(scratch_reserve(ctx, 2, sizeof(uint64_t) * ((size_t) n_ecs + 1), &d_off));
    J_TRY(scratch_reserve(ctx, 3, 64, &d_status));
    if (total) CU_TRY(ctx, copy_h2d(ctx, d_ecs, ecs_concat, total, ctx->stream));
    CU_TRY(ctx, copy_h2d(ctx, d_off, ecs_offsets, sizeof(uint64_t) * ((size_t) n_ecs + 1), ctx->stream));
    jpeg_sm100_huff_table tables[8];
    memcpy(tables, dc, sizeof(jpeg_sm100_huff_table) * 4);
    memcpy(tables + 4, ac, sizeof(jpeg_sm100_huff_table) * 4);
    ctx->hint_interval_bytes = total / n_ecs;  // how the parallel decoders cut the intervals
    const int k3 = jpeg_huffman_decode_scan(ctx, scan, reinterpret_cast<uint8_t *>(d_ecs), reinterpret_cast<uint64_t *>(d_off),
                                            n_ecs, interval, extend & (JPEG_SM100_SCAN_EXTEND | JPEG_SM100_SCAN_T81), tables, 1, &ps.sp, reinterpret_cast<int32_t *>(d_status));
    ctx->hint_interval_bytes = 0;
    J_TRY(k3);
    int32_t status = 0;
    CU_TRY(ctx, copy_d2h(ctx, &status, d_status, sizeof status, ctx->stream));
    for (uint32_t p = 0; p < n_planes; ++p)
        if (ps.bytes[p])
            CU_TRY(ctx, copy_d2h(ctx, planes[p].coef, ps.sp.plane[p].coef, ps.bytes[p], ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return status;
}

static int idct_host(jpeg_sm100_ctx *ctx, const int16_t *coef, uint32_t ux, uint32_t uy, const uint16_t quanta[64],
                     int precision, void *samples, int sample_bytes)
{
    REQUIRE_CTX(ctx);
    if (!quanta || ((uint64_t) ux * uy && (!coef || !samples))) return JPEG_SM100_ERR_INVALID_ARGUMENT;
    jpeg_sm100_plane_i16 hp = {const_cast<int16_t *>(coef), (int32_t) ux, (int32_t) uy};
    PlaneSet             ps;
    J_TRY(upload_spectral(ctx, 0, &hp, 1, nullptr, ps));
    jpeg_sm100_dev_planar pl;
    J_TRY(alloc_planar(ctx, 4, ps.sp, sample_bytes, pl));
    J_TRY(jpeg_sm100_dev_idct(ctx, &ps.sp, quanta, precision, &pl));
    const size_t bytes = (size_t) 64 * ux * uy * sample_bytes;
    if (bytes) CU_TRY(ctx, copy_d2h(ctx, samples, pl.plane[0].samples, bytes, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return JPEG_SM100_OK;
}
JPEG_API int jpeg_sm100_idct(jpeg_sm100_ctx *ctx, const int16_t *coef, uint32_t ux, uint32_t uy, const uint16_t quanta[64],
                             int precision, uint16_t *samples)
{
    return idct_host(ctx, coef, ux, uy, quanta, precision, samples, 2);
}
JPEG_API int jpeg_sm100_idct_u8(jpeg_sm100_ctx *ctx, const int16_t *coef, uint32_t ux, uint32_t uy,
                                const uint16_t quanta[64], uint8_t *samples)
{
    return idct_host(ctx, coef, ux, uy, quanta, 8, samples, 1);
}

// uploads host uint16 planes; returns a device planar view
static int upload_planar_u16(jpeg_sm100_ctx *ctx, int slot, const jpeg_sm100_plane_u16 *planes, uint32_t n, bool copy,
                             jpeg_sm100_dev_planar &pl)
{
    if (n < 1 || n > 4 || !planes) return JPEG_SM100_ERR_INVALID_ARGUMENT;
    jpeg_sm100_dev_spectral geo;
    memset(&geo, 0, sizeof geo);
    geo.n_images = 1;
    geo.n_planes = n;
    for (uint32_t p = 0; p < n; ++p) {
        if (planes[p].units_x < 0 || planes[p].units_y < 0) return JPEG_SM100_ERR_INVALID_ARGUMENT;
        geo.plane[p].units_x = planes[p].units_x;
        geo.plane[p].units_y = planes[p].units_y;
        geo.plane[p].factor_x = planes[p].factor_x;
        geo.plane[p].factor_y = planes[p].factor_y;
    }
    J_TRY(alloc_planar(ctx, slot, geo, 2, pl));
    if (copy)
        for (uint32_t p = 0; p < n; ++p) {
            const size_t bytes = (size_t) 128 * planes[p].units_x * planes[p].units_y;
            if (bytes) CU_TRY(ctx, copy_h2d(ctx, pl.plane[p].samples, planes[p].samples, bytes, ctx->stream));
        }
    return JPEG_SM100_OK;
}

JPEG_API int jpeg_sm100_interleave(jpeg_sm100_ctx *ctx, const jpeg_sm100_plane_u16 *planes, uint32_t n, uint32_t sx,
                                   uint32_t sy, int cosited, uint16_t *out)
{
    REQUIRE_CTX(ctx);
    jpeg_sm100_dev_planar pl;
    J_TRY(upload_planar_u16(ctx, 4, planes, n, true, pl));
    const size_t bytes = (size_t) sx * sy * n * 2;
    void        *d_out = nullptr;
    J_TRY(scratch_reserve(ctx, 5, bytes + 64, &d_out));
    J_TRY(jpeg_color_interleave(ctx, &pl, sx, sy, cosited, reinterpret_cast<uint16_t *>(d_out)));
    if (bytes) CU_TRY(ctx, copy_d2h(ctx, out, d_out, bytes, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return JPEG_SM100_OK;
}

static int unpack_host(jpeg_sm100_ctx *ctx, const uint16_t *il, uint64_t n_px, int arity, uint8_t *out, bool to_rgb)
{
    REQUIRE_CTX(ctx);
    if (arity != 1 && arity != 3) return JPEG_SM100_ERR_INVALID_ARGUMENT;
    void *d_il = nullptr, *d_out = nullptr;
    J_TRY(scratch_reserve(ctx, 5, n_px * arity * 2 + 64, &d_il));
    J_TRY(scratch_reserve(ctx, 6, n_px * 3 + 64, &d_out));
    if (n_px) CU_TRY(ctx, copy_h2d(ctx, d_il, il, n_px * arity * 2, ctx->stream));
    J_TRY(jpeg_color_unpack(ctx, reinterpret_cast<uint16_t *>(d_il), n_px, arity, reinterpret_cast<uint8_t *>(d_out), to_rgb));
    if (n_px) CU_TRY(ctx, copy_d2h(ctx, out, d_out, n_px * 3, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return JPEG_SM100_OK;
}
JPEG_API int jpeg_sm100_unpack_rgb8(jpeg_sm100_ctx *ctx, const uint16_t *il, uint64_t n, int arity, uint8_t *rgb)
{
    return unpack_host(ctx, il, n, arity, rgb, true);
}
JPEG_API int jpeg_sm100_unpack_ycc8(jpeg_sm100_ctx *ctx, const uint16_t *il, uint64_t n, int arity, uint8_t *ycc)
{
    return unpack_host(ctx, il, n, arity, ycc, false);
}

JPEG_API int jpeg_sm100_spectral_to_rgb8(jpeg_sm100_ctx *ctx, const jpeg_sm100_plane_i16 *planes, uint32_t n,
                                         const uint16_t *quanta, const int32_t *factors, uint32_t sx, uint32_t sy,
                                         int cosited, uint8_t *rgb)
{
    REQUIRE_CTX(ctx);
    if (!planes || !quanta || !factors || (n != 1 && n != 3)) return JPEG_SM100_ERR_INVALID_ARGUMENT;
    PlaneSet ps;
    J_TRY(upload_spectral(ctx, 0, planes, n, factors, ps));
    const size_t bytes = (size_t) sx * sy * 3;
    void        *d_rgb = nullptr;
    J_TRY(scratch_reserve(ctx, 6, bytes + 64, &d_rgb));
    J_TRY(spectral_to_rgb8_dev(ctx, &ps.sp, quanta, sx, sy, cosited, reinterpret_cast<uint8_t *>(d_rgb)));
    if (bytes) CU_TRY(ctx, copy_d2h(ctx, rgb, d_rgb, bytes, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return JPEG_SM100_OK;
}

// batched host-buffer decode.  Entropy decoding is latency-bound (one thread per restart interval), so it runs once
// for the whole batch; the bandwidth-bound back half is cut into chunks whose device->host copies (the e2e limiter:
// 3 bytes of RGB per pixel leave the device) overlap the kernels of the next chunk:
//     stream  : H2D ecs -> memset -> K3 (whole batch) -> per chunk k: K1 -> K2 into RGB buffer (k & 1)
//     copy_out: D2H of chunk k's RGB (waits for its K2; K2 of chunk k + 2 waits for this copy)
namespace {

// shared body of the two batch entry points.  raw_offsets == nullptr: `bytes` are unstuffed ECS bytes and `offsets` the
// n_images * n_ecs + 1 ECS offsets; else `bytes` are raw scan bytes and the GPU lexer produces both.
// On every exit path the caller's buffers must be quiescent: copies queued by earlier groups may still be writing `rgb`
struct StreamDrain {
    jpeg_sm100_ctx *ctx;
    bool            armed = true;
    ~StreamDrain()
    {
        if (!armed) return;
        cudaStreamSynchronize(ctx->stream);
        if (ctx->copy_out) cudaStreamSynchronize(ctx->copy_out);
    }
};

int decode_batch_common(jpeg_sm100_ctx *ctx, const jpeg_sm100_scan_desc *scan, uint32_t n_images, const uint8_t *bytes,
                        const uint64_t *offsets, const uint64_t *raw_offsets, const uint64_t *raw_lengths, uint32_t n_ecs,
                        uint64_t interval, const jpeg_sm100_huff_table *tables_in, int tables_shared, const uint16_t *quanta,
                        uint32_t sx, uint32_t sy, int cosited, uint8_t *rgb, int32_t *status)
{
    if (scan->band_lo != 0 || scan->band_hi != 64 || scan->bit_hi >= 0) return JPEG_SM100_ERR_UNSUPPORTED;
    const uint32_t n_planes = (uint32_t) scan->n_comp;
    if (n_planes != 1 && n_planes != 3) return JPEG_SM100_ERR_UNSUPPORTED;
    int scx = 0, scy = 0;
    for (uint32_t c = 0; c < n_planes; ++c) {
        if (scan->comp[c].plane != (int) c) return JPEG_SM100_ERR_UNSUPPORTED;  // planes in scan order
        scx = scan->comp[c].factor_x > scx ? scan->comp[c].factor_x : scx;
        scy = scan->comp[c].factor_y > scy ? scan->comp[c].factor_y : scy;
    }
    if (scx < 1 || scy < 1) return JPEG_SM100_ERR_INVALID_ARGUMENT;
    if (!ctx->copy_out) CU_TRY(ctx, cudaStreamCreateWithFlags(&ctx->copy_out, cudaStreamNonBlocking));
    // Table errors belong to their image (decode.swift:3186-3203 raises them per file): validate every set before anything is
    // launched; an image with a missing / invalid table decodes with a stand-in set (its output is undefined) and reports its error.
    std::vector<int32_t>               table_err(n_images, 0);
    std::vector<jpeg_sm100_huff_table> patched;
    const jpeg_sm100_huff_table       *tables = tables_in;
    {
        int good = -1, any_bad = 0;
        const uint32_t n_sets = tables_shared ? 1u : n_images;
        for (uint32_t i = 0; i < n_sets; ++i) {
            const int e = jpeg_huffman_validate_tables(scan, tables_in + (size_t) 8 * i);
            if (e == JPEG_SM100_ERR_INVALID_ARGUMENT) return e;
            if (tables_shared) std::fill(table_err.begin(), table_err.end(), e);
            else table_err[i] = e;
            if (e) any_bad = 1;
            else if (good < 0) good = (int) i;
        }
        if (any_bad && good < 0) {  // nothing decodable in the batch
            if (status) for (uint32_t i = 0; i < n_images; ++i) status[i] = table_err[i];
            return table_err[0];
        }
        if (any_bad) {
            patched.assign(tables_in, tables_in + (size_t) 8 * n_sets);
            for (uint32_t i = 0; i < n_sets; ++i)
                if (table_err[i]) std::copy(tables_in + (size_t) 8 * good, tables_in + (size_t) 8 * good + 8, patched.begin() + (size_t) 8 * i);
            tables = patched.data();
        }
    }
    StreamDrain drain{ctx};

    // chunk size: ~100 MB of RGB per chunk, at least 1 image, at most the batch
    const size_t rgb_per_image = (size_t) sx * sy * 3;
    static const size_t chunk_mb = [] { const char *e = getenv("JPEG_SM100_CHUNK_MB"); return e && atoi(e) > 0 ? (size_t) atoi(e) : (size_t) 100; }();
    uint32_t     chunk = (uint32_t) ((chunk_mb << 20) / (rgb_per_image ? rgb_per_image : 1));
    chunk = chunk < 1 ? 1 : (chunk > n_images ? n_images : chunk);
    const uint32_t n_chunks = (n_images + chunk - 1) / chunk + 4;  // (+ the partial chunk each group of images may end with)
    while (ctx->events.size() < (size_t) 2 * n_chunks) {
        cudaEvent_t ev;
        CU_TRY(ctx, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        ctx->events.push_back(ev);
    }
    cudaEvent_t *ev_k = ctx->events.data(), *ev_out = ev_k + n_chunks;

    jpeg_sm100_dev_spectral sp;
    memset(&sp, 0, sizeof sp);
    sp.n_images = n_images;
    sp.n_planes = n_planes;
    size_t coef_total = 0, coef_off[4];
    for (uint32_t p = 0; p < n_planes; ++p) {
        const int fx = scan->comp[p].factor_x, fy = scan->comp[p].factor_y;
        sp.plane[p].units_x = units_of((int) sx * fx, 8 * scx);
        sp.plane[p].units_y = units_of((int) sy * fy, 8 * scy);
        sp.plane[p].factor_x = fx;
        sp.plane[p].factor_y = fy;
        sp.plane[p].image_stride = (uint64_t) 64 * sp.plane[p].units_x * sp.plane[p].units_y;
        coef_off[p] = coef_total;
        coef_total += align_up((size_t) sp.plane[p].image_stride * 2 * n_images, 1024);
    }
    void *d_coef = nullptr, *d_ecs = nullptr, *d_off = nullptr, *d_status = nullptr, *d_rgb = nullptr, *d_raw = nullptr;
    J_TRY(scratch_reserve(ctx, 0, coef_total + 1024, &d_coef));
    for (uint32_t p = 0; p < n_planes; ++p) sp.plane[p].coef = reinterpret_cast<int16_t *>(reinterpret_cast<uint8_t *>(d_coef) + coef_off[p]);
    const uint64_t n_off = (uint64_t) n_images * n_ecs + 1;
    uint64_t       in_bytes = 0;
    if (raw_offsets)
        for (uint32_t i = 0; i < n_images; ++i) in_bytes = raw_offsets[i] + raw_lengths[i] > in_bytes ? raw_offsets[i] + raw_lengths[i] : in_bytes;
    else
        in_bytes = offsets[n_off - 1];
    // The batch is cut into groups of images that go through the whole pipeline one after the other on the stream, so that the
    // first RGB leaves the device after one group's upload + entropy decode instead of the whole batch's: the device -> host
    // copy of 3 bytes per pixel is what bounds this call, and it should start early and never pause.
    uint32_t group = (n_images + 3) / 4;
    group = group < 8 ? 8 : group;
    group = (group + chunk - 1) / chunk * chunk;  // whole chunks
    group = group > n_images ? n_images : group;
    const uint32_t n_groups = (n_images + group - 1) / group;
    const size_t rgb_chunk = align_up(rgb_per_image * chunk, 256);
    J_TRY(scratch_reserve(ctx, 1, in_bytes + 16 * ((size_t) n_groups + 1) + 64, &d_ecs));
    J_TRY(scratch_reserve(ctx, 2, (n_off + n_groups) * 8, &d_off));
    J_TRY(scratch_reserve(ctx, 3, 2 * sizeof(int32_t) * n_images + 64, &d_status));
    J_TRY(scratch_reserve(ctx, 6, 2 * rgb_chunk + 64, &d_rgb));
    if (raw_offsets) J_TRY(scratch_reserve(ctx, 5, in_bytes + 64, &d_raw));
    int32_t *d_lex_status = reinterpret_cast<int32_t *>(d_status) + n_images;
    if (!raw_offsets) CU_TRY(ctx, cudaMemsetAsync(d_lex_status, 0, sizeof(int32_t) * n_images, ctx->stream));

    uint32_t k = 0;  // running chunk index (the two RGB buffers alternate over the whole call)
    for (uint32_t g = 0; g < n_groups; ++g) {
        const uint32_t g0 = g * group, gn = (g0 + group <= n_images) ? group : n_images - g0;
        // ---- this group's bytes: host -> device, at the same offsets they have on the host
        uint64_t b0, b1;
        if (raw_offsets) {
            b0 = UINT64_MAX, b1 = 0;
            for (uint32_t i = g0; i < g0 + gn; ++i) {
                b0 = raw_offsets[i] < b0 ? raw_offsets[i] : b0;
                b1 = raw_offsets[i] + raw_lengths[i] > b1 ? raw_offsets[i] + raw_lengths[i] : b1;
            }
            if (b1 < b0) b0 = b1 = 0;
        } else {
            b0 = offsets[(uint64_t) g0 * n_ecs], b1 = offsets[(uint64_t) (g0 + gn) * n_ecs];
        }
        uint8_t  *g_ecs;
        uint64_t *g_off = reinterpret_cast<uint64_t *>(d_off) + (uint64_t) g0 * n_ecs + g;  // gn * n_ecs + 1 entries
        if (raw_offsets) {
            if (b1 > b0) CU_TRY(ctx, copy_h2d(ctx, reinterpret_cast<uint8_t *>(d_raw) + b0, bytes + b0, b1 - b0, ctx->stream));
            // the lexer writes the group's unstuffed bytes (never more than the raw ones) to a 16-byte aligned region of its own
            g_ecs = reinterpret_cast<uint8_t *>(d_ecs) + align_up(b0, 16) + 16 * (size_t) g;
            LexPlan plan;
            J_TRY(jpeg_lex_count(ctx, reinterpret_cast<uint8_t *>(d_raw), raw_offsets + g0, raw_lengths + g0, gn, &plan));
            J_TRY(jpeg_lex_scatter(ctx, reinterpret_cast<uint8_t *>(d_raw), &plan, n_ecs, g_ecs, g_off, d_lex_status + g0));
        } else {
            g_ecs = reinterpret_cast<uint8_t *>(d_ecs);  // offsets stay absolute
            if (b1 > b0) CU_TRY(ctx, copy_h2d(ctx, g_ecs + b0, bytes + b0, b1 - b0, ctx->stream));
            CU_TRY(ctx, copy_h2d(ctx, g_off, offsets + (uint64_t) g0 * n_ecs, ((uint64_t) gn * n_ecs + 1) * 8, ctx->stream));
        }
        // ---- entropy decode of the group.  Spectral planes start zeroed (decode.swift:2241-2256): SCAN_FRESH
        jpeg_sm100_dev_spectral spg = sp;
        spg.n_images = gn;
        for (uint32_t p = 0; p < n_planes; ++p) spg.plane[p].coef = sp.plane[p].coef + (size_t) g0 * sp.plane[p].image_stride;
        ctx->hint_interval_bytes = (b1 - b0) / ((uint64_t) gn * n_ecs);
        const int k3 = jpeg_huffman_decode_scan(ctx, scan, g_ecs, g_off, n_ecs, interval, JPEG_SM100_SCAN_FRESH,
                                                tables + (tables_shared ? 0 : (size_t) 8 * g0), tables_shared, &spg,
                                                reinterpret_cast<int32_t *>(d_status) + g0);
        ctx->hint_interval_bytes = 0;
        J_TRY(k3);
        // ---- transform + colour, chunk by chunk; each chunk's RGB goes home on the copy stream while the next is computed
        for (uint32_t c0 = 0; c0 < gn; c0 += chunk, ++k) {
            const uint32_t i0 = g0 + c0, cnt = (c0 + chunk <= gn) ? chunk : gn - c0;
            uint8_t       *rgb_buf = reinterpret_cast<uint8_t *>(d_rgb) + (k & 1) * rgb_chunk;
            jpeg_sm100_dev_spectral spk = sp;
            spk.n_images = cnt;
            for (uint32_t p = 0; p < n_planes; ++p) spk.plane[p].coef = sp.plane[p].coef + (size_t) i0 * sp.plane[p].image_stride;
            if (k >= 2) CU_TRY(ctx, cudaStreamWaitEvent(ctx->stream, ev_out[k - 2], 0));  // RGB buffer (k & 1) is free again
            J_TRY(spectral_to_rgb8_dev(ctx, &spk, quanta, sx, sy, cosited, rgb_buf));  // fused K1+K2, or K1 -> scratch planes -> K2
            CU_TRY(ctx, cudaEventRecord(ev_k[k], ctx->stream));
            CU_TRY(ctx, cudaStreamWaitEvent(ctx->copy_out, ev_k[k], 0));
            CU_TRY(ctx, copy_d2h(ctx, rgb + (size_t) i0 * rgb_per_image, rgb_buf, rgb_per_image * cnt, ctx->copy_out));
            CU_TRY(ctx, cudaEventRecord(ev_out[k], ctx->copy_out));
        }
    }
    std::vector<int32_t> st(2 * (size_t) n_images, 0);
    CU_TRY(ctx, copy_d2h(ctx, st.data(), d_status, 2 * sizeof(int32_t) * n_images, ctx->stream));
    // the batch call waits tens of milliseconds for PCIe.  Default: poll the two streams and yield the core between polls (no
    // wake-up latency when cores are free; a blocking-sync event costs ~5 % of the call at N = 1).  JPEG_SM100_WAIT=block sleeps on
    // blocking-sync events instead -- for hosts where many of these calls (one per context / per GPU process) share few cores.
    {
        const char *wm = getenv("JPEG_SM100_WAIT");
        const bool  block = wm && strcmp(wm, "block") == 0;
        for (cudaStream_t st_ : {ctx->stream, ctx->copy_out}) {
            if (block) {
                if (!ctx->wait_event) CU_TRY(ctx, cudaEventCreateWithFlags(&ctx->wait_event, cudaEventBlockingSync | cudaEventDisableTiming));
                CU_TRY(ctx, cudaEventRecord(ctx->wait_event, st_));
                CU_TRY(ctx, cudaEventSynchronize(ctx->wait_event));
            } else {
                cudaError_t q;
                while ((q = cudaStreamQuery(st_)) == cudaErrorNotReady) sched_yield();
                CU_TRY(ctx, q);
            }
        }
    }
    drain.armed = false;  // both streams are idle
    int first = 0;
    for (uint32_t i = 0; i < n_images; ++i) {
        // a lexer error is raised before the scan is pushed (decode.swift:3929), so it wins over a decode error; a table error
        // is raised when the scan is pushed, before its first bit is read
        const int32_t v = st[n_images + i] ? st[n_images + i] : (table_err[i] ? table_err[i] : st[i]);
        if (status) status[i] = v;
        if (v && !first) first = v;
    }
    return first;
}

}  // namespace

JPEG_API int jpeg_sm100_decode_batch_rgb8(jpeg_sm100_ctx *ctx, const jpeg_sm100_scan_desc *scan, uint32_t n_images,
                                          const uint8_t *ecs_concat, const uint64_t *ecs_offsets, uint32_t n_ecs,
                                          uint64_t interval, const jpeg_sm100_huff_table *tables, int tables_shared,
                                          const uint16_t *quanta, uint32_t sx, uint32_t sy, int cosited, uint8_t *rgb,
                                          int32_t *status)
{
    REQUIRE_CTX(ctx);
    if (!scan || !ecs_offsets || !tables || !quanta || !rgb || n_images == 0 || n_ecs == 0) return JPEG_SM100_ERR_INVALID_ARGUMENT;
    return decode_batch_common(ctx, scan, n_images, ecs_concat, ecs_offsets, nullptr, nullptr, n_ecs, interval, tables, tables_shared,
                               quanta, sx, sy, cosited, rgb, status);
}

JPEG_API int jpeg_sm100_decode_batch_raw_rgb8(jpeg_sm100_ctx *ctx, const jpeg_sm100_scan_desc *scan, uint32_t n_images,
                                              const uint8_t *raw_concat, const uint64_t *raw_offsets, const uint64_t *raw_lengths,
                                              uint32_t n_ecs, uint64_t interval, const jpeg_sm100_huff_table *tables,
                                              int tables_shared, const uint16_t *quanta, uint32_t sx, uint32_t sy, int cosited,
                                              uint8_t *rgb, int32_t *status)
{
    REQUIRE_CTX(ctx);
    if (!scan || !raw_concat || !raw_offsets || !raw_lengths || !tables || !quanta || !rgb || n_images == 0 || n_ecs == 0)
        return JPEG_SM100_ERR_INVALID_ARGUMENT;
    return decode_batch_common(ctx, scan, n_images, raw_concat, nullptr, raw_offsets, raw_lengths, n_ecs, interval, tables,
                               tables_shared, quanta, sx, sy, cosited, rgb, status);
}

// ---- N1: lexer entry points ----
JPEG_API int jpeg_sm100_dev_lex_scan(jpeg_sm100_ctx *ctx, const uint8_t *d_raw, const uint64_t *raw_offsets,
                                     const uint64_t *raw_lengths, uint32_t n_images, uint32_t n_ecs, uint8_t *d_ecs,
                                     uint64_t *d_ecs_offsets, int32_t *d_status)
{
    REQUIRE_CTX(ctx);
    if (!d_raw || !raw_offsets || !raw_lengths || !d_ecs || !d_ecs_offsets || n_images == 0 || n_ecs == 0) return JPEG_SM100_ERR_INVALID_ARGUMENT;
    if ((reinterpret_cast<uintptr_t>(d_raw) & 15) || (reinterpret_cast<uintptr_t>(d_ecs) & 15)) return JPEG_SM100_ERR_INVALID_ARGUMENT;
    LexPlan plan;
    J_TRY(jpeg_lex_count(ctx, d_raw, raw_offsets, raw_lengths, n_images, &plan));
    return jpeg_lex_scatter(ctx, d_raw, &plan, n_ecs, d_ecs, d_ecs_offsets, d_status);
}

namespace {
// uploads one image's raw scan bytes, lexes them on the device; leaves d_ecs (slot 1) / d_off (slot 2) ready
int lex_one(jpeg_sm100_ctx *ctx, const uint8_t *raw, uint64_t raw_len, uint32_t *n_ecs, void **d_ecs, void **d_off, int32_t *lex_status)
{
    void *d_raw = nullptr, *d_st = nullptr;
    J_TRY(scratch_reserve(ctx, 5, raw_len + 64, &d_raw));
    J_TRY(scratch_reserve(ctx, 1, raw_len + 64, d_ecs));
    J_TRY(scratch_reserve(ctx, 3, 64, &d_st));
    if (raw_len) CU_TRY(ctx, copy_h2d(ctx, d_raw, raw, raw_len, ctx->stream));
    const uint64_t off0 = 0;
    LexPlan        plan;
    J_TRY(jpeg_lex_count(ctx, reinterpret_cast<uint8_t *>(d_raw), &off0, &raw_len, 1, &plan));
    // the number of segments is data: one 4-byte read-back before the scatter pass can lay out the offsets
    J_TRY(jpeg_lex_scatter(ctx, reinterpret_cast<uint8_t *>(d_raw), &plan, 0, nullptr, nullptr, nullptr));
    uint32_t splits = 0;
    CU_TRY(ctx, copy_d2h(ctx, &splits, plan.d_n_splits, 4, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    *n_ecs = splits + 1;
    J_TRY(scratch_reserve(ctx, 2, 8 * ((size_t) *n_ecs + 1), d_off));
    J_TRY(jpeg_lex_scatter(ctx, reinterpret_cast<uint8_t *>(d_raw), &plan, *n_ecs, reinterpret_cast<uint8_t *>(*d_ecs),
                           reinterpret_cast<uint64_t *>(*d_off), reinterpret_cast<int32_t *>(d_st)));
    CU_TRY(ctx, copy_d2h(ctx, lex_status, d_st, 4, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return JPEG_SM100_OK;
}
}  // namespace

JPEG_API int jpeg_sm100_lex_scan(jpeg_sm100_ctx *ctx, const uint8_t *raw, uint64_t raw_len, uint8_t *ecs, uint64_t ecs_capacity,
                                 uint64_t *ecs_offsets, uint32_t offsets_capacity, uint32_t *n_ecs)
{
    REQUIRE_CTX(ctx);
    if ((!raw && raw_len) || !ecs_offsets || !n_ecs) return JPEG_SM100_ERR_INVALID_ARGUMENT;
    void   *d_ecs = nullptr, *d_off = nullptr;
    int32_t st = 0;
    J_TRY(lex_one(ctx, raw, raw_len, n_ecs, &d_ecs, &d_off, &st));
    if (st) return st;
    if (offsets_capacity < *n_ecs + 1) return JPEG_SM100_ERR_NO_MEMORY;
    CU_TRY(ctx, copy_d2h(ctx, ecs_offsets, d_off, 8 * ((size_t) *n_ecs + 1), ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    const uint64_t total = ecs_offsets[*n_ecs];
    if (total > ecs_capacity || (total && !ecs)) return JPEG_SM100_ERR_NO_MEMORY;
    if (total) CU_TRY(ctx, copy_d2h(ctx, ecs, d_ecs, total, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return JPEG_SM100_OK;
}

JPEG_API int jpeg_sm100_decode_scan_raw(jpeg_sm100_ctx *ctx, const jpeg_sm100_scan_desc *scan, const uint8_t *raw, uint64_t raw_len,
                                        uint64_t interval, int extend, const jpeg_sm100_huff_table dc[4],
                                        const jpeg_sm100_huff_table ac[4], jpeg_sm100_plane_i16 *planes, uint32_t n_planes)
{
    REQUIRE_CTX(ctx);
    if (!scan || !planes || !dc || !ac || (!raw && raw_len)) return JPEG_SM100_ERR_INVALID_ARGUMENT;
    void    *d_ecs = nullptr, *d_off = nullptr, *d_status = nullptr;
    uint32_t n_ecs = 0;
    int32_t  st = 0;
    J_TRY(lex_one(ctx, raw, raw_len, &n_ecs, &d_ecs, &d_off, &st));
    if (st) return st;
    if (interval == JPEG_SM100_INTERVAL_NONE) {
        if (n_ecs > 1) return JPEG_SM100_ERR_MISSING_INTERVAL;  // decode.swift:3708-3720
        n_ecs = 1;
    }
    PlaneSet ps;
    J_TRY(upload_spectral(ctx, 0, planes, n_planes, nullptr, ps));
    J_TRY(scratch_reserve(ctx, 3, 64, &d_status));
    jpeg_sm100_huff_table tables[8];
    memcpy(tables, dc, sizeof(jpeg_sm100_huff_table) * 4);
    memcpy(tables + 4, ac, sizeof(jpeg_sm100_huff_table) * 4);
    ctx->hint_interval_bytes = raw_len / n_ecs;
    const int k3 = jpeg_huffman_decode_scan(ctx, scan, reinterpret_cast<uint8_t *>(d_ecs), reinterpret_cast<uint64_t *>(d_off), n_ecs,
                                            interval, extend & (JPEG_SM100_SCAN_EXTEND | JPEG_SM100_SCAN_T81), tables, 1, &ps.sp, reinterpret_cast<int32_t *>(d_status));
    ctx->hint_interval_bytes = 0;
    J_TRY(k3);
    int32_t status = 0;
    CU_TRY(ctx, copy_d2h(ctx, &status, d_status, sizeof status, ctx->stream));
    for (uint32_t p = 0; p < n_planes; ++p)
        if (ps.bytes[p])
            CU_TRY(ctx, copy_d2h(ctx, planes[p].coef, ps.sp.plane[p].coef, ps.bytes[p], ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return status;
}

// ---- encode ----
JPEG_API int jpeg_sm100_pack_rgb8(jpeg_sm100_ctx *ctx, const uint8_t *rgb, uint64_t n_px, int arity, uint16_t *il)
{
    REQUIRE_CTX(ctx);
    if (arity != 1 && arity != 3) return JPEG_SM100_ERR_INVALID_ARGUMENT;
    void *d_rgb = nullptr, *d_il = nullptr;
    J_TRY(scratch_reserve(ctx, 6, n_px * 3 + 64, &d_rgb));
    J_TRY(scratch_reserve(ctx, 5, n_px * arity * 2 + 64, &d_il));
    if (n_px) CU_TRY(ctx, copy_h2d(ctx, d_rgb, rgb, n_px * 3, ctx->stream));
    J_TRY(jpeg_color_pack_rgb(ctx, reinterpret_cast<uint8_t *>(d_rgb), n_px, arity, reinterpret_cast<uint16_t *>(d_il)));
    if (n_px) CU_TRY(ctx, copy_d2h(ctx, il, d_il, n_px * arity * 2, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return JPEG_SM100_OK;
}

JPEG_API int jpeg_sm100_decompose(jpeg_sm100_ctx *ctx, const uint16_t *il, uint32_t sx, uint32_t sy,
                                  jpeg_sm100_plane_u16 *planes, uint32_t n)
{
    REQUIRE_CTX(ctx);
    jpeg_sm100_dev_planar pl;
    J_TRY(upload_planar_u16(ctx, 4, planes, n, false, pl));
    const size_t bytes = (size_t) sx * sy * n * 2;
    void        *d_il = nullptr;
    J_TRY(scratch_reserve(ctx, 5, bytes + 64, &d_il));
    if (bytes) CU_TRY(ctx, copy_h2d(ctx, d_il, il, bytes, ctx->stream));
    J_TRY(jpeg_color_decompose(ctx, d_il, false, sx, sy, &pl));
    for (uint32_t p = 0; p < n; ++p) {
        const size_t pb = (size_t) 128 * planes[p].units_x * planes[p].units_y;
        if (pb) CU_TRY(ctx, copy_d2h(ctx, planes[p].samples, pl.plane[p].samples, pb, ctx->stream));
    }
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return JPEG_SM100_OK;
}

JPEG_API int jpeg_sm100_fdct(jpeg_sm100_ctx *ctx, const uint16_t *samples, uint32_t ux, uint32_t uy,
                             const uint16_t quanta[64], int precision, int16_t *coef)
{
    REQUIRE_CTX(ctx);
    if (!quanta || ((uint64_t) ux * uy && (!samples || !coef))) return JPEG_SM100_ERR_INVALID_ARGUMENT;
    jpeg_sm100_plane_u16  hp = {const_cast<uint16_t *>(samples), (int32_t) ux, (int32_t) uy, 1, 1};
    jpeg_sm100_dev_planar pl;
    J_TRY(upload_planar_u16(ctx, 4, &hp, 1, true, pl));
    jpeg_sm100_plane_i16 geo = {nullptr, (int32_t) ux, (int32_t) uy};
    PlaneSet             ps;
    // reuse upload_spectral for the allocation only (no source pointer => nothing copied when empty)
    memset(&ps, 0, sizeof ps);
    {
        void *base = nullptr;
        ps.bytes[0] = (size_t) 128 * ux * uy;
        J_TRY(scratch_reserve(ctx, 0, ps.bytes[0] + 1024, &base));
        ps.sp.n_images = 1;
        ps.sp.n_planes = 1;
        ps.sp.plane[0].coef = reinterpret_cast<int16_t *>(base);
        ps.sp.plane[0].image_stride = ps.bytes[0] / 2;
        ps.sp.plane[0].units_x = geo.units_x;
        ps.sp.plane[0].units_y = geo.units_y;
        ps.sp.plane[0].factor_x = ps.sp.plane[0].factor_y = 1;
    }
    J_TRY(jpeg_sm100_dev_fdct(ctx, &pl, quanta, precision, &ps.sp));
    if (ps.bytes[0]) CU_TRY(ctx, copy_d2h(ctx, coef, ps.sp.plane[0].coef, ps.bytes[0], ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return JPEG_SM100_OK;
}

JPEG_API int jpeg_sm100_rgb8_to_spectral(jpeg_sm100_ctx *ctx, const uint8_t *rgb, uint32_t sx, uint32_t sy,
                                         jpeg_sm100_plane_i16 *planes, uint32_t n, const uint16_t *quanta,
                                         const int32_t *factors)
{
    REQUIRE_CTX(ctx);
    if (!planes || !quanta || !factors || (n != 1 && n != 3)) return JPEG_SM100_ERR_INVALID_ARGUMENT;
    // device coefficient planes (no upload), 8-bit sample planes, RGB input
    PlaneSet ps;
    memset(&ps, 0, sizeof ps);
    size_t total = 0, off[4];
    for (uint32_t p = 0; p < n; ++p) {
        ps.bytes[p] = (size_t) 128 * planes[p].units_x * planes[p].units_y;
        off[p] = total;
        total += align_up(ps.bytes[p], 1024);
    }
    void *base = nullptr;
    J_TRY(scratch_reserve(ctx, 0, total + 1024, &base));
    ps.sp.n_images = 1;
    ps.sp.n_planes = n;
    for (uint32_t p = 0; p < n; ++p) {
        ps.sp.plane[p].coef = reinterpret_cast<int16_t *>(reinterpret_cast<uint8_t *>(base) + off[p]);
        ps.sp.plane[p].image_stride = ps.bytes[p] / 2;
        ps.sp.plane[p].units_x = planes[p].units_x;
        ps.sp.plane[p].units_y = planes[p].units_y;
        ps.sp.plane[p].factor_x = factors[2 * p];
        ps.sp.plane[p].factor_y = factors[2 * p + 1];
    }
    jpeg_sm100_dev_planar pl;
    J_TRY(alloc_planar(ctx, 4, ps.sp, 1, pl));
    const size_t rgb_bytes = (size_t) sx * sy * 3;
    void        *d_rgb = nullptr;
    J_TRY(scratch_reserve(ctx, 6, rgb_bytes + 64, &d_rgb));
    if (rgb_bytes) CU_TRY(ctx, copy_h2d(ctx, d_rgb, rgb, rgb_bytes, ctx->stream));
    J_TRY(jpeg_color_decompose(ctx, d_rgb, true, sx, sy, &pl));
    J_TRY(jpeg_sm100_dev_fdct(ctx, &pl, quanta, 8, &ps.sp));
    for (uint32_t p = 0; p < n; ++p)
        if (ps.bytes[p]) CU_TRY(ctx, copy_d2h(ctx, planes[p].coef, ps.sp.plane[p].coef, ps.bytes[p], ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return JPEG_SM100_OK;
}

JPEG_API int jpeg_sm100_encode_scan(jpeg_sm100_ctx *ctx, const jpeg_sm100_scan_desc *scan, const jpeg_sm100_plane_i16 *planes,
                                    uint32_t n_planes, uint64_t interval_mcus, jpeg_sm100_huff_table dc_out[4],
                                    jpeg_sm100_huff_table ac_out[4], uint8_t *ecs, uint64_t ecs_capacity, uint64_t *ecs_len)
{
    REQUIRE_CTX(ctx);
    if (!scan || !planes || !dc_out || !ac_out || !ecs_len) return JPEG_SM100_ERR_INVALID_ARGUMENT;
    PlaneSet ps;
    J_TRY(upload_spectral(ctx, 0, planes, n_planes, nullptr, ps));
    // worst case: every coefficient costs 16 + 16 bits, doubled by stuffing, plus RST markers
    uint64_t blocks = 0;
    for (uint32_t p = 0; p < n_planes; ++p) blocks += (uint64_t) (planes[p].units_x + 4) * (planes[p].units_y + 4);
    const uint64_t cap = blocks * 64 * 8 + 4096;
    void          *d_ecs = nullptr, *d_len = nullptr;
    J_TRY(scratch_reserve(ctx, 1, cap, &d_ecs));
    J_TRY(scratch_reserve(ctx, 3, 64, &d_len));
    jpeg_sm100_huff_table tables[8];
    uint64_t              needed = 0;
    J_TRY(jpeg_huffman_encode_scan(ctx, scan, &ps.sp, interval_mcus, tables, reinterpret_cast<uint8_t *>(d_ecs), cap,
                                   reinterpret_cast<uint64_t *>(d_len), &needed));
    memcpy(dc_out, tables, sizeof(jpeg_sm100_huff_table) * 4);
    memcpy(ac_out, tables + 4, sizeof(jpeg_sm100_huff_table) * 4);
    *ecs_len = needed;
    if (needed > ecs_capacity) return JPEG_SM100_ERR_NO_MEMORY;
    if (needed) CU_TRY(ctx, copy_d2h(ctx, ecs, d_ecs, needed, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return JPEG_SM100_OK;
}

// =====================================================================================================================
// layer A with the image resident on the device (jpeg_sm100_spectral): JPEG.Context's one Spectral per file, kept in HBM
// =====================================================================================================================
struct jpeg_sm100_spectral {
    uint32_t n_planes = 0;
    int32_t  ux[4] = {0, 0, 0, 0}, uy[4] = {0, 0, 0, 0};
    int16_t *coef[4] = {nullptr, nullptr, nullptr, nullptr};  // one allocation, plane p at coef[p]
    void    *base = nullptr;
    size_t   bytes = 0;
};

namespace {

int spectral_layout(uint32_t n, const int32_t *units_xy, size_t off[4], size_t *total)
{
    if (n < 1 || n > 4 || !units_xy) return JPEG_SM100_ERR_INVALID_ARGUMENT;
    size_t t = 0;
    for (uint32_t p = 0; p < n; ++p) {
        if (units_xy[2 * p] < 0 || units_xy[2 * p + 1] < 0) return JPEG_SM100_ERR_INVALID_ARGUMENT;
        off[p] = t;
        t += align_up((size_t) 128 * units_xy[2 * p] * units_xy[2 * p + 1], 1024);
    }
    *total = t + 1024;
    return JPEG_SM100_OK;
}

void spectral_view(const jpeg_sm100_spectral *s, const int32_t *factors, jpeg_sm100_dev_spectral &sp)
{
    memset(&sp, 0, sizeof sp);
    sp.n_images = 1;
    sp.n_planes = s->n_planes;
    for (uint32_t p = 0; p < s->n_planes; ++p) {
        sp.plane[p].coef = s->coef[p];
        sp.plane[p].image_stride = (uint64_t) 64 * s->ux[p] * s->uy[p];
        sp.plane[p].units_x = s->ux[p];
        sp.plane[p].units_y = s->uy[p];
        sp.plane[p].factor_x = factors ? factors[2 * p] : 1;
        sp.plane[p].factor_y = factors ? factors[2 * p + 1] : 1;
    }
}

}  // namespace

JPEG_API int jpeg_sm100_spectral_create(jpeg_sm100_ctx *ctx, uint32_t n_planes, const int32_t *units_xy, jpeg_sm100_spectral **out)
{
    REQUIRE_CTX(ctx);
    if (!out) return JPEG_SM100_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    size_t off[4], total = 0;
    J_TRY(spectral_layout(n_planes, units_xy, off, &total));
    jpeg_sm100_spectral *s = new jpeg_sm100_spectral();
    if (cudaMalloc(&s->base, total) != cudaSuccess) {
        delete s;
        cudaGetLastError();
        return JPEG_SM100_ERR_NO_MEMORY;
    }
    s->bytes = total;
    s->n_planes = n_planes;
    for (uint32_t p = 0; p < n_planes; ++p) {
        s->ux[p] = units_xy[2 * p], s->uy[p] = units_xy[2 * p + 1];
        s->coef[p] = reinterpret_cast<int16_t *>(reinterpret_cast<uint8_t *>(s->base) + off[p]);
    }
    const cudaError_t e = cudaMemsetAsync(s->base, 0, total, ctx->stream);  // new Spectral planes are zero (decode.swift:2241-2256)
    if (e != cudaSuccess) {
        cudaFree(s->base);
        delete s;
        return jpeg_cuda_fail(ctx, e, "cudaMemsetAsync", __FILE__, __LINE__);
    }
    *out = s;
    return JPEG_SM100_OK;
}

JPEG_API void jpeg_sm100_spectral_destroy(jpeg_sm100_ctx *ctx, jpeg_sm100_spectral *s)
{
    if (!s) return;
    if (ctx) {
        cudaSetDevice(ctx->device);
        cudaStreamSynchronize(ctx->stream);
    }
    if (s->base) cudaFree(s->base);
    delete s;
}

JPEG_API int jpeg_sm100_spectral_resize(jpeg_sm100_ctx *ctx, jpeg_sm100_spectral *s, const int32_t *units_xy)
{
    REQUIRE_CTX(ctx);
    if (!s) return JPEG_SM100_ERR_INVALID_ARGUMENT;
    size_t off[4], total = 0;
    J_TRY(spectral_layout(s->n_planes, units_xy, off, &total));
    void *nb = nullptr;
    if (cudaMalloc(&nb, total) != cudaSuccess) {
        cudaGetLastError();
        return JPEG_SM100_ERR_NO_MEMORY;
    }
    CU_TRY(ctx, cudaMemsetAsync(nb, 0, total, ctx->stream));
    for (uint32_t p = 0; p < s->n_planes; ++p) {
        const int32_t nx = units_xy[2 * p], ny = units_xy[2 * p + 1];
        const int32_t cx = nx < s->ux[p] ? nx : s->ux[p], cy = ny < s->uy[p] ? ny : s->uy[p];
        if (cx > 0 && cy > 0)  // rows of 128-byte blocks: the common region keeps its coefficients (decode.swift:2456-2495)
            CU_TRY(ctx, cudaMemcpy2DAsync(reinterpret_cast<uint8_t *>(nb) + off[p], (size_t) 128 * nx, s->coef[p], (size_t) 128 * s->ux[p],
                                          (size_t) 128 * cx, cy, cudaMemcpyDeviceToDevice, ctx->stream));
    }
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFree(s->base);
    s->base = nb;
    s->bytes = total;
    for (uint32_t p = 0; p < s->n_planes; ++p) {
        s->ux[p] = units_xy[2 * p], s->uy[p] = units_xy[2 * p + 1];
        s->coef[p] = reinterpret_cast<int16_t *>(reinterpret_cast<uint8_t *>(nb) + off[p]);
    }
    return JPEG_SM100_OK;
}

static int spectral_check_planes(const jpeg_sm100_spectral *s, const jpeg_sm100_plane_i16 *planes, uint32_t n)
{
    if (!s || !planes || n != s->n_planes) return JPEG_SM100_ERR_INVALID_ARGUMENT;
    for (uint32_t p = 0; p < n; ++p)
        if (planes[p].units_x != s->ux[p] || planes[p].units_y != s->uy[p] || ((size_t) s->ux[p] * s->uy[p] && !planes[p].coef))
            return JPEG_SM100_ERR_INVALID_ARGUMENT;
    return JPEG_SM100_OK;
}

JPEG_API int jpeg_sm100_spectral_upload(jpeg_sm100_ctx *ctx, jpeg_sm100_spectral *s, const jpeg_sm100_plane_i16 *planes, uint32_t n)
{
    REQUIRE_CTX(ctx);
    J_TRY(spectral_check_planes(s, planes, n));
    for (uint32_t p = 0; p < n; ++p) {
        const size_t b = (size_t) 128 * s->ux[p] * s->uy[p];
        if (b) CU_TRY(ctx, copy_h2d(ctx, s->coef[p], planes[p].coef, b, ctx->stream));
    }
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return JPEG_SM100_OK;
}

JPEG_API int jpeg_sm100_spectral_download(jpeg_sm100_ctx *ctx, const jpeg_sm100_spectral *s, jpeg_sm100_plane_i16 *planes, uint32_t n)
{
    REQUIRE_CTX(ctx);
    J_TRY(spectral_check_planes(s, planes, n));
    for (uint32_t p = 0; p < n; ++p) {
        const size_t b = (size_t) 128 * s->ux[p] * s->uy[p];
        if (b) CU_TRY(ctx, copy_d2h(ctx, planes[p].coef, s->coef[p], b, ctx->stream));
    }
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return JPEG_SM100_OK;
}

static int spectral_decode_common(jpeg_sm100_ctx *ctx, jpeg_sm100_spectral *s, const jpeg_sm100_scan_desc *scan, void *d_ecs, void *d_off,
                                  uint32_t n_ecs, uint64_t bytes, uint64_t interval, int extend, const jpeg_sm100_huff_table dc[4],
                                  const jpeg_sm100_huff_table ac[4])
{
    void *d_status = nullptr;
    J_TRY(scratch_reserve(ctx, 3, 64, &d_status));
    jpeg_sm100_huff_table tables[8];
    memcpy(tables, dc, sizeof(jpeg_sm100_huff_table) * 4);
    memcpy(tables + 4, ac, sizeof(jpeg_sm100_huff_table) * 4);
    jpeg_sm100_dev_spectral sp;
    spectral_view(s, nullptr, sp);
    ctx->hint_interval_bytes = n_ecs ? bytes / n_ecs : 0;
    const int k3 = jpeg_huffman_decode_scan(ctx, scan, reinterpret_cast<uint8_t *>(d_ecs), reinterpret_cast<uint64_t *>(d_off), n_ecs, interval,
                                            extend & (JPEG_SM100_SCAN_EXTEND | JPEG_SM100_SCAN_T81), tables, 1, &sp, reinterpret_cast<int32_t *>(d_status));
    ctx->hint_interval_bytes = 0;
    J_TRY(k3);
    int32_t status = 0;  // the reference throws at the first failing interval: the host needs the verdict before the next scan
    CU_TRY(ctx, copy_d2h(ctx, &status, d_status, sizeof status, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return status;
}

JPEG_API int jpeg_sm100_spectral_decode_scan(jpeg_sm100_ctx *ctx, jpeg_sm100_spectral *s, const jpeg_sm100_scan_desc *scan,
                                             const uint8_t *ecs_concat, const uint64_t *ecs_offsets, uint32_t n_ecs, uint64_t interval,
                                             int extend, const jpeg_sm100_huff_table dc[4], const jpeg_sm100_huff_table ac[4])
{
    REQUIRE_CTX(ctx);
    if (!s || !scan || !dc || !ac || (n_ecs && !ecs_offsets)) return JPEG_SM100_ERR_INVALID_ARGUMENT;
    if (n_ecs == 0) return JPEG_SM100_OK;
    if (interval == JPEG_SM100_INTERVAL_NONE) n_ecs = 1;  // decode.swift:3500: the stride yields one element
    const uint64_t total = ecs_offsets[n_ecs];
    void          *d_ecs = nullptr, *d_off = nullptr;
    J_TRY(scratch_reserve(ctx, 1, total + 64, &d_ecs));
    J_TRY(scratch_reserve(ctx, 2, sizeof(uint64_t) * ((size_t) n_ecs + 1), &d_off));
    if (total) CU_TRY(ctx, copy_h2d(ctx, d_ecs, ecs_concat, total, ctx->stream));
    CU_TRY(ctx, copy_h2d(ctx, d_off, ecs_offsets, sizeof(uint64_t) * ((size_t) n_ecs + 1), ctx->stream));
    return spectral_decode_common(ctx, s, scan, d_ecs, d_off, n_ecs, total, interval, extend, dc, ac);
}

JPEG_API int jpeg_sm100_spectral_decode_scan_raw(jpeg_sm100_ctx *ctx, jpeg_sm100_spectral *s, const jpeg_sm100_scan_desc *scan,
                                                 const uint8_t *raw, uint64_t raw_len, uint64_t interval, int extend,
                                                 const jpeg_sm100_huff_table dc[4], const jpeg_sm100_huff_table ac[4])
{
    REQUIRE_CTX(ctx);
    if (!s || !scan || !dc || !ac || (!raw && raw_len)) return JPEG_SM100_ERR_INVALID_ARGUMENT;
    void    *d_ecs = nullptr, *d_off = nullptr;
    uint32_t n_ecs = 0;
    int32_t  st = 0;
    J_TRY(lex_one(ctx, raw, raw_len, &n_ecs, &d_ecs, &d_off, &st));
    if (st) return st;
    if (interval == JPEG_SM100_INTERVAL_NONE) {
        if (n_ecs > 1) return JPEG_SM100_ERR_MISSING_INTERVAL;  // decode.swift:3708-3720
        n_ecs = 1;
    }
    return spectral_decode_common(ctx, s, scan, d_ecs, d_off, n_ecs, raw_len, interval, extend, dc, ac);
}

JPEG_API int jpeg_sm100_spectral_encode_scan(jpeg_sm100_ctx *ctx, const jpeg_sm100_spectral *s, const jpeg_sm100_scan_desc *scan,
                                             uint64_t interval_mcus, jpeg_sm100_huff_table dc_out[4], jpeg_sm100_huff_table ac_out[4],
                                             uint8_t *ecs, uint64_t ecs_capacity, uint64_t *ecs_len)
{
    REQUIRE_CTX(ctx);
    if (!s || !scan || !dc_out || !ac_out || !ecs_len) return JPEG_SM100_ERR_INVALID_ARGUMENT;
    jpeg_sm100_dev_spectral sp;
    spectral_view(s, nullptr, sp);
    uint64_t blocks = 0;
    for (uint32_t p = 0; p < s->n_planes; ++p) blocks += (uint64_t) (s->ux[p] + 4) * (s->uy[p] + 4);
    const uint64_t cap = blocks * 64 * 8 + 4096;
    void          *d_ecs = nullptr, *d_len = nullptr;
    J_TRY(scratch_reserve(ctx, 1, cap, &d_ecs));
    J_TRY(scratch_reserve(ctx, 3, 64, &d_len));
    jpeg_sm100_huff_table tables[8];
    uint64_t              needed = 0;
    J_TRY(jpeg_huffman_encode_scan(ctx, scan, &sp, interval_mcus, tables, reinterpret_cast<uint8_t *>(d_ecs), cap,
                                   reinterpret_cast<uint64_t *>(d_len), &needed));
    memcpy(dc_out, tables, sizeof(jpeg_sm100_huff_table) * 4);
    memcpy(ac_out, tables + 4, sizeof(jpeg_sm100_huff_table) * 4);
    *ecs_len = needed;
    if (needed > ecs_capacity) return JPEG_SM100_ERR_NO_MEMORY;
    if (needed) CU_TRY(ctx, copy_d2h(ctx, ecs, d_ecs, needed, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return JPEG_SM100_OK;
}

JPEG_API int jpeg_sm100_spectral_idct(jpeg_sm100_ctx *ctx, const jpeg_sm100_spectral *s, const uint16_t *quanta, int precision,
                                      jpeg_sm100_plane_u16 *planes, uint32_t n)
{
    REQUIRE_CTX(ctx);
    if (!s || !quanta || !planes || n != s->n_planes) return JPEG_SM100_ERR_INVALID_ARGUMENT;
    jpeg_sm100_dev_spectral sp;
    spectral_view(s, nullptr, sp);
    jpeg_sm100_dev_planar pl;
    J_TRY(alloc_planar(ctx, 4, sp, 2, pl));
    J_TRY(jpeg_sm100_dev_idct(ctx, &sp, quanta, precision, &pl));
    for (uint32_t p = 0; p < n; ++p) {
        if (planes[p].units_x != s->ux[p] || planes[p].units_y != s->uy[p]) return JPEG_SM100_ERR_INVALID_ARGUMENT;
        const size_t b = (size_t) 128 * s->ux[p] * s->uy[p];
        if (b && planes[p].samples) CU_TRY(ctx, copy_d2h(ctx, planes[p].samples, pl.plane[p].samples, b, ctx->stream));
    }
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return JPEG_SM100_OK;
}

JPEG_API int jpeg_sm100_spectral_rgb8(jpeg_sm100_ctx *ctx, const jpeg_sm100_spectral *s, const uint16_t *quanta, const int32_t *factors,
                                      uint32_t sx, uint32_t sy, int cosited, uint8_t *rgb)
{
    REQUIRE_CTX(ctx);
    if (!s || !quanta || !factors || (s->n_planes != 1 && s->n_planes != 3)) return JPEG_SM100_ERR_INVALID_ARGUMENT;
    jpeg_sm100_dev_spectral sp;
    spectral_view(s, factors, sp);
    const size_t bytes = (size_t) sx * sy * 3;
    void        *d_rgb = nullptr;
    J_TRY(scratch_reserve(ctx, 6, bytes + 64, &d_rgb));
    J_TRY(spectral_to_rgb8_dev(ctx, &sp, quanta, sx, sy, cosited, reinterpret_cast<uint8_t *>(d_rgb)));
    if (bytes) CU_TRY(ctx, copy_d2h(ctx, rgb, d_rgb, bytes, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return JPEG_SM100_OK;
}

JPEG_API void jpeg_sm100_transfer_counts(jpeg_sm100_ctx *ctx, uint64_t *h2d, uint64_t *d2h)
{
    if (h2d) *h2d = ctx ? ctx->h2d_bytes : 0;
    if (d2h) *d2h = ctx ? ctx->d2h_bytes : 0;
}

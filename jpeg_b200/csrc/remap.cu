// remap.cu -- restart intervals that are not whole MCU rows (ITU-T T.81 E.1.4: interval e starts at MCU e * Ri wherever that
// falls in its row).  The reference places intervals by integer division on ROWS (decode.swift:3205-3207, 2897-2899) and its
// encoder writes no DRI at all, so this is an extension (decode: JPEG_SM100_SCAN_T81; encode: any interval_mcus).
//
// Every entropy kernel of this library walks an interval as whole rows of the MCU grid.  Rather than teaching a dozen kernels a
// second placement, the scan is run on a VIRTUAL grid of the same MCUs in the same order: g = gcd(Ri, W * H) columns and
// W * H / g rows.  MCU m of the real grid (row m / W, column m % W) is MCU m of the virtual grid (row m / g, column m % g), every
// interval is Ri / g whole virtual rows, and the last, shorter interval is whole rows too (W * H mod Ri is a multiple of g).
// One kernel copies the blocks of the scan's components into virtual planes before the scan (progressive scans read what earlier
// scans left; blocks of partial MCUs outside their plane read as zero, decode.swift:1459-1464) and back afterwards (blocks
// outside the plane are dropped, decode.swift:1470-1475): 128 bytes per block each way, 8 lanes per block.
#include "common.cuh"

namespace {

struct RemapParams {
    int16_t *real[4], *virt[4];
    uint64_t real_stride[4], virt_stride[4];
    int32_t  ux[4], uy[4], fx[4], fy[4];
    int32_t  W, g, mcu_blocks;
    uint32_t M;
    uint8_t  blk_comp[12], blk_dx[12], blk_dy[12];
};

template <bool TO_VIRTUAL>
__global__ void __launch_bounds__(256) k_remap_intervals(const __grid_constant__ RemapParams R, const uint32_t n_images)
{
    const uint64_t per_image = (uint64_t) R.M * (uint32_t) R.mcu_blocks, total = per_image * n_images;
    const uint32_t part = threadIdx.x & 7u;
    for (uint64_t i = ((uint64_t) blockIdx.x * blockDim.x + threadIdx.x) >> 3; i < total; i += ((uint64_t) gridDim.x * blockDim.x) >> 3) {
        const uint32_t img = (uint32_t) (i / per_image);
        const uint64_t rem = i - (uint64_t) img * per_image;
        const uint32_t m = (uint32_t) (rem / (uint32_t) R.mcu_blocks), b = (uint32_t) (rem - (uint64_t) m * (uint32_t) R.mcu_blocks);
        const int      c = R.blk_comp[b];
        if (!R.real[c]) continue;
        const uint32_t ry = m / (uint32_t) R.W, rx = m - ry * (uint32_t) R.W, vy = m / (uint32_t) R.g, vx = m - vy * (uint32_t) R.g;
        const uint32_t x = rx * (uint32_t) R.fx[c] + R.blk_dx[b], y = ry * (uint32_t) R.fy[c] + R.blk_dy[b];
        const bool     in_plane = x < (uint32_t) R.ux[c] && y < (uint32_t) R.uy[c];
        uint4 *v = reinterpret_cast<uint4 *>(R.virt[c] + (size_t) img * R.virt_stride[c] +
                                             64 * ((size_t) R.g * R.fx[c] * (vy * (uint32_t) R.fy[c] + R.blk_dy[b]) + vx * (uint32_t) R.fx[c] + R.blk_dx[b])) + part;
        uint4 *r = reinterpret_cast<uint4 *>(R.real[c] + (size_t) img * R.real_stride[c] + 64 * ((size_t) R.ux[c] * y + x)) + part;
        if (TO_VIRTUAL) *v = in_plane ? *r : make_uint4(0, 0, 0, 0);
        else if (in_plane) *r = *v;
    }
}

uint64_t gcd64(uint64_t a, uint64_t b)
{
    while (b) {
        const uint64_t t = a % b;
        a = b, b = t;
    }
    return a;
}

}  // namespace

// true: the scan has to run on the virtual grid (a restart interval that is neither absent nor a whole number of MCU rows)
bool jpeg_virtual_scan_needed(const jpeg_sm100_scan_desc *scan, const jpeg_sm100_dev_spectral *sp, uint64_t interval)
{
    if (!scan || !sp || interval == 0 || interval == JPEG_SM100_INTERVAL_NONE || scan->n_comp < 1 || scan->n_comp > 4) return false;
    int64_t W;
    if (scan->n_comp > 1) W = scan->blocks_x;
    else {
        const int p = scan->comp[0].plane;
        if (p < 0 || p >= (int) sp->n_planes) return false;
        W = sp->plane[p].units_x;
    }
    return W > 0 && interval % (uint64_t) W != 0;
}

// scratch slot 16 belongs to this file.  Fills `v` (virtual planes, scan description on the virtual grid, copy parameters).
int jpeg_virtual_scan_setup(jpeg_sm100_ctx *ctx, const jpeg_sm100_scan_desc *scan, const jpeg_sm100_dev_spectral *sp, uint64_t interval,
                            JpegVirtualScan *v)
{
    static_assert(sizeof(RemapParams) <= sizeof(v->params), "JpegVirtualScan::params holds the copy parameters");
    const bool interleaved = scan->n_comp > 1;
    int64_t    W, H;
    if (interleaved) W = scan->blocks_x, H = scan->blocks_y;
    else {
        const int p = scan->comp[0].plane;
        if (p < 0 || p >= (int) sp->n_planes) return JPEG_SM100_ERR_INVALID_ARGUMENT;
        W = sp->plane[p].units_x, H = sp->plane[p].units_y;
    }
    if (W <= 0 || H <= 0 || (uint64_t) W * (uint64_t) H > 0x7fffffffull) return JPEG_SM100_ERR_UNSUPPORTED;
    const uint64_t M = (uint64_t) W * (uint64_t) H, g = gcd64(interval, M);
    if (M / g > 0xffffull) return JPEG_SM100_ERR_UNSUPPORTED;  // the decoders count MCU rows in 16 bits (more than 65535 virtual rows: a huge image with an interval coprime to its MCU count)
    RemapParams    R;
    memset(&R, 0, sizeof R);
    R.W = (int32_t) W, R.g = (int32_t) g, R.M = (uint32_t) M;
    v->scan = *scan;
    v->scan.blocks_x = (int32_t) g, v->scan.blocks_y = (int32_t) (M / g);
    memset(&v->sp, 0, sizeof v->sp);
    v->sp.n_images = sp->n_images;
    v->sp.n_planes = (uint32_t) scan->n_comp;
    size_t off[4], total = 0;
    int    volume = 0;
    for (int c = 0; c < scan->n_comp; ++c) {
        const int p = scan->comp[c].plane;
        if (p >= (int) sp->n_planes) return JPEG_SM100_ERR_INVALID_ARGUMENT;
        const int fx = interleaved ? scan->comp[c].factor_x : 1, fy = interleaved ? scan->comp[c].factor_y : 1;
        if (fx < 1 || fy < 1 || fx > 4 || fy > 4) return JPEG_SM100_ERR_INVALID_ARGUMENT;
        R.fx[c] = fx, R.fy[c] = fy;
        for (int dy = 0; dy < fy; ++dy)
            for (int dx = 0; dx < fx; ++dx) {
                if (volume >= 12) return JPEG_SM100_ERR_INVALID_ARGUMENT;
                R.blk_comp[volume] = (uint8_t) c, R.blk_dx[volume] = (uint8_t) dx, R.blk_dy[volume] = (uint8_t) dy;
                ++volume;
            }
        auto &vp = v->sp.plane[c];
        vp.units_x = (int32_t) (g * fx), vp.units_y = (int32_t) (M / g * fy);
        vp.factor_x = scan->comp[c].factor_x, vp.factor_y = scan->comp[c].factor_y;
        vp.image_stride = (uint64_t) 64 * vp.units_x * vp.units_y;
        v->scan.comp[c].plane = p < 0 ? -1 : c;
        off[c] = total;
        if (p >= 0) {
            if ((reinterpret_cast<uintptr_t>(sp->plane[p].coef) & 15) || (sp->plane[p].image_stride & 7)) return JPEG_SM100_ERR_UNSUPPORTED;
            R.real[c] = sp->plane[p].coef, R.real_stride[c] = sp->plane[p].image_stride;
            R.ux[c] = sp->plane[p].units_x, R.uy[c] = sp->plane[p].units_y;
            total += (size_t) vp.image_stride * 2 * sp->n_images + 1024;
            total = (total + 1023) / 1024 * 1024;
        }
    }
    R.mcu_blocks = volume;
    void *base = nullptr;
    J_TRY(scratch_reserve(ctx, 16, total + 1024, &base));
    for (int c = 0; c < scan->n_comp; ++c)
        if (R.real[c]) {
            R.virt[c] = reinterpret_cast<int16_t *>(reinterpret_cast<uint8_t *>(base) + off[c]);
            R.virt_stride[c] = v->sp.plane[c].image_stride;
            v->sp.plane[c].coef = R.virt[c];
        }
    memcpy(v->params, &R, sizeof R);
    return JPEG_SM100_OK;
}

int jpeg_virtual_scan_copy(jpeg_sm100_ctx *ctx, const JpegVirtualScan *v, bool to_virtual)
{
    RemapParams R;
    memcpy(&R, v->params, sizeof R);
    const uint64_t items = (uint64_t) R.M * (uint32_t) R.mcu_blocks * v->sp.n_images;
    if (items == 0) return JPEG_SM100_OK;
    const uint32_t grid = (uint32_t) std::min<uint64_t>((items * 8 + 255) / 256, (uint64_t) ctx->sm_count * 32);
    if (to_virtual) k_remap_intervals<true><<<grid, 256, 0, ctx->stream>>>(R, v->sp.n_images);
    else k_remap_intervals<false><<<grid, 256, 0, ctx->stream>>>(R, v->sp.n_images);
    LAUNCH_CHECK(ctx);
    return JPEG_SM100_OK;
}

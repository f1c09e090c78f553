// color.cu -- K2 (upsample + interleave + YCbCr->RGB + pack) and K4 (RGB->YCbCr + box downsample).
//
// Replaces Planar.interleaved(cosite:) (decode.swift:4182-4276), Color.unpack for RGB / YCbCr
// (jpeg.swift:441-453, 493-572), RGB.pack (jpeg.swift:463-478, 584-599) and Rectangular.decomposed()
// (encode.swift:389-425).  All arithmetic is the reference's binary32 sequence, never contracted.
#include "common.cuh"
#include "pixel_core.cuh"

namespace {

struct PlanarView {
    const void *samples[4];
    uint64_t    image_stride[4];  // samples
    int32_t     width[4];         // 8 * units_x
    int32_t     height[4];        // 8 * units_y
    int32_t     fx[4], fy[4];
    int32_t     n_planes;
    int32_t     scale_x, scale_y;
    int32_t     size_x, size_y;
    int32_t     cosited;
    uint32_t    n_images;
};

template <typename T>
__device__ __forceinline__ float sample_at(const T *pl, int w, int x, int y)
{
    return (float) pl[(size_t) w * y + x];
}

// decode.swift:4236-4265: one output sample of plane p at pixel (x, y), as the UInt16 the reference stores
template <typename T>
__device__ __forceinline__ uint32_t upsampled(const PlanarView &V, int p, uint32_t img, int x, int y)
{
    const T *pl = reinterpret_cast<const T *>(V.samples[p]) + (size_t) img * V.image_stride[p];
    const int w = V.width[p];
    if (V.n_planes == 1 || (V.fx[p] == V.scale_x && V.fy[p] == V.scale_y)) return (uint32_t) pl[(size_t) w * y + x];
    int ax, ay, bx, by, cx, cy;
    if (V.cosited) {
        ax = ay = 0;
        bx = V.fx[p], by = V.fy[p];
        cx = V.scale_x, cy = V.scale_y;
    } else {
        ax = V.fx[p] - V.scale_x, ay = V.fy[p] - V.scale_y;
        bx = 2 * V.fx[p], by = 2 * V.fy[p];
        cx = 2 * V.scale_x, cy = 2 * V.scale_y;
    }
    const int dx = w - 1, dy = V.height[p] - 1;
    const int nx = ax + bx * x, ny = ay + by * y;
    const int ix = nx / cx, rx = nx - ix * cx;  // truncating, like quotientAndRemainder
    const int iy = ny / cy, ry = ny - iy * cy;
    const int jx = min(ix + 1, dx), jy = min(iy + 1, dy);
    const float tx = fmaxf(0.0f, fminf(__fdiv_rn((float) rx, (float) cx), 1.0f));
    const float ty = fmaxf(0.0f, fminf(__fdiv_rn((float) ry, (float) cy), 1.0f));
    const float u00 = sample_at(pl, w, ix, iy), u01 = sample_at(pl, w, jx, iy);
    const float u10 = sample_at(pl, w, ix, jy), u11 = sample_at(pl, w, jx, jy);
    const float v0 = fadd(fmul(u00, fsub(1.0f, tx)), fmul(u01, tx));
    const float v1 = fadd(fmul(u10, fsub(1.0f, tx)), fmul(u11, tx));
    const float r = fadd(fmul(v0, fsub(1.0f, ty)), fmul(v1, ty));
    return (uint32_t) __float2int_rz(roundf(r));  // .rounded(): to nearest, ties away from zero
}

// jpeg.swift:441-453  YCbCr.rgb
__device__ __forceinline__ void ycc_to_rgb(uint32_t y8, uint32_t cb8, uint32_t cr8, float &r, float &g, float &b)
{
    const float Y = (float) (y8 & 0xffu), db = fsub((float) (cb8 & 0xffu), 128.0f),
                dr = fsub((float) (cr8 & 0xffu), 128.0f);
    r = fadd(fadd(Y, fmul(0.00000f, db)), fmul(1.40200f, dr));
    g = fadd(fadd(Y, fmul(-0.34414f, db)), fmul(-0.71414f, dr));
    b = fadd(fadd(Y, fmul(1.77200f, db)), fmul(0.00000f, dr));
}
__device__ __forceinline__ uint32_t clamp_u8(float v) { return (uint32_t) __float2int_rz(fmaxf(0.0f, fminf(v, 255.0f))); }

// ---- generic planes -> RGB8: any sampling factors, centred or co-sited; 4 pixels per thread -------------------
template <typename T>
__global__ void __launch_bounds__(256) k_planar_to_rgb8(const __grid_constant__ PlanarView V, uint8_t *__restrict__ rgb)
{
    const uint32_t groups_x = (uint32_t) (V.size_x + 3) / 4;
    const uint64_t per_image = (uint64_t) groups_x * V.size_y;
    const uint64_t total = per_image * V.n_images;
    const bool     word_rows = (V.size_x & 3) == 0;
    for (uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t) gridDim.x * blockDim.x) {
        const uint32_t img = (uint32_t) (i / per_image);
        const uint32_t rem = (uint32_t) (i - (uint64_t) img * per_image);
        const int      y = (int) (rem / groups_x);
        const int      x0 = (int) (rem - (uint32_t) y * groups_x) * 4;
        uint8_t       *dst = rgb + ((size_t) img * V.size_y + y) * (size_t) V.size_x * 3 + (size_t) x0 * 3;
        uint32_t       out[12];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int x = min(x0 + j, V.size_x - 1);
            uint32_t  Y = upsampled<T>(V, 0, img, x, y), cb = 128, cr = 128;
            if (V.n_planes == 3) {
                cb = upsampled<T>(V, 1, img, x, y);
                cr = upsampled<T>(V, 2, img, x, y);
            }
            float r, g, b;
            ycc_to_rgb(Y, cb, cr, r, g, b);
            out[3 * j] = clamp_u8(r);
            out[3 * j + 1] = clamp_u8(g);
            out[3 * j + 2] = clamp_u8(b);
        }
        if (word_rows) {
            uint32_t *d32 = reinterpret_cast<uint32_t *>(dst);
            d32[0] = out[0] | (out[1] << 8) | (out[2] << 16) | (out[3] << 24);
            d32[1] = out[4] | (out[5] << 8) | (out[6] << 16) | (out[7] << 24);
            d32[2] = out[8] | (out[9] << 8) | (out[10] << 16) | (out[11] << 24);
        } else {
            const int n = min(4, V.size_x - x0) * 3;
            for (int k = 0; k < n; ++k) dst[k] = (uint8_t) out[k];
        }
    }
}

// ---- fast paths: 8-bit planes -> RGB8 (ycc_to_rgb_fast, byte_of: pixel_core.cuh) ----------------------------------------------
// 8 pixels (24 bytes) of one row: Y bytes in (y0, y1), chroma as integers cb[8], cr[8]
__device__ __forceinline__ void emit_rgb8x8(uint2 yy, const int (&cb)[8], const int (&cr)[8], uint8_t *dst, bool vec,
                                            int n_valid)
{
    float r[8], g[8], b[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float Y = (float) byte_of(j < 4 ? yy.x : yy.y, j & 3);
        ycc_to_rgb_fast(Y, fsub((float) cb[j], 128.0f), fsub((float) cr[j], 128.0f), r[j], g[j], b[j]);
    }
    uint32_t w[6];
    w[0] = pack4_u8_trunc(r[0], g[0], b[0], r[1]);
    w[1] = pack4_u8_trunc(g[1], b[1], r[2], g[2]);
    w[2] = pack4_u8_trunc(b[2], r[3], g[3], b[3]);
    w[3] = pack4_u8_trunc(r[4], g[4], b[4], r[5]);
    w[4] = pack4_u8_trunc(g[5], b[5], r[6], g[6]);
    w[5] = pack4_u8_trunc(b[6], r[7], g[7], b[7]);
    if (vec) {
        uint2 *d = reinterpret_cast<uint2 *>(dst);
        d[0] = make_uint2(w[0], w[1]);
        d[1] = make_uint2(w[2], w[3]);
        d[2] = make_uint2(w[4], w[5]);
    } else {
#pragma unroll
        for (int q = 0; q < 24; ++q)
            if (q < n_valid * 3) dst[q] = (uint8_t) (w[q >> 2] >> (8 * (q & 3)));
    }
}

// 4:2:0, centred.  Thread = 8 x 2 pixels (columns 8t..8t+7 of rows 2r+1, 2r+2; row 0 is covered by r = -1).  For these
// factors every interpolation weight is a multiple of 1/4, every float product/sum in decode.swift:4258-4264 is exact,
// and `.rounded()` of the exact value equals (9a + 3b + 3c + d + 8) >> 4 -- integer arithmetic, bit-identical.
__global__ void __launch_bounds__(128)
k_ycc420_to_rgb8(const __grid_constant__ PlanarView V, uint8_t *__restrict__ rgb)
{
    const int      W = V.size_x, H = V.size_y;
    const int      groups_x = (W + 7) / 8;
    const int      row_pairs = H / 2 + 1;  // r = -1 .. ceil((H-1)/2)-1
    const uint64_t per_image = (uint64_t) groups_x * row_pairs;
    const uint64_t total = per_image * V.n_images;
    const int      yw = V.width[0], cw = V.width[1], ch = V.height[1];
    const bool     vec = (W & 7) == 0;
    for (uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t) gridDim.x * blockDim.x) {
        const uint32_t img = (uint32_t) (i / per_image);
        const uint32_t rem = (uint32_t) (i - (uint64_t) img * per_image);
        const int      rp = (int) (rem / groups_x);
        const int      gx = (int) (rem - (uint32_t) rp * groups_x);
        const int      r = rp - 1;  // chroma row pair (r, r+1) feeds luma rows 2r+1, 2r+2
        const int      x0 = 8 * gx, c0 = 4 * gx;
        const uint8_t *Yp = reinterpret_cast<const uint8_t *>(V.samples[0]) + (size_t) img * V.image_stride[0];
        const uint8_t *Cp[2] = {reinterpret_cast<const uint8_t *>(V.samples[1]) + (size_t) img * V.image_stride[1],
                                reinterpret_cast<const uint8_t *>(V.samples[2]) + (size_t) img * V.image_stride[2]};
        const int ra = max(r, 0), rb = min(r + 1, ch - 1);
        // horizontally interpolated chroma, scaled by 4, for the 8 columns, rows ra and rb
        int hc[2][2][8];
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) {
                const uint8_t *row = Cp[c] + (size_t) cw * (rr ? rb : ra);
                const uint32_t mid = __ldg(reinterpret_cast<const uint32_t *>(row + c0));  // chroma c0..c0+3
                const int      left = __ldg(row + max(c0 - 1, 0));
                const int      right = __ldg(row + min(c0 + 4, cw - 1));
                const int      s[6] = {left, (int) byte_of(mid, 0), (int) byte_of(mid, 1), (int) byte_of(mid, 2),
                                       (int) byte_of(mid, 3), right};
                // pixel x = 8g + j: even x = 2m uses (c[m-1], c[m]) weights (1, 3); odd x = 2m+1 uses (c[m], c[m+1]) weights (3, 1)
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int m = j >> 1;
                    hc[c][rr][j] = (j & 1) ? 3 * s[m + 1] + s[m + 2] : s[m] + 3 * s[m + 1];
                }
                if (x0 == 0) hc[c][rr][0] = 4 * s[1];  // x = 0: t clamps to 0 -> u[0] alone (decode.swift:4250)
            }
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const int y = 2 * r + 1 + k;
            if (y < 0 || y >= H) continue;
            // vertical weights: row 2r+1 -> (3, 1), row 2r+2 -> (1, 3); for y = 0 (r = -1) both rows are c[0],
            // which reproduces the reference's t = 0 clamp; at the bottom rb clamps to the padded plane edge
            const int   wa = (k == 0) ? 3 : 1, wb = 4 - wa;
            const uint2 yy = __ldg(reinterpret_cast<const uint2 *>(Yp + (size_t) yw * y + x0));
            int         cb[8], cr[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                cb[j] = (wa * hc[0][0][j] + wb * hc[0][1][j] + 8) >> 4;
                cr[j] = (wa * hc[1][0][j] + wb * hc[1][1][j] + 8) >> 4;
            }
            uint8_t *dst = rgb + ((size_t) img * H + y) * (size_t) W * 3 + (size_t) x0 * 3;
            emit_rgb8x8(yy, cb, cr, dst, vec, min(8, W - x0));
        }
    }
}

// ---- 4:2:0 centred, second generation: the same mapping as k_ycc420_to_rgb8 with a third fewer instructions --------------
// k_ycc420_to_rgb8 is bound by instruction issue (ncu: 63 % issue-active, 37 instructions per pixel, DRAM at 32 %), so this
// version does the integer interpolation two pixels at a time in packed 16-bit halves (every intermediate is < 4096):
//   horizontal  P = 3 * (s1 | s1 << 16) + (s0 | s2 << 16) + (2 | 2 << 16)      2 PRMT + 1 IMAD per pixel pair
//   vertical    V = ((3 * Pa + Pb) >> 4 & 0x00ff00ff) ^ 0x00800080             1 IMAD + 1 SHF + 1 LOP3 per pixel pair
// (3 Pa + Pb carries 3 * 2 + 2 = 8, the rounding term of (9a + 3b + 3c + d + 8) >> 4).  The final XOR turns each chroma byte
// into the two's-complement byte of (c - 128), so one signed byte -> float conversion yields the exact (c - 128.0f) the
// reference computes (jpeg.swift:447) and the two subtractions per pixel disappear.

__global__ void __launch_bounds__(256)
k_ycc420_to_rgb8_v2(const __grid_constant__ PlanarView V, uint8_t *__restrict__ rgb)
{
    // grid = (column groups / blockDim, row pairs (strided), images): no index divisions (a flat 64-bit index cost ~60 of the
    // ~450 instructions a thread spends on its 16 pixels)
    const int      W = V.size_x, H = V.size_y;
    const int      groups_x = (W + 7) / 8;
    const int      row_pairs = H / 2 + 1;  // r = -1 .. ceil((H-1)/2)-1
    const int      yw = V.width[0], cw = V.width[1], ch = V.height[1];
    const bool     vec = (W & 7) == 0;
    const int      gx = (int) (blockIdx.x * blockDim.x + threadIdx.x);
    const uint32_t img = blockIdx.z;
    if (gx >= groups_x) return;
    for (int rp = (int) blockIdx.y; rp < row_pairs; rp += (int) gridDim.y) {
        const int      r = rp - 1;  // chroma row pair (r, r+1) feeds luma rows 2r+1, 2r+2
        const int      x0 = 8 * gx, c0 = 4 * gx;
        const uint8_t *Yp = reinterpret_cast<const uint8_t *>(V.samples[0]) + (size_t) img * V.image_stride[0];
        const int      ra = max(r, 0), rb = min(r + 1, ch - 1);
        // P[c][rr][m]: horizontally interpolated chroma (x 4, + 2) of pixels 2m (low half) and 2m + 1 (high half)
        uint32_t P[2][2][4];
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const uint8_t *Cp = reinterpret_cast<const uint8_t *>(V.samples[1 + c]) + (size_t) img * V.image_stride[1 + c];
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) {
                const uint8_t *row = Cp + (size_t) cw * (rr ? rb : ra);
                const uint32_t mid = __ldg(reinterpret_cast<const uint32_t *>(row + c0));  // chroma c0 .. c0+3
                // x = 0 clamps t to 0 (decode.swift:4250): u[0] alone, which is what left = row[0] gives; the right neighbour
                // clamps to the padded plane's last column (decode.swift:4244)
                const uint32_t left = __ldg(row + max(c0 - 1, 0)), right = __ldg(row + min(c0 + 4, cw - 1));
                // pixel 2m: (c[m-1], c[m]) weights (1, 3); pixel 2m+1: (c[m], c[m+1]) weights (3, 1)
                P[c][rr][0] = __byte_perm(mid, 0, 0x4040) * 3u + (__byte_perm(mid, left, 0x5154) + 0x00020002u);
                P[c][rr][1] = __byte_perm(mid, 0, 0x4141) * 3u + (__byte_perm(mid, 0, 0x4240) + 0x00020002u);
                P[c][rr][2] = __byte_perm(mid, 0, 0x4242) * 3u + (__byte_perm(mid, 0, 0x4341) + 0x00020002u);
                P[c][rr][3] = __byte_perm(mid, 0, 0x4343) * 3u + (__byte_perm(mid, right, 0x5452) + 0x00020002u);
            }
        }
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const int y = 2 * r + 1 + k;
            if (y < 0 || y >= H) continue;
            // vertical weights: row 2r+1 -> (3, 1), row 2r+2 -> (1, 3); for y = 0 (r = -1) both rows are c[0], which reproduces
            // the reference's t = 0 clamp; at the bottom rb clamps to the padded plane edge
            const uint2 yy = __ldg(reinterpret_cast<const uint2 *>(Yp + (size_t) yw * y + x0));
            float       rr_[8], gg_[8], bb_[8];
#pragma unroll
            for (int m = 0; m < 4; ++m) {
                const uint32_t tb = (k == 0) ? P[0][0][m] * 3u + P[0][1][m] : P[0][1][m] * 3u + P[0][0][m];
                const uint32_t tr = (k == 0) ? P[1][0][m] * 3u + P[1][1][m] : P[1][1][m] * 3u + P[1][0][m];
                const uint32_t vb = ((tb >> 4) & 0x00ff00ffu) ^ 0x00800080u, vr = ((tr >> 4) & 0x00ff00ffu) ^ 0x00800080u;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int   j = 2 * m + h;
                    const float Y = (float) byte_of(j < 4 ? yy.x : yy.y, j & 3);
                    ycc_to_rgb_fast(Y, s8_to_float(vb, 2 * h), s8_to_float(vr, 2 * h), rr_[j], gg_[j], bb_[j]);
                }
            }
            uint32_t w[6];
            w[0] = pack4_u8_trunc(rr_[0], gg_[0], bb_[0], rr_[1]);
            w[1] = pack4_u8_trunc(gg_[1], bb_[1], rr_[2], gg_[2]);
            w[2] = pack4_u8_trunc(bb_[2], rr_[3], gg_[3], bb_[3]);
            w[3] = pack4_u8_trunc(rr_[4], gg_[4], bb_[4], rr_[5]);
            w[4] = pack4_u8_trunc(gg_[5], bb_[5], rr_[6], gg_[6]);
            w[5] = pack4_u8_trunc(bb_[6], rr_[7], gg_[7], bb_[7]);
            uint8_t *dst = rgb + ((size_t) img * H + y) * (size_t) W * 3 + (size_t) x0 * 3;
            if (vec) {
                uint2 *d = reinterpret_cast<uint2 *>(dst);
                d[0] = make_uint2(w[0], w[1]);
                d[1] = make_uint2(w[2], w[3]);
                d[2] = make_uint2(w[4], w[5]);
            } else {
                const int n_valid = min(8, W - x0);
#pragma unroll
                for (int q = 0; q < 24; ++q)
                    if (q < n_valid * 3) dst[q] = (uint8_t) (w[q >> 2] >> (8 * (q & 3)));
            }
        }
    }
}

// ---- the same two fast paths with the RGB rows staged in shared memory and written by the TMA (SASS: UBLKCP) ------------
// A thread's 24 bytes per row sit 24 bytes apart: as direct stores that is three 8-byte stores per row per thread, each warp
// store touching every sector of a 768-byte span a third at a time (partial-sector writes, 2.7 TB/s).  Staged, a tile row
// of up to 1024 pixels (3072 bytes) leaves as ONE bulk copy shared -> global: whole lines, no LSU store traffic.
// Requires W % 16 == 0 (row segments are multiples of 16 bytes) and a 16-byte aligned RGB base.
constexpr int RGB_TILE_PX = 1024;  // 128 threads x 8 pixels

__device__ __forceinline__ void bulk_store(void *gdst, const void *ssrc, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read()
{
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}

// ROWS = 2: 4:2:0 centred (thread = 8 x 2 pixels, rows 2r+1 and 2r+2 as in k_ycc420_to_rgb8); ROWS = 1: 4:4:4
template <int ROWS>
__global__ void __launch_bounds__(128)
k_ycc_to_rgb8_tma(const __grid_constant__ PlanarView V, uint8_t *__restrict__ rgb)
{
    __shared__ __align__(128) uint8_t sbuf[2][ROWS][RGB_TILE_PX * 3];
    const int      W = V.size_x, H = V.size_y;
    const int      tiles_x = (W + RGB_TILE_PX - 1) / RGB_TILE_PX;
    const int      rows_y = ROWS == 2 ? H / 2 + 1 : H;  // ROWS == 2: r = -1 .. ceil((H-1)/2)-1
    const uint64_t per_image = (uint64_t) tiles_x * rows_y;
    const uint64_t total = per_image * V.n_images;
    const int      yw = V.width[0], cw = V.width[1], ch = V.height[1];
    const int      tid = threadIdx.x;
    int            stage = 0;
    for (uint64_t t = blockIdx.x; t < total; t += gridDim.x, stage ^= 1) {
        const uint32_t img = (uint32_t) (t / per_image);
        const uint32_t rem = (uint32_t) (t - (uint64_t) img * per_image);
        const int      ry = (int) (rem / tiles_x), tx = (int) (rem - (uint32_t) ry * tiles_x);
        const int      x0 = tx * RGB_TILE_PX + 8 * tid;
        const int      tile_px = min(RGB_TILE_PX, W - tx * RGB_TILE_PX);
        // the bulk copies that read this stage two tiles ago must be done with it (one younger group may still be in flight)
        if (tid == 0) bulk_wait_read<1>();
        __syncthreads();
        const uint8_t *Yp = reinterpret_cast<const uint8_t *>(V.samples[0]) + (size_t) img * V.image_stride[0];
        const uint8_t *Cp[2] = {reinterpret_cast<const uint8_t *>(V.samples[1]) + (size_t) img * V.image_stride[1],
                                reinterpret_cast<const uint8_t *>(V.samples[2]) + (size_t) img * V.image_stride[2]};
        if (ROWS == 2) {
            // chroma row pair (r, r+1) feeds luma rows 2r+1, 2r+2.  Every lane runs the loads and shuffles (lanes past the
            // right edge of the image read a harmless in-plane word); only lanes inside the image convert and store.
            const int  r = ry - 1, c0 = min(x0 >> 1, cw - 4);
            const bool valid = x0 < W;
            const int  ra = max(r, 0), rb = min(r + 1, ch - 1);
            int        hc[2][2][8];  // horizontally interpolated chroma, scaled by 4 (see k_ycc420_to_rgb8)
#pragma unroll
            for (int c = 0; c < 2; ++c)
#pragma unroll
                for (int rr = 0; rr < 2; ++rr) {
                    const uint8_t *row = Cp[c] + (size_t) cw * (rr ? rb : ra);
                    const uint32_t mid = __ldg(reinterpret_cast<const uint32_t *>(row + c0));
                    // neighbours: the adjacent lanes hold them; only the lanes at the ends of a warp go to memory
                    uint32_t lt = __shfl_up_sync(0xffffffffu, mid, 1) >> 24, rt = __shfl_down_sync(0xffffffffu, mid, 1) & 0xffu;
                    if ((tid & 31) == 0) lt = __ldg(row + max(c0 - 1, 0));
                    if ((tid & 31) == 31 || c0 + 4 > cw - 1) rt = __ldg(row + min(c0 + 4, cw - 1));  // plane edge: clamp (decode.swift:4244)
                    const int s[6] = {(int) lt, (int) byte_of(mid, 0), (int) byte_of(mid, 1), (int) byte_of(mid, 2),
                                      (int) byte_of(mid, 3), (int) rt};
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int m = j >> 1;
                        hc[c][rr][j] = (j & 1) ? 3 * s[m + 1] + s[m + 2] : s[m] + 3 * s[m + 1];
                    }
                    if (x0 == 0) hc[c][rr][0] = 4 * s[1];
                }
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const int y = 2 * r + 1 + k;
                if (y < 0 || y >= H || !valid) continue;
                const int   wa = (k == 0) ? 3 : 1, wb = 4 - wa;
                const uint2 yy = __ldg(reinterpret_cast<const uint2 *>(Yp + (size_t) yw * y + x0));
                int         cb[8], cr[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    cb[j] = (wa * hc[0][0][j] + wb * hc[0][1][j] + 8) >> 4;
                    cr[j] = (wa * hc[1][0][j] + wb * hc[1][1][j] + 8) >> 4;
                }
                emit_rgb8x8(yy, cb, cr, &sbuf[stage][k][24 * tid], true, 8);
            }
        } else {
            if (x0 < W) {
                uint2 plv[3];
#pragma unroll
                for (int p = 0; p < 3; ++p) {
                    const uint8_t *base = reinterpret_cast<const uint8_t *>(V.samples[p]) + (size_t) img * V.image_stride[p];
                    plv[p] = __ldg(reinterpret_cast<const uint2 *>(base + (size_t) V.width[p] * ry + x0));
                }
                int cb[8], cr[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    cb[j] = (int) byte_of(j < 4 ? plv[1].x : plv[1].y, j & 3);
                    cr[j] = (int) byte_of(j < 4 ? plv[2].x : plv[2].y, j & 3);
                }
                emit_rgb8x8(plv[0], cb, cr, &sbuf[stage][0][24 * tid], true, 8);
            }
        }
        fence_proxy_async();  // generic-proxy writes to shared memory -> visible to the bulk copy engine
        __syncthreads();
        if (tid == 0) {
#pragma unroll
            for (int k = 0; k < ROWS; ++k) {
                const int y = ROWS == 2 ? 2 * (ry - 1) + 1 + k : ry;
                if (y < 0 || y >= H) continue;
                bulk_store(rgb + (((size_t) img * H + y) * (size_t) W + (size_t) tx * RGB_TILE_PX) * 3, &sbuf[stage][k][0],
                           (uint32_t) tile_px * 3u);
            }
            bulk_commit();
        }
    }
    if (tid == 0) bulk_wait_read<0>();  // shared memory must outlive the copies that read it
}

// 4:4:4 (and any layout whose planes all have factor == scale): no resampling; thread = 8 pixels of one row
__global__ void __launch_bounds__(256)
k_ycc444_to_rgb8(const __grid_constant__ PlanarView V, uint8_t *__restrict__ rgb)
{
    // grid = (column groups / blockDim, rows (strided), images): no index divisions (see k_ycc420_to_rgb8_v2)
    const int      W = V.size_x, H = V.size_y;
    const int      groups_x = (W + 7) / 8;
    const bool     vec = (W & 7) == 0;
    const int      gx = (int) (blockIdx.x * blockDim.x + threadIdx.x);
    const uint32_t img = blockIdx.z;
    if (gx >= groups_x) return;
    for (int y = (int) blockIdx.y; y < H; y += (int) gridDim.y) {
        const int x0 = 8 * gx;
        uint2     pl[3];
#pragma unroll
        for (int p = 0; p < 3; ++p) {
            const uint8_t *base = reinterpret_cast<const uint8_t *>(V.samples[p]) + (size_t) img * V.image_stride[p];
            pl[p] = __ldg(reinterpret_cast<const uint2 *>(base + (size_t) V.width[p] * y + x0));
        }
        int cb[8], cr[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            cb[j] = (int) byte_of(j < 4 ? pl[1].x : pl[1].y, j & 3);
            cr[j] = (int) byte_of(j < 4 ? pl[2].x : pl[2].y, j & 3);
        }
        uint8_t *dst = rgb + ((size_t) img * H + y) * (size_t) W * 3 + (size_t) x0 * 3;
        emit_rgb8x8(pl[0], cb, cr, dst, vec, min(8, W - x0));
    }
}

// ---- planes -> interleaved uint16 (Rectangular.values) ------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
k_interleave(const __grid_constant__ PlanarView V, uint16_t *__restrict__ out)
{
    const uint64_t per_image = (uint64_t) V.size_x * V.size_y;
    const uint64_t total = per_image * V.n_images;
    for (uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t) gridDim.x * blockDim.x) {
        const uint32_t img = (uint32_t) (i / per_image);
        const uint32_t rem = (uint32_t) (i - (uint64_t) img * per_image);
        const int      y = (int) (rem / (uint32_t) V.size_x), x = (int) (rem - (uint32_t) y * V.size_x);
        for (int p = 0; p < V.n_planes; ++p) out[i * V.n_planes + p] = (uint16_t) upsampled<T>(V, p, img, x, y);
    }
}

// jpeg.swift:551-572 RGB.unpack / 493-513 YCbCr.unpack
template <bool TO_RGB>
__global__ void __launch_bounds__(256)
k_unpack(const uint16_t *__restrict__ il, uint64_t n_px, int arity, uint8_t *__restrict__ out)
{
    for (uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x; i < n_px; i += (uint64_t) gridDim.x * blockDim.x) {
        uint32_t Y, cb = 128, cr = 128;
        if (arity == 1) Y = il[i];
        else {
            Y = il[3 * i];
            cb = il[3 * i + 1];
            cr = il[3 * i + 2];
        }
        if (TO_RGB) {
            float r, g, b;
            ycc_to_rgb(Y, cb, cr, r, g, b);
            out[3 * i] = (uint8_t) clamp_u8(r);
            out[3 * i + 1] = (uint8_t) clamp_u8(g);
            out[3 * i + 2] = (uint8_t) clamp_u8(b);
        } else {
            out[3 * i] = (uint8_t) Y;
            out[3 * i + 1] = (uint8_t) cb;
            out[3 * i + 2] = (uint8_t) cr;
        }
    }
}

// ---- encode side ---------------------------------------------------------------------------------------------------
// jpeg.swift:463-478 RGB.ycc
__device__ __forceinline__ void rgb_to_ycc(uint32_t R8, uint32_t G8, uint32_t B8, uint32_t &y, uint32_t &cb, uint32_t &cr)
{
    const float R = (float) R8, G = (float) G8, B = (float) B8;
    const float fy = fadd(fadd(fadd(0.0f, fmul(0.2990f, R)), fmul(0.5870f, G)), fmul(0.1140f, B));
    const float fb = fadd(fadd(fadd(128.0f, fmul(-0.1687f, R)), fmul(-0.3313f, G)), fmul(0.5000f, B));
    const float fr = fadd(fadd(fadd(128.0f, fmul(0.5000f, R)), fmul(-0.4187f, G)), fmul(-0.0813f, B));
    y = clamp_u8(fy);
    cb = clamp_u8(fb);
    cr = clamp_u8(fr);
}

// jpeg.swift:584-599 RGB.pack
__global__ void __launch_bounds__(256)
k_pack_rgb(const uint8_t *__restrict__ rgb, uint64_t n_px, int arity, uint16_t *__restrict__ il)
{
    for (uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x; i < n_px; i += (uint64_t) gridDim.x * blockDim.x) {
        uint32_t y, cb, cr;
        rgb_to_ycc(rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2], y, cb, cr);
        if (arity == 1) il[i] = (uint16_t) y;
        else {
            il[3 * i] = (uint16_t) y;
            il[3 * i + 1] = (uint16_t) cb;
            il[3 * i + 2] = (uint16_t) cr;
        }
    }
}

// encode.swift:389-425 Rectangular.decomposed(): box filter with edge clamp, truncating mean, over the padded plane.
// SRC_RGB8: the source is RGB8 and the colour conversion is fused (pack + decomposed in one pass: K4).
template <typename T, bool SRC_RGB8>
__global__ void __launch_bounds__(256)
k_decompose(const void *__restrict__ src, const __grid_constant__ PlanarView V, int p)
{
    const int      w = V.width[p], h = V.height[p];
    const uint64_t per_image = (uint64_t) w * h;
    const uint64_t total = per_image * V.n_images;
    const int      rx = V.scale_x / V.fx[p], ry = V.scale_y / V.fy[p];
    const float    magnitude = (float) (rx * ry);
    T             *out = reinterpret_cast<T *>(const_cast<void *>(V.samples[p]));
    for (uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t) gridDim.x * blockDim.x) {
        const uint32_t img = (uint32_t) (i / per_image);
        const uint32_t rem = (uint32_t) (i - (uint64_t) img * per_image);
        const int      y = (int) (rem / (uint32_t) w), x = (int) (rem - (uint32_t) y * w);
        const int      bx = x * V.scale_x / V.fx[p], by = y * V.scale_y / V.fy[p];
        int            sum = 0;
        for (int yy = by; yy < by + ry; ++yy)
            for (int xx = bx; xx < bx + rx; ++xx) {
                const int    ix = min(xx, V.size_x - 1), iy = min(yy, V.size_y - 1);
                const size_t px = ((size_t) img * V.size_y + iy) * V.size_x + ix;
                if (SRC_RGB8) {
                    const uint8_t *s = reinterpret_cast<const uint8_t *>(src) + 3 * px;
                    uint32_t       c[3];
                    rgb_to_ycc(s[0], s[1], s[2], c[0], c[1], c[2]);
                    sum += (int) c[V.n_planes == 1 ? 0 : p];
                } else
                    sum += reinterpret_cast<const uint16_t *>(src)[px * V.n_planes + p];
            }
        out[(size_t) img * V.image_stride[p] + (size_t) w * y + x] = (T) __float2int_rz(__fdiv_rn((float) sum, magnitude));
    }
}

// K4 fast paths: RGB8 -> 8-bit YCbCr planes in one pass (RGB.pack jpeg.swift:584-599 fused with decomposed() encode.swift:389-425)
// for 4:2:0 (SUB = true: thread = 8 x 2 pixels -> 16 Y, 4 Cb, 4 Cr) and 4:4:4 (thread = 8 pixels), image sizes that are whole
// MCUs (so the planes have no padding to fill and no edge to clamp).  Each pixel is converted to integer YCbCr first, exactly
// as RGB.pack does; the 2 x 2 box mean UInt16(Float(sum) / Float(4)) of encode.swift:413-420 is exact and truncating: sum >> 2.
// The generic kernel reads every RGB pixel three times (once per plane) with 1-byte loads: 8.0 ms per 64 4K frames; this
// reads them once with 8-byte loads.
template <bool SUB>
__global__ void __launch_bounds__(256)
k_rgb8_to_ycc_planes(const uint8_t *__restrict__ rgb, const __grid_constant__ PlanarView V)
{
    constexpr int  ROWS = SUB ? 2 : 1;
    // grid = (column groups / blockDim, rows (strided), images): no index divisions (see k_ycc420_to_rgb8_v2)
    const int      W = V.size_x, H = V.size_y;
    const int      groups_x = W / 8, rows_y = H / ROWS;
    const int      gx = (int) (blockIdx.x * blockDim.x + threadIdx.x);
    const uint32_t img = blockIdx.z;
    if (gx >= groups_x) return;
    for (int ry = (int) blockIdx.y; ry < rows_y; ry += (int) gridDim.y) {
        const int      x0 = 8 * gx, y0 = ROWS * ry;
        uint32_t       cbs[ROWS][8], crs[ROWS][8];
#pragma unroll
        for (int k = 0; k < ROWS; ++k) {
            const uint2 *src = reinterpret_cast<const uint2 *>(rgb + (((size_t) img * H + y0 + k) * W + x0) * 3);
            const uint2  a = __ldg(src), b = __ldg(src + 1), c = __ldg(src + 2);
            const uint32_t w[6] = {a.x, a.y, b.x, b.y, c.x, c.y};
            uint32_t       yv[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const uint32_t R = byte_of(w[(3 * j) >> 2], (3 * j) & 3), G = byte_of(w[(3 * j + 1) >> 2], (3 * j + 1) & 3),
                               B = byte_of(w[(3 * j + 2) >> 2], (3 * j + 2) & 3);
                rgb_to_ycc(R, G, B, yv[j], cbs[k][j], crs[k][j]);
            }
            uint8_t *Yp = reinterpret_cast<uint8_t *>(const_cast<void *>(V.samples[0])) + (size_t) img * V.image_stride[0];
            *reinterpret_cast<uint2 *>(Yp + (size_t) V.width[0] * (y0 + k) + x0) =
                make_uint2(yv[0] | yv[1] << 8 | yv[2] << 16 | yv[3] << 24, yv[4] | yv[5] << 8 | yv[6] << 16 | yv[7] << 24);
        }
        uint8_t *Cb = reinterpret_cast<uint8_t *>(const_cast<void *>(V.samples[1])) + (size_t) img * V.image_stride[1];
        uint8_t *Cr = reinterpret_cast<uint8_t *>(const_cast<void *>(V.samples[2])) + (size_t) img * V.image_stride[2];
        if (SUB) {
            uint32_t pb = 0, pr = 0;
#pragma unroll
            for (int m = 0; m < 4; ++m) {
                pb |= ((cbs[0][2 * m] + cbs[0][2 * m + 1] + cbs[ROWS - 1][2 * m] + cbs[ROWS - 1][2 * m + 1]) >> 2) << (8 * m);
                pr |= ((crs[0][2 * m] + crs[0][2 * m + 1] + crs[ROWS - 1][2 * m] + crs[ROWS - 1][2 * m + 1]) >> 2) << (8 * m);
            }
            *reinterpret_cast<uint32_t *>(Cb + (size_t) V.width[1] * ry + 4 * gx) = pb;
            *reinterpret_cast<uint32_t *>(Cr + (size_t) V.width[2] * ry + 4 * gx) = pr;
        } else {
            *reinterpret_cast<uint2 *>(Cb + (size_t) V.width[1] * y0 + x0) =
                make_uint2(cbs[0][0] | cbs[0][1] << 8 | cbs[0][2] << 16 | cbs[0][3] << 24, cbs[0][4] | cbs[0][5] << 8 | cbs[0][6] << 16 | cbs[0][7] << 24);
            *reinterpret_cast<uint2 *>(Cr + (size_t) V.width[2] * y0 + x0) =
                make_uint2(crs[0][0] | crs[0][1] << 8 | crs[0][2] << 16 | crs[0][3] << 24, crs[0][4] | crs[0][5] << 8 | crs[0][6] << 16 | crs[0][7] << 24);
        }
    }
}

// threads per CTA for the kernels whose CTAs span a row of 8-pixel groups: the multiple of 32 (64..256) that wastes the fewest
// lanes on the last CTA of the row
inline uint32_t row_threads(uint32_t groups_x)
{
    uint32_t bt = 128, best = ~0u;
    for (uint32_t t = 256; t >= 64; t -= 32) {
        const uint32_t waste = (groups_x + t - 1) / t * t - groups_x;
        if (waste < best) best = waste, bt = t;
    }
    return bt;
}

int fill_view(const jpeg_sm100_dev_planar *pl, uint32_t sx, uint32_t sy, int cosited, PlanarView &V)
{
    if (!pl || pl->n_planes < 1 || pl->n_planes > 4) return JPEG_SM100_ERR_INVALID_ARGUMENT;
    if (pl->sample_bytes != 1 && pl->sample_bytes != 2) return JPEG_SM100_ERR_INVALID_ARGUMENT;
    memset(&V, 0, sizeof V);
    V.n_planes = (int) pl->n_planes;
    V.n_images = pl->n_images;
    V.size_x = (int) sx;
    V.size_y = (int) sy;
    V.cosited = cosited ? 1 : 0;
    for (uint32_t p = 0; p < pl->n_planes; ++p) {
        V.samples[p] = pl->plane[p].samples;
        V.image_stride[p] = pl->plane[p].image_stride;
        V.width[p] = 8 * pl->plane[p].units_x;
        V.height[p] = 8 * pl->plane[p].units_y;
        V.fx[p] = pl->plane[p].factor_x;
        V.fy[p] = pl->plane[p].factor_y;
        if (V.fx[p] < 1 || V.fy[p] < 1) return JPEG_SM100_ERR_INVALID_ARGUMENT;
        V.scale_x = V.fx[p] > V.scale_x ? V.fx[p] : V.scale_x;
        V.scale_y = V.fy[p] > V.scale_y ? V.fy[p] : V.scale_y;
    }
    // every pixel of the size_x x size_y rectangle must map into its plane (the reference indexes Planar.Plane with bounds
    // checks and would trap, decode.swift:1586-1597): ceil(size * factor / scale) <= 8 * units
    for (uint32_t p = 0; p < pl->n_planes; ++p) {
        const int64_t need_x = ((int64_t) sx * V.fx[p] + V.scale_x - 1) / V.scale_x, need_y = ((int64_t) sy * V.fy[p] + V.scale_y - 1) / V.scale_y;
        if (need_x > V.width[p] || need_y > V.height[p]) return JPEG_SM100_ERR_PRECONDITION;
    }
    return JPEG_SM100_OK;
}

inline uint32_t grid_for(jpeg_sm100_ctx *ctx, uint64_t work, int block, int per_sm)
{
    uint64_t g = (work + block - 1) / block;
    uint64_t cap = (uint64_t) ctx->sm_count * per_sm;
    if (g > cap) g = cap;
    return g ? (uint32_t) g : 1u;
}

}  // namespace

int jpeg_color_planar_to_rgb8(jpeg_sm100_ctx *ctx, const jpeg_sm100_dev_planar *pl, uint32_t sx, uint32_t sy,
                              int cosited, uint8_t *d_rgb)
{
    PlanarView V;
    J_TRY(fill_view(pl, sx, sy, cosited, V));
    if (V.n_planes != 1 && V.n_planes != 3) return JPEG_SM100_ERR_UNSUPPORTED;
    if ((uint64_t) sx * sy * pl->n_images == 0) return JPEG_SM100_OK;
    const char *color_env = getenv("JPEG_SM100_COLOR");  // A/B validation: "generic" (reference-literal kernels), "tma" (bulk-copy stores)
    const bool  no_fast = color_env && strcmp(color_env, "generic") == 0;
    const bool is420 = V.n_planes == 3 && pl->sample_bytes == 1 && !cosited && V.fx[0] == 2 && V.fy[0] == 2 &&
                       V.fx[1] == 1 && V.fy[1] == 1 && V.fx[2] == 1 && V.fy[2] == 1 &&
                       V.width[1] == V.width[2] && V.height[1] == V.height[2] &&
                       (reinterpret_cast<uintptr_t>(V.samples[0]) & 7) == 0 && (V.image_stride[0] & 7) == 0 &&
                       (reinterpret_cast<uintptr_t>(V.samples[1]) & 3) == 0 && (V.image_stride[1] & 3) == 0 &&
                       (reinterpret_cast<uintptr_t>(V.samples[2]) & 3) == 0 && (V.image_stride[2] & 3) == 0 &&
                       (reinterpret_cast<uintptr_t>(d_rgb) & 7) == 0;
    bool is444 = V.n_planes == 3 && pl->sample_bytes == 1 && (reinterpret_cast<uintptr_t>(d_rgb) & 7) == 0;
    for (int p = 0; p < 3 && is444; ++p)
        is444 = V.fx[p] == V.scale_x && V.fy[p] == V.scale_y && (reinterpret_cast<uintptr_t>(V.samples[p]) & 7) == 0 &&
                (V.image_stride[p] & 7) == 0;
    // rows staged in shared memory and written by bulk copies: needs 16-byte row segments
    // measured on B200 (4K 4:2:0, 64 frames): direct stores 0.87 ms, staged + bulk copies 0.91 ms -- the kernel is bound by
    // instruction issue (63 % issue-active, 37 instructions per pixel), not by its stores; the TMA variant is opt-in ("tma")
    const bool no_tma = !(color_env && strcmp(color_env, "tma") == 0);
    const bool tma_ok = !no_tma && (sx % 16) == 0 && (reinterpret_cast<uintptr_t>(d_rgb) & 15) == 0;
    if ((is444 || is420) && !no_fast && tma_ok) {
        const uint64_t tiles = (uint64_t) ((sx + RGB_TILE_PX - 1) / RGB_TILE_PX) * (is444 ? sy : sy / 2 + 1) * pl->n_images;
        if (is444) k_ycc_to_rgb8_tma<1><<<grid_for(ctx, tiles * 128, 128, 8), 128, 0, ctx->stream>>>(V, d_rgb);
        else k_ycc_to_rgb8_tma<2><<<grid_for(ctx, tiles * 128, 128, 8), 128, 0, ctx->stream>>>(V, d_rgb);
        LAUNCH_CHECK(ctx);
        return JPEG_SM100_OK;
    }
    if (is444 && !no_fast) {
        const uint32_t groups_x = (sx + 7) / 8, bt = row_threads(groups_x);
        for (uint32_t i0 = 0; i0 < pl->n_images; i0 += 65535) {
            PlanarView     Vz = V;
            const uint32_t nz = std::min<uint32_t>(65535, pl->n_images - i0);
            for (int p = 0; p < 3; ++p)
                Vz.samples[p] = reinterpret_cast<const uint8_t *>(V.samples[p]) + (size_t) i0 * V.image_stride[p];
            const dim3 grid((groups_x + bt - 1) / bt, std::min<uint32_t>(sy, 4096), nz);
            k_ycc444_to_rgb8<<<grid, bt, 0, ctx->stream>>>(Vz, d_rgb + (size_t) i0 * sx * sy * 3);
        }
        LAUNCH_CHECK(ctx);
        return JPEG_SM100_OK;
    }
    if (is420 && !no_fast) {
        const uint64_t work = (uint64_t) ((sx + 7) / 8) * (sy / 2 + 1) * pl->n_images;
        if (color_env && strcmp(color_env, "direct") == 0)  // first generation, kept for A/B validation
            k_ycc420_to_rgb8<<<grid_for(ctx, work, 128, 16), 128, 0, ctx->stream>>>(V, d_rgb);
        else {
            const uint32_t groups_x = (sx + 7) / 8, row_pairs = sy / 2 + 1, bt = row_threads(groups_x);
            for (uint32_t i0 = 0; i0 < pl->n_images; i0 += 65535) {  // gridDim.z limit
                PlanarView     Vz = V;
                const uint32_t nz = std::min<uint32_t>(65535, pl->n_images - i0);
                for (int p = 0; p < 3; ++p)
                    Vz.samples[p] = reinterpret_cast<const uint8_t *>(V.samples[p]) + (size_t) i0 * V.image_stride[p];
                const dim3 grid((groups_x + bt - 1) / bt, std::min<uint32_t>(row_pairs, 4096), nz);
                k_ycc420_to_rgb8_v2<<<grid, bt, 0, ctx->stream>>>(Vz, d_rgb + (size_t) i0 * sx * sy * 3);
            }
        }
        LAUNCH_CHECK(ctx);
        return JPEG_SM100_OK;
    }
    const uint64_t work = (uint64_t) ((sx + 3) / 4) * sy * pl->n_images;
    if (pl->sample_bytes == 1) k_planar_to_rgb8<uint8_t><<<grid_for(ctx, work, 256, 8), 256, 0, ctx->stream>>>(V, d_rgb);
    else k_planar_to_rgb8<uint16_t><<<grid_for(ctx, work, 256, 8), 256, 0, ctx->stream>>>(V, d_rgb);
    LAUNCH_CHECK(ctx);
    return JPEG_SM100_OK;
}

int jpeg_color_interleave(jpeg_sm100_ctx *ctx, const jpeg_sm100_dev_planar *pl, uint32_t sx, uint32_t sy, int cosited,
                          uint16_t *d_out)
{
    PlanarView V;
    J_TRY(fill_view(pl, sx, sy, cosited, V));
    const uint64_t work = (uint64_t) sx * sy * pl->n_images;
    if (work == 0) return JPEG_SM100_OK;
    if (pl->sample_bytes == 1) k_interleave<uint8_t><<<grid_for(ctx, work, 256, 8), 256, 0, ctx->stream>>>(V, d_out);
    else k_interleave<uint16_t><<<grid_for(ctx, work, 256, 8), 256, 0, ctx->stream>>>(V, d_out);
    LAUNCH_CHECK(ctx);
    return JPEG_SM100_OK;
}

int jpeg_color_unpack(jpeg_sm100_ctx *ctx, const uint16_t *d_il, uint64_t n_px, int arity, uint8_t *d_out, bool to_rgb)
{
    if (arity != 1 && arity != 3) return JPEG_SM100_ERR_INVALID_ARGUMENT;
    if (n_px == 0) return JPEG_SM100_OK;
    if (to_rgb) k_unpack<true><<<grid_for(ctx, n_px, 256, 8), 256, 0, ctx->stream>>>(d_il, n_px, arity, d_out);
    else k_unpack<false><<<grid_for(ctx, n_px, 256, 8), 256, 0, ctx->stream>>>(d_il, n_px, arity, d_out);
    LAUNCH_CHECK(ctx);
    return JPEG_SM100_OK;
}

int jpeg_color_pack_rgb(jpeg_sm100_ctx *ctx, const uint8_t *d_rgb, uint64_t n_px, int arity, uint16_t *d_il)
{
    if (arity != 1 && arity != 3) return JPEG_SM100_ERR_INVALID_ARGUMENT;
    if (n_px == 0) return JPEG_SM100_OK;
    k_pack_rgb<<<grid_for(ctx, n_px, 256, 8), 256, 0, ctx->stream>>>(d_rgb, n_px, arity, d_il);
    LAUNCH_CHECK(ctx);
    return JPEG_SM100_OK;
}

// src_is_rgb8: fused colour conversion (K4) else uint16 interleaved source (plain decomposed())
int jpeg_color_decompose(jpeg_sm100_ctx *ctx, const void *d_src, bool src_is_rgb8, uint32_t sx, uint32_t sy,
                         const jpeg_sm100_dev_planar *pl)
{
    PlanarView V;
    J_TRY(fill_view(pl, sx, sy, 0, V));
    if (src_is_rgb8 && V.n_planes != 1 && V.n_planes != 3) return JPEG_SM100_ERR_UNSUPPORTED;
    {   // fused fast paths (see k_rgb8_to_ycc_planes)
        const char *e = getenv("JPEG_SM100_COLOR");
        const bool  generic = e && strcmp(e, "generic") == 0;
        bool        ok = src_is_rgb8 && !generic && V.n_planes == 3 && pl->sample_bytes == 1 && pl->n_images > 0 &&
                  (reinterpret_cast<uintptr_t>(d_src) & 7) == 0 && V.fx[1] == 1 && V.fy[1] == 1 && V.fx[2] == 1 && V.fy[2] == 1;
        const bool sub = ok && V.fx[0] == 2 && V.fy[0] == 2, full = ok && V.fx[0] == 1 && V.fy[0] == 1;
        const int  mx = sub ? 16 : 8;
        ok = (sub || full) && sx % mx == 0 && sy % mx == 0 && sx > 0 && sy > 0;
        for (int p = 0; p < 3 && ok; ++p)
            ok = V.width[p] == (int) (sx * V.fx[p] / V.scale_x) && V.height[p] == (int) (sy * V.fy[p] / V.scale_y) &&
                 (reinterpret_cast<uintptr_t>(V.samples[p]) & 7) == 0 && (V.image_stride[p] & 7) == 0;
        if (ok) {
            const uint32_t groups_x = sx / 8, rows_y = sy / (sub ? 2 : 1), bt = row_threads(groups_x);
            for (uint32_t i0 = 0; i0 < pl->n_images; i0 += 65535) {
                PlanarView     Vz = V;
                const uint32_t nz = std::min<uint32_t>(65535, pl->n_images - i0);
                for (int p = 0; p < 3; ++p)
                    Vz.samples[p] = reinterpret_cast<const uint8_t *>(V.samples[p]) + (size_t) i0 * V.image_stride[p];
                const uint8_t *src = reinterpret_cast<const uint8_t *>(d_src) + (size_t) i0 * sx * sy * 3;
                const dim3     grid((groups_x + bt - 1) / bt, std::min<uint32_t>(rows_y, 4096), nz);
                if (sub) k_rgb8_to_ycc_planes<true><<<grid, bt, 0, ctx->stream>>>(src, Vz);
                else k_rgb8_to_ycc_planes<false><<<grid, bt, 0, ctx->stream>>>(src, Vz);
            }
            LAUNCH_CHECK(ctx);
            return JPEG_SM100_OK;
        }
    }
    for (int p = 0; p < V.n_planes; ++p) {
        if (V.scale_x % V.fx[p] || V.scale_y % V.fy[p]) return JPEG_SM100_ERR_UNSUPPORTED;
        const uint64_t work = (uint64_t) V.width[p] * V.height[p] * V.n_images;
        if (work == 0) continue;
        const uint32_t g = grid_for(ctx, work, 256, 8);
        if (pl->sample_bytes == 1) {
            if (src_is_rgb8) k_decompose<uint8_t, true><<<g, 256, 0, ctx->stream>>>(d_src, V, p);
            else k_decompose<uint8_t, false><<<g, 256, 0, ctx->stream>>>(d_src, V, p);
        } else {
            if (src_is_rgb8) k_decompose<uint16_t, true><<<g, 256, 0, ctx->stream>>>(d_src, V, p);
            else k_decompose<uint16_t, false><<<g, 256, 0, ctx->stream>>>(d_src, V, p);
        }
        LAUNCH_CHECK(ctx);
    }
    return JPEG_SM100_OK;
}

// fused.cu -- K1+K2 in one kernel: coefficients -> RGB8 for 4:2:0 (centred chroma) images of 8-bit samples.
//
// Replaces the chain Spectral.idct() -> Planar.interleaved(cosite: false) -> unpack(as: RGB.self) (reference
// decode.swift:4154 -> 4182-4276 -> jpeg.swift:441-453) without the sample planes in between: 128 B of coefficients per block in,
// 3 B of RGB per pixel out -- 6 B per pixel of HBM traffic for 4:2:0 instead of the 9 B of K1 followed by K2 (the planes alone
// are 2 x 12.4 MB per 4K frame).
//
// B200 design
//   * A CTA owns a strip of FX MCUs x a band of MCU rows of one image and walks down the band, one MCU row per step.  The centred
//     4:2:0 filter needs the chroma row above and below every luma row pair (decode.swift:4236-4265), so the rows of a step are
//     shifted up by one: step R emits luma rows 16R-1 .. 16R+14 from chroma rows 8R-1 .. 8R+7 and keeps the last luma and chroma
//     row of the step in a ring for the next one (17 luma rows / 9 chroma rows in shared memory).  Only the first step of a band
//     re-transforms the MCU row above it; horizontally every step transforms one extra chroma block on either side of the strip.
//   * Coefficient tiles arrive by TMA (cp.async.bulk.tensor, 4-D map {64, units_x, units_y, image}, 128B-swizzled, out-of-plane
//     blocks zero-filled) into a 2-stage mbarrier ring: the tiles of step i+2 are in flight while step i is transformed.
//   * IDCT: one thread per block exactly as in K1 (registers only, constant-bank multipliers); warps 0-2 take the 2 x 48 luma
//     blocks, warp 3 the 26 Cb blocks, warp 4 the 26 Cr blocks, so every warp reads ONE quantisation table from the constant bank.
//   * Colour: the packed 16-bit interpolation of k_ycc420_to_rgb8_v2 on rows read from shared memory, 8 x 2 pixels per task.
//   * Arithmetic is K1's and K2's, instruction for instruction: the output is bit-identical to the staged path.
#include "common.cuh"
#include "pixel_core.cuh"

namespace {

constexpr int FX = 24;               // MCUs per strip (4:2:0: 48 luma blocks, 24 + 2 chroma blocks per block row)
constexpr int FT = 160;              // threads: 3 luma warps + 1 Cb warp + 1 Cr warp
constexpr int F_YB = 2 * FX;         // luma blocks per block row of the strip
constexpr int F_CB = FX + 2;         // chroma blocks per row incl. one halo block on either side
constexpr int F_Y_BYTES = 2 * F_YB * 128;                  // 12288
constexpr int F_C_BYTES = F_CB * 128;                      // 3328
constexpr int F_C_SLOT = (F_C_BYTES + 1023) / 1024 * 1024;  // 4096 (128B-swizzled tiles start on 1024-byte boundaries)
constexpr int F_STAGE = F_Y_BYTES + 2 * F_C_SLOT;          // 20480
constexpr int F_STAGES = 2;
constexpr int F_YW = 16 * FX;        // luma ring row: 384 bytes
constexpr int F_CW = 8 * F_CB;       // chroma ring row: 208 bytes
constexpr int F_YROWS = 17, F_CROWS = 9;
constexpr int F_SMEM = F_STAGES * F_STAGE + F_YROWS * F_YW + 2 * F_CROWS * F_CW + 64 + 3 * 256 + 1024;

struct FusedParams {
    float    q[3][64];   // modulated quanta per plane, q[h*8+k]
    float    level;      // 128.5
    int32_t  W, H;       // image size in pixels
    int32_t  ux0, uy0, ux1, uy1;
    int32_t  n_seg, n_band, band_rows;  // strips per MCU row, bands per image, MCU rows per band
    uint32_t n_items;    // n_images * n_seg * n_band
    uint8_t *rgb;
};

struct Step {  // one MCU row of one work item
    uint32_t item;
    int32_t  img, seg, R, R0, R1;
    bool     valid;
};

__device__ __forceinline__ void step_first(Step &s, const FusedParams &P, uint32_t item)
{
    s.item = item;
    s.valid = item < P.n_items;
    if (!s.valid) return;
    const uint32_t per_img = (uint32_t) (P.n_seg * P.n_band);
    s.img = (int32_t) (item / per_img);
    const uint32_t rem = item - (uint32_t) s.img * per_img;
    const int32_t  band = (int32_t) (rem / (uint32_t) P.n_seg);
    s.seg = (int32_t) (rem - (uint32_t) band * (uint32_t) P.n_seg);
    s.R0 = band * P.band_rows;
    s.R1 = min(s.R0 + P.band_rows, P.uy1);
    s.R = s.R0 > 0 ? s.R0 - 1 : 0;  // a band below the first re-transforms the MCU row above it (its last rows feed the filter)
}
__device__ __forceinline__ void step_next(Step &s, const FusedParams &P, uint32_t stride)
{
    if (!s.valid) return;
    if (++s.R < s.R1) return;
    step_first(s, P, s.item + stride);
}

// the IDCT of one block (K1's transform_and_store, decode.swift:4101-4133) into a ring of rows in shared memory
// (qs: the plane's multipliers in shared memory -- one instantiation of the transform for all three planes: with one copy per
// plane reading the constant bank the kernel's code is 3 x 800 instructions and the warps of an SM sub-partition run three
// different copies of it)
__device__ __forceinline__ void idct_to_ring(const uint32_t (&w)[32], const float *qs, const float level, uint8_t *ring, const int stride,
                                             const int rows, int pos)
{
    float v[8][8];
#pragma unroll
    for (int h = 0; h < 8; ++h)
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int      z = zigzag_index(k, h);
            const uint32_t word = w[z >> 1];
            const short    s = (z & 1) ? (short) (word >> 16) : (short) (word & 0xffffu);
            v[h][k] = fmul(qs[h * 8 + k], (float) s);
        }
#pragma unroll
    for (int k = 0; k < 8; ++k) idct8_noshift(v[0][k], v[1][k], v[2][k], v[3][k], v[4][k], v[5][k], v[6][k], v[7][k]);
#pragma unroll
    for (int y = 0; y < 8; ++y) {
        idct8(v[y][0], v[y][1], v[y][2], v[y][3], v[y][4], v[y][5], v[y][6], v[y][7], level);
        uint2 o;
        o.x = pack4_u8_trunc(v[y][0], v[y][1], v[y][2], v[y][3]);
        o.y = pack4_u8_trunc(v[y][4], v[y][5], v[y][6], v[y][7]);
        *reinterpret_cast<uint2 *>(ring + pos * stride) = o;
        if (++pos == rows) pos = 0;
    }
}

__device__ __forceinline__ void tma_load_4d(void *smem_dst, const CUtensorMap *tmap, int c0, int c1, int c2, int c3, uint64_t *bar)
{
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                 ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar))
                 : "memory");
}

__global__ void __launch_bounds__(FT, 4)
k_idct_rgb420(const __grid_constant__ CUtensorMap tm_y, const __grid_constant__ CUtensorMap tm_cb, const __grid_constant__ CUtensorMap tm_cr,
              const __grid_constant__ FusedParams P)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t  *tiles = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t  *ys = tiles + F_STAGES * F_STAGE;    // [17][384]
    uint8_t  *cs = ys + F_YROWS * F_YW;           // [2][9][208]
    uint64_t *full = reinterpret_cast<uint64_t *>(cs + 2 * F_CROWS * F_CW);
    float    *sq = reinterpret_cast<float *>(full + 8);  // [3][64]
    const int tid = (int) threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 192; i += FT) sq[i] = P.q[i >> 6][i & 63];

    auto issue = [&](const Step &s, int stage) {  // thread 0: the three tiles of one step
        uint8_t *dst = tiles + stage * F_STAGE;
        mbar_arrive_expect_tx(&full[stage], F_Y_BYTES + 2 * F_C_BYTES);
        tma_load_4d(dst, &tm_y, 0, s.seg * F_YB, 2 * s.R, s.img, &full[stage]);
        tma_load_4d(dst + F_Y_BYTES, &tm_cb, 0, s.seg * FX - 1, s.R, s.img, &full[stage]);
        tma_load_4d(dst + F_Y_BYTES + F_C_SLOT, &tm_cr, 0, s.seg * FX - 1, s.R, s.img, &full[stage]);
    };

    Step cur, ahead;  // the step being processed; the next step to request from the TMA (two steps on)
    step_first(cur, P, blockIdx.x);
    ahead = cur;
    if (tid == 0) {
        prefetch_tmap(&tm_y);
        prefetch_tmap(&tm_cb);
        prefetch_tmap(&tm_cr);
        for (int s = 0; s < F_STAGES; ++s) mbar_init(&full[s], 1);
        fence_mbar_init();
        fence_proxy_async();
    }
    __syncthreads();
    for (int s = 0; s < F_STAGES; ++s) {
        if (tid == 0 && ahead.valid) issue(ahead, s);
        step_next(ahead, P, gridDim.x);
    }

    // this thread's block of a step: luma block (by, bx) of the strip, or chroma block j (0 and F_CB - 1 are the halo blocks)
    const int  plane = warp < 3 ? 0 : warp - 2;
    const float *qs = sq + 64 * plane;
    const int  yb_row = tid / F_YB, yb_col = tid - yb_row * F_YB;  // (plane 0)
    const int  cj = tid - (plane == 1 ? 96 : 128);                 // (planes 1, 2)
    const int  tile_off = plane == 0 ? tid * 128 : F_Y_BYTES + (plane - 1) * F_C_SLOT + cj * 128;
    const int  tile_row = plane == 0 ? tid : cj;  // row inside its 1024-aligned tile: the swizzle key
    const bool has_block = plane == 0 ? true : cj < F_CB;
    const bool vec = (P.W & 7) == 0 && (reinterpret_cast<uintptr_t>(P.rgb) & 7) == 0;
    const int  cw = 8 * P.ux1;

    for (uint32_t it = 0; cur.valid; ++it) {
        const int      stage = (int) (it & 1u);
        const uint32_t parity = (it >> 1) & 1u;
        const int      R = cur.R, seg = cur.seg;
        const bool     pre = R < cur.R0;  // transform only: this MCU row belongs to the band above
        mbar_wait(&full[stage], parity);
        uint32_t w[32];
        {
            const uint4 *row = reinterpret_cast<const uint4 *>(tiles + stage * F_STAGE + tile_off);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const uint4 c = has_block ? row[j ^ (tile_row & 7)] : make_uint4(0, 0, 0, 0);
                w[4 * j + 0] = c.x, w[4 * j + 1] = c.y, w[4 * j + 2] = c.z, w[4 * j + 3] = c.w;
            }
        }
        __syncthreads();  // stage drained (and the previous step's colour pass is done with the rings): refill it
        if (tid == 0 && ahead.valid) issue(ahead, stage);
        step_next(ahead, P, gridDim.x);

        // ---- K1: blocks of MCU row R -> the rings.  Ring slot of luma row y: (y + 1) % 17, of chroma row r: (r + 1) % 9
        const int rm17 = R % F_YROWS, rm9 = R % F_CROWS;
        {
            uint8_t *ring;
            int      stride, rows, pos;
            bool     go;
            if (plane == 0) {
                const int bx = seg * F_YB + yb_col, by = 2 * R + yb_row;
                go = bx < P.ux0 && by < P.uy0 && !(pre && yb_row == 0);
                pos = (8 * yb_row + 1 - rm17) % F_YROWS;  // (16 R + 8 row + 1) mod 17, 16 = -1 (mod 17)
                ring = ys + 8 * yb_col, stride = F_YW, rows = F_YROWS;
            } else {
                const int cbx = seg * FX - 1 + cj;
                go = has_block && cbx >= 0 && cbx < P.ux1;
                pos = (1 - rm9) % F_CROWS;  // (8 R + 1) mod 9, 8 = -1 (mod 9)
                ring = cs + (plane - 1) * F_CROWS * F_CW + 8 * cj, stride = F_CW, rows = F_CROWS;
            }
            if (pos < 0) pos += rows;
            if (go) idct_to_ring(w, qs, P.level, ring, stride, rows, pos);
        }
        __syncthreads();

        // ---- K2: chroma row pairs r = 8R-1 .. 8R+6 (and 8R+7 under the last MCU row of the image) -> luma rows 2r+1, 2r+2
        if (!pre) {
            const int x_strip = seg * 16 * FX;
            const int n_groups = min(F_YB, (P.W - x_strip + 7) >> 3);
            const int n_q = (R + 1 == P.uy1) ? 9 : 8;
            const int ybase = (F_YROWS - rm17) % F_YROWS;  // slot of luma row 16R-1
            const int cbase = (F_CROWS - rm9) % F_CROWS;   // slot of chroma row 8R-1
            for (int task = tid; task < n_q * F_YB; task += FT) {
                const int q = task / F_YB, g = task - q * F_YB;
                if (g >= n_groups) continue;
                const int r = 8 * R - 1 + q;
                const int y0 = 2 * r + 1;
                if (y0 >= P.H) continue;
                // slots of chroma rows max(r, 0) and min(r + 1, 8 uy1 - 1) (decode.swift:4244-4250: the filter clamps at the plane edges)
                int sa = cbase + q + (r < 0 ? 1 : 0), sb = cbase + q + (q == 8 ? 0 : 1);
                sa -= sa >= F_CROWS ? F_CROWS : 0, sb -= sb >= F_CROWS ? F_CROWS : 0;
                sb -= sb >= F_CROWS ? F_CROWS : 0;
                const int c0 = seg * 8 * FX + 4 * g;  // first chroma column of the group (in the plane)
                uint32_t  Pq[2][2][4];
#pragma unroll
                for (int c = 0; c < 2; ++c)
#pragma unroll
                    for (int rr = 0; rr < 2; ++rr) {
                        const uint8_t *row = cs + (c * F_CROWS + (rr ? sb : sa)) * F_CW + 8 + 4 * g;  // ring column of c0
                        const uint32_t mid = *reinterpret_cast<const uint32_t *>(row);
                        uint32_t       left = *reinterpret_cast<const uint32_t *>(row - 4) >> 24;
                        uint32_t       right = *reinterpret_cast<const uint32_t *>(row + 4) & 0xffu;
                        if (c0 == 0) left = mid & 0xffu;            // x = 0: t clamps to 0 (decode.swift:4250)
                        if (c0 + 4 >= cw) right = mid >> 24;        // last column of the padded plane (decode.swift:4244)
                        Pq[c][rr][0] = __byte_perm(mid, 0, 0x4040) * 3u + (__byte_perm(mid, left, 0x5154) + 0x00020002u);
                        Pq[c][rr][1] = __byte_perm(mid, 0, 0x4141) * 3u + (__byte_perm(mid, 0, 0x4240) + 0x00020002u);
                        Pq[c][rr][2] = __byte_perm(mid, 0, 0x4242) * 3u + (__byte_perm(mid, 0, 0x4341) + 0x00020002u);
                        Pq[c][rr][3] = __byte_perm(mid, 0, 0x4343) * 3u + (__byte_perm(mid, right, 0x5452) + 0x00020002u);
                    }
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    const int y = y0 + k;
                    if (y < 0 || y >= P.H) continue;
                    int sy = ybase + 2 * q + k;
                    sy -= sy >= F_YROWS ? F_YROWS : 0;
                    const uint2 yy = *reinterpret_cast<const uint2 *>(ys + sy * F_YW + 8 * g);
                    float       rr_[8], gg_[8], bb_[8];
#pragma unroll
                    for (int m = 0; m < 4; ++m) {
                        const uint32_t tb = (k == 0) ? Pq[0][0][m] * 3u + Pq[0][1][m] : Pq[0][1][m] * 3u + Pq[0][0][m];
                        const uint32_t tr = (k == 0) ? Pq[1][0][m] * 3u + Pq[1][1][m] : Pq[1][1][m] * 3u + Pq[1][0][m];
                        const uint32_t vb = ((tb >> 4) & 0x00ff00ffu) ^ 0x00800080u, vr = ((tr >> 4) & 0x00ff00ffu) ^ 0x00800080u;
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const int   j = 2 * m + h;
                            const float Y = (float) byte_of(j < 4 ? yy.x : yy.y, j & 3);
                            ycc_to_rgb_fast(Y, s8_to_float(vb, 2 * h), s8_to_float(vr, 2 * h), rr_[j], gg_[j], bb_[j]);
                        }
                    }
                    uint32_t o[6];
                    o[0] = pack4_u8_trunc(rr_[0], gg_[0], bb_[0], rr_[1]);
                    o[1] = pack4_u8_trunc(gg_[1], bb_[1], rr_[2], gg_[2]);
                    o[2] = pack4_u8_trunc(bb_[2], rr_[3], gg_[3], bb_[3]);
                    o[3] = pack4_u8_trunc(rr_[4], gg_[4], bb_[4], rr_[5]);
                    o[4] = pack4_u8_trunc(gg_[5], bb_[5], rr_[6], gg_[6]);
                    o[5] = pack4_u8_trunc(bb_[6], rr_[7], gg_[7], bb_[7]);
                    const int x0 = x_strip + 8 * g;
                    uint8_t  *dst = P.rgb + (((size_t) cur.img * P.H + y) * (size_t) P.W + (size_t) x0) * 3;
                    if (vec) {
                        uint2 *d = reinterpret_cast<uint2 *>(dst);
                        d[0] = make_uint2(o[0], o[1]);
                        d[1] = make_uint2(o[2], o[3]);
                        d[2] = make_uint2(o[4], o[5]);
                    } else {
                        const int n_valid = min(8, P.W - x0);
#pragma unroll
                        for (int b = 0; b < 24; ++b)
                            if (b < n_valid * 3) dst[b] = (uint8_t) (o[b >> 2] >> (8 * (b & 3)));
                    }
                }
            }
        }
        step_next(cur, P, gridDim.x);
    }
}

typedef CUresult (*encode_tiled_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_map(jpeg_sm100_ctx *ctx, CUtensorMap *tm, const jpeg_sm100_dev_spectral *sp, int p, uint32_t box_x, uint32_t box_y)
{
    const auto      &pl = sp->plane[p];
    const cuuint64_t gdim[4] = {64, (cuuint64_t) pl.units_x, (cuuint64_t) pl.units_y, sp->n_images};
    const cuuint64_t gstride[3] = {128, (cuuint64_t) 128 * pl.units_x, (cuuint64_t) pl.image_stride * 2};
    const cuuint32_t box[4] = {64, box_x, box_y, 1};
    const cuuint32_t estride[4] = {1, 1, 1, 1};
    const CUresult   r = reinterpret_cast<encode_tiled_fn>(ctx->encode_tiled)(
        tm, CU_TENSOR_MAP_DATA_TYPE_UINT16, 4, pl.coef, gdim, gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE,
        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        ctx->last_error = "cuTensorMapEncodeTiled (fused kernel) failed (" + std::to_string((int) r) + ")";
        return JPEG_SM100_ERR_CUDA;
    }
    return JPEG_SM100_OK;
}

}  // namespace

int resolve_encode_tiled(jpeg_sm100_ctx *ctx);  // idct.cu

// Coefficients -> RGB8 in one kernel.  Returns JPEG_FUSED_NOT_APPLICABLE (> 0) when the geometry is not one the fused kernel
// takes (the caller then runs K1 followed by K2), else a status.  quanta: three 64-entry tables in zig-zag order.
int jpeg_fused_spectral_to_rgb8(jpeg_sm100_ctx *ctx, const jpeg_sm100_dev_spectral *sp, const uint16_t *quanta, uint32_t sx, uint32_t sy,
                                int cosited, uint8_t *d_rgb)
{
    constexpr int NOT_APPLICABLE = 1;
    // Opt-in (JPEG_SM100_FUSE=1, read per call).  Measured on B200, 64 x 4K 4:2:0: 1.36 ms against 1.02 ms for K1 followed by K2.
    // Both halves are bound by instruction issue, not by HBM (ncu: K1 337 M + K2 574 M warp instructions at 74 % / 83 % issue-active;
    // this kernel 965 M at 66 % with 20 resident warps per SM and two CTA barriers per step): saving 3 of 9 bytes per pixel of DRAM
    // traffic (3.2 GB measured = the algorithmic 6 B/px) buys nothing while the instruction count stays the same.
    const char *fuse_env = getenv("JPEG_SM100_FUSE");
    const bool  off = !(fuse_env && atoi(fuse_env) != 0);
    if (off || cosited || !sp || sp->n_planes != 3 || sp->n_images == 0 || sx == 0 || sy == 0) return NOT_APPLICABLE;
    const auto &Y = sp->plane[0], &Cb = sp->plane[1], &Cr = sp->plane[2];
    if (!(Y.factor_x == 2 && Y.factor_y == 2 && Cb.factor_x == 1 && Cb.factor_y == 1 && Cr.factor_x == 1 && Cr.factor_y == 1)) return NOT_APPLICABLE;
    if (Cb.units_x != Cr.units_x || Cb.units_y != Cr.units_y || Cb.units_x < 1 || Cb.units_y < 1) return NOT_APPLICABLE;
    if (!(Y.units_x == 2 * Cb.units_x || Y.units_x == 2 * Cb.units_x - 1) || !(Y.units_y == 2 * Cb.units_y || Y.units_y == 2 * Cb.units_y - 1))
        return NOT_APPLICABLE;
    // every pixel must map into its plane (what fill_view checks for K2)
    if (sx > 8u * (uint32_t) Y.units_x || sy > 8u * (uint32_t) Y.units_y || (sx + 1) / 2 > 8u * (uint32_t) Cb.units_x ||
        (sy + 1) / 2 > 8u * (uint32_t) Cb.units_y)
        return NOT_APPLICABLE;
    // the filter's clamp at the bottom edge is the plane's last row: planes taller than the picture needs keep the staged path
    if ((uint32_t) Cb.units_y != (sy + 15) / 16) return NOT_APPLICABLE;
    if (sp->n_images > 0x7fffffffu || (uint64_t) sx * sy > 0x7fffffffull) return NOT_APPLICABLE;
    for (int p = 0; p < 3; ++p)
        if ((reinterpret_cast<uintptr_t>(sp->plane[p].coef) & 15) || ((sp->plane[p].image_stride * 2) & 15) ||
            sp->plane[p].image_stride < (uint64_t) 64 * sp->plane[p].units_x * sp->plane[p].units_y)
            return NOT_APPLICABLE;
    J_TRY(resolve_encode_tiled(ctx));
    CUtensorMap tm[3];
    J_TRY(make_map(ctx, &tm[0], sp, 0, F_YB, 2));
    J_TRY(make_map(ctx, &tm[1], sp, 1, F_CB, 1));
    J_TRY(make_map(ctx, &tm[2], sp, 2, F_CB, 1));
    FusedParams P;
    for (int p = 0; p < 3; ++p) modulate_quanta(quanta + 64 * p, 0.125f, P.q[p]);
    P.level = 128.5f;
    P.W = (int32_t) sx, P.H = (int32_t) sy;
    P.ux0 = Y.units_x, P.uy0 = Y.units_y, P.ux1 = Cb.units_x, P.uy1 = Cb.units_y;
    // strips must cover the pixels, not the padded plane
    const int mcus_x = (int) ((sx + 15) / 16);
    P.n_seg = (mcus_x + FX - 1) / FX;
    const char *band_env = getenv("JPEG_SM100_FUSE_BAND");  // MCU rows per band (tuning / tests)
    const int   env_band = band_env ? atoi(band_env) : 0;
    // bands: enough work items for ~8 waves over the resident CTAs, at least 8 MCU rows each (a band re-transforms one MCU row)
    const int      mcu_rows = P.uy1;
    const uint64_t slots = (uint64_t) ctx->sm_count * 4;
    int            rows = env_band > 0 ? env_band : 16;
    if (env_band <= 0) {
        const uint64_t strips = (uint64_t) sp->n_images * P.n_seg;
        while (rows > 8 && strips * ((mcu_rows + rows - 1) / rows) < 8 * slots) rows -= 2;
    }
    P.band_rows = rows;
    P.n_band = (mcu_rows + rows - 1) / rows;
    const uint64_t items = (uint64_t) sp->n_images * P.n_seg * P.n_band;
    if (items > 0x7fffffffull) return NOT_APPLICABLE;
    P.n_items = (uint32_t) items;
    P.rgb = d_rgb;
    if (!ctx->fused_smem_set) {
        CU_TRY(ctx, cudaFuncSetAttribute(k_idct_rgb420, cudaFuncAttributeMaxDynamicSharedMemorySize, F_SMEM));
        ctx->fused_smem_set = true;
    }
    const uint32_t grid = (uint32_t) (items < slots ? items : slots);
    k_idct_rgb420<<<grid, FT, F_SMEM, ctx->stream>>>(tm[0], tm[1], tm[2], P);
    LAUNCH_CHECK(ctx);
    return JPEG_SM100_OK;
}

// huffman_decode.cu -- K3: restart-interval-parallel entropy decode, all five scan kinds.
//
// Replaces Spectral.decode(ecss:interval:scan:tables:extend:) (reference decode.swift:3476-3551) and everything
// below it: the ten per-kind decoders (decode.swift:2880-3445), composites (2773-2872), EXTEND (2742-2754), the
// two-level Huffman LUT (310-351, 1037-1265) and the 1-padded bitstream (jpeg.swift:1873-1916).
//
// B200 design
//   * The only parallel axis the format offers is the restart interval (the reference decoder itself walks them in
//     a serial loop, decode.swift:3500): one THREAD decodes one interval; a warp holds 32 intervals of one image.
//   * Entropy decoding is a serial dependency chain per thread, so the kernel is latency- not bandwidth-bound.
//     Each warp is its own CTA so that the ~n_images x n_intervals / 32 warps spread over all 148 x 4 schedulers.
//   * The sequential / DC-first decoders are written as a flat per-symbol state machine: every loop trip decodes
//     exactly one Huffman symbol whatever block the lane is in, so lanes never wait for each other at block or MCU
//     boundaries (trip count = max over lanes of the interval's symbol count, not the sum of per-block maxima).
//   * The reference's two-level (8 + 8 bit) LUT is reproduced entry for entry and staged in shared memory
//     (a few KB for real tables); oversized tables fall back to global memory.
//   * Coefficients go straight to their final zig-zag slot in the Spectral.Plane layout, 64 * (units_x * y + x) + z.
#include "common.cuh"

namespace {

constexpr int WARP = 32;
#ifndef FAST_BITS_N
#define FAST_BITS_N 9
#endif
constexpr int FAST_BITS = FAST_BITS_N;               // codes of up to FAST_BITS bits resolve with one shared-memory load
constexpr int FAST_ENTRIES = 1 << FAST_BITS;
constexpr int MAX_LUT_SMEM = 44 * 1024;  // entries staged in shared memory up to this size

// ---- host: LUT construction (decode.swift:310-351 size, 1037-1240 decoder) ----------------------------------
struct LutHeader {  // per image: 8 tables (dc0..3, ac0..3)
    int32_t  n[8];
    int32_t  zeta[8];
    uint32_t offset[8];  // entry offset into the image's entry array
    int32_t  present[8];
    uint32_t fast[8];        // entry offset of the table's FAST_ENTRIES-entry fast table
    uint32_t total_entries;  // reference (two-level) entries
    uint32_t total_all;      // reference + fast entries
    uint32_t pad[2];
};
static_assert(sizeof(LutHeader) % 16 == 0, "header must keep 16-byte alignment");

bool huff_size(const uint8_t counts[16], int &n, int &z)
{
    int interior = 1;
    for (int l = 0; l < 8; ++l) {
        if (!(interior > 0)) return false;
        interior = 2 * interior - counts[l];
    }
    n = 256 - interior;
    z = n;
    for (int i = 0; i < 8; ++i) {
        if (!(interior > 0)) return false;
        z += (int) counts[8 + i] << (7 - i);
        interior = 2 * interior - counts[8 + i];
    }
    return interior > 0;
}

// host -> device hand-over of one table set: the sized header plus the raw BITS / HUFFVAL of its eight slots
struct RawSet {
    LutHeader header;
    uint8_t   counts[8][16];
    uint8_t   values[8][256];
};

// (symbol | length << 8) of the reference LUT -> the fast decoders' entry:
//     byte0 = code length, byte1 = extra bits, byte2 = zig-zag advance (run + 1; 64 for EOB), byte3 = length + extra bits
// 0 if the sequential decoders need the careful path for it (DC magnitude category > 16, or EOBn with n > 0).
// With z the zig-zag position before the symbol (0 for DC): the coefficient lands at z + advance - 1 (an EOB lands beyond
// 63, i.e. nowhere) and the next position is z + advance.
__host__ __device__ inline uint32_t fast_entry(uint32_t ref, bool dc)
{
    const uint32_t len = ref >> 8, sym = ref & 0xffu;
    if (dc) return sym > 16u ? 0u : (len | (sym << 8) | (1u << 16) | ((len + sym) << 24));
    const uint32_t size = sym & 15u, run = sym >> 4;
    if (size == 0u && run != 0u && run != 15u) return 0u;
    const uint32_t adv = sym == 0u ? 64u : run + 1u;
    return len | (size << 8) | (adv << 16) | ((len + size) << 24);
}

// decode.swift:1037-1240 Table.Huffman.decoder(): level l (codes of l+1 bits) contributes `0x8080 >> l & 0xff` clones
// of (symbol, l+1) per leaf: 128, 64, ... 1 in the level-0 table, then 128 ... 1 again in the 256-entry sub-tables.
__global__ void __launch_bounds__(128) k_build_luts(const RawSet *__restrict__ raw, uint8_t *__restrict__ luts, size_t stride)
{
    const int        ti = blockIdx.x;
    const RawSet    &r = raw[blockIdx.y];
    uint8_t         *dst = luts + stride * blockIdx.y;
    const LutHeader &h = r.header;
    if (ti == 0)
        for (uint32_t i = threadIdx.x; i < sizeof(LutHeader) / 4; i += blockDim.x)
            reinterpret_cast<uint32_t *>(dst)[i] = reinterpret_cast<const uint32_t *>(&h)[i];
    if (!h.present[ti]) return;
    uint16_t     *entries = reinterpret_cast<uint16_t *>(dst + sizeof(LutHeader)) + h.offset[ti];
    const uint8_t *counts = r.counts[ti], *values = r.values[ti];
    // leaf j of level l starts at  sum_{l' < l} counts[l'] * clones(l')  +  j * clones(l)
    uint32_t level_start = 0, leaf_base = 0;
    for (int l = 0; l < 16; ++l) {
        const uint32_t clones = (0x8080u >> l) & 0xffu;
        for (uint32_t j = threadIdx.x; j < counts[l]; j += blockDim.x) {
            const uint16_t e = (uint16_t) (values[leaf_base + j] | ((l + 1) << 8));
            uint16_t      *o = entries + level_start + j * clones;
            for (uint32_t c = 0; c < clones; ++c) o[c] = e;
        }
        level_start += counts[l] * clones;
        leaf_base += counts[l];
    }
    // FAST_BITS-bit fast table, one 32-bit entry per prefix: byte0 = code length (0: code longer than 11 bits, not a code, or
    // a symbol the sequential decoders reject -> reference lookup), byte1 = extra bits, byte2 = zero run, byte3 = 1: EOB
    __syncthreads();
    uint32_t *fast = reinterpret_cast<uint32_t *>(reinterpret_cast<uint16_t *>(dst + sizeof(LutHeader)) + h.fast[ti]);
    const int n = h.n[ti], zeta = h.zeta[ti];
    for (uint32_t i = threadIdx.x; i < (uint32_t) FAST_ENTRIES; i += blockDim.x) {
        const uint32_t cw = i << (16 - FAST_BITS);
        const int      hi = (int) (cw >> 8);
        uint32_t       e = 0x1000u;
        if (hi < n) e = entries[hi];
        else if ((int) cw < zeta) e = entries[(int) cw - 255 * n];
        fast[i] = (e >> 8) <= (uint32_t) FAST_BITS ? fast_entry(e, ti < 4) : 0u;
    }
}

// ---- device side ---------------------------------------------------------------------------------------------
struct ScanParams {
    int32_t  kind;  // 0 sequential, 1 dc first, 2 dc refine, 3 ac first, 4 ac refine
    int32_t  band_lo, band_hi, al;
    int32_t  n_comp;
    int32_t  W, H;  // iteration grid: MCUs (interleaved) or the plane's units (single component)
    int32_t  extend;
    uint32_t n_ecs;
    uint64_t interval;  // MCUs per interval; UINT64_MAX = none
    int16_t *plane[4];
    uint64_t image_stride[4];
    int32_t  ux[4], uy[4], fx[4], fy[4];
    int32_t  dc[4], ac[4];  // LUT indices: dc slot, 4 + ac slot
    int32_t  mcu_blocks;
    uint8_t  blk_comp[12], blk_dx[12], blk_dy[12];
    const uint8_t  *ecs;
    const uint64_t *offsets;
    const uint8_t  *luts;        // per image: LutHeader + entries
    uint64_t        lut_stride;  // bytes per image (0 = shared)
    uint32_t        lut_smem;    // 1: stage entries in shared memory
    int32_t        *status;      // n_images * n_ecs
};

struct BitReader {
    const uint32_t *wp;      // next aligned word to load
    uint64_t        acc;     // MSB-aligned bit buffer
    int32_t         navail;  // bits in acc
    int64_t         loaded;  // bytes of this interval already moved into acc (may be negative before start)
    int64_t         nbytes;
    int64_t         pos;     // bits consumed
    int64_t         count;   // 8 * nbytes

    __device__ __forceinline__ void load_word()
    {
        uint32_t be = 0xffffffffu;
        if (loaded < nbytes) {
            be = __byte_perm(__ldg(wp), 0, 0x0123);
            const int64_t valid = nbytes - loaded;  // bytes of this word that belong to the interval
            if (valid < 4) be |= 0xffffffffu >> (8 * (int) valid);  // jpeg.swift:1881-1887: pad with 1-bits
        }
        wp += 1;
        loaded += 4;
        acc |= (uint64_t) be << (32 - navail);
        navail += 32;
    }
    __device__ __forceinline__ void init(const uint8_t *base, int64_t n)
    {
        nbytes = n;
        count = 8 * n;
        pos = 0;
        const int lead = (int) (reinterpret_cast<uintptr_t>(base) & 3);
        wp = reinterpret_cast<const uint32_t *>(base - lead);
        acc = 0;
        navail = 0;
        loaded = -lead;
        // first word: drop the `lead` bytes that precede the interval
        uint32_t be = 0xffffffffu;
        if (n > 0) {
            be = __byte_perm(__ldg(wp), 0, 0x0123);
            const int64_t valid = n + lead;  // bytes of the word up to the interval end
            if (valid < 4) be |= 0xffffffffu >> (8 * (int) valid);
        }
        wp += 1;
        loaded += 4;
        acc = (uint64_t) be << (32 + 8 * lead);
        navail = 32 - 8 * lead;
        if (navail <= 32) load_word();
    }
    __device__ __forceinline__ void refill()
    {
        if (navail <= 32) load_word();
    }
    __device__ __forceinline__ uint32_t peek16() const { return (uint32_t) (acc >> 48); }
    __device__ __forceinline__ void     consume(int n)
    {
        acc <<= n;
        navail -= n;
        pos += n;
    }
    // consume an arbitrary (possibly > 32) number of bits -- only reachable with corrupt DC symbols
    __device__ void consume_long(int n)
    {
        while (n > 0) {
            refill();
            const int k = n < 16 ? n : 16;
            consume(k);
            n -= k;
        }
    }
};

// decode.swift:2742-2754 EXTEND with the reference's masking shifts
__device__ __forceinline__ int extend16(int binade, uint32_t tail)
{
    const uint32_t t = tail & 0xffffu;
    const uint32_t sign = t >> ((binade - 1) & 15);
    const uint32_t high = ((0xffffu + sign) << (binade & 15)) & 0xffffu;
    const uint32_t low = (t + (sign ^ 1u)) & 0xffffu;
    return (int) (short) (high | low);
}

struct Lut {
    const uint16_t *entries;  // shared or global
    int32_t         n[8], zeta[8];
    uint32_t        offset[8];
};

// decode.swift:1246-1265 Decoder[codeword] -> symbol | length << 8
__device__ __forceinline__ uint32_t lut_lookup(const uint16_t *entries, int n, int zeta, uint32_t off, uint32_t cw)
{
    const int  i = (int) (cw >> 8);
    const bool direct = i < n;
    const bool valid = direct || ((int) cw < zeta);
    const int  idx = direct ? i : (int) cw - 255 * n;
    uint32_t   e = 0x1000u;  // (symbol 0, length 16)
    if (valid) e = entries[off + idx];
    return e;
}

#define FAIL_LANE(code)                                                                                              \
    do {                                                                                                             \
        err = (code);                                                                                                \
        goto finished;                                                                                               \
    } while (0)

// ---- flat per-symbol state machine: sequential (kind 0) and DC-first (kind 1) scans --------------------------------
template <bool LUT_SMEM>
__global__ void __launch_bounds__(WARP) k_decode_flat(const __grid_constant__ ScanParams P)
{
    extern __shared__ __align__(16) uint8_t smem[];
    const uint32_t img = blockIdx.y;
    const uint32_t e = blockIdx.x * WARP + threadIdx.x;
    const uint8_t *lut_img = P.luts + (size_t) img * P.lut_stride;
    // the table header always lives in shared memory; the entries too when they fit
    const LutHeader *hdr = reinterpret_cast<const LutHeader *>(smem);
    const uint16_t  *entries = reinterpret_cast<const uint16_t *>(lut_img + sizeof(LutHeader));
    {
        const uint32_t total = reinterpret_cast<const LutHeader *>(lut_img)->total_entries;
        const uint32_t words = (uint32_t) sizeof(LutHeader) / 4 + (LUT_SMEM ? (total + 1) / 2 : 0);
        uint32_t      *dst = reinterpret_cast<uint32_t *>(smem);
        const uint32_t *src = reinterpret_cast<const uint32_t *>(lut_img);
        for (uint32_t i = threadIdx.x; i < words; i += WARP) dst[i] = src[i];
        __syncwarp();
        if (LUT_SMEM) entries = reinterpret_cast<const uint16_t *>(smem + sizeof(LutHeader));
    }
    if (e >= P.n_ecs) return;

    int err = 0;
    // rows of this interval: decode.swift:3205-3207 / 2897-2899 (integer division, clamped; we never grow planes)
    int64_t r0, r1;
    if (P.interval == UINT64_MAX) {
        r0 = 0;
        r1 = P.H;
    } else {
        r0 = (int64_t) (((uint64_t) e * P.interval) / (uint32_t) P.W);  // interval < 2^32 (checked on the host)
        r1 = (int64_t) (((uint64_t) (e + 1) * P.interval) / (uint32_t) P.W);
        if (r0 > P.H) r0 = P.H;
        if (r1 > P.H) r1 = P.H;
    }
    BitReader br;
    {
        const uint64_t o0 = P.offsets[(size_t) img * P.n_ecs + e], o1 = P.offsets[(size_t) img * P.n_ecs + e + 1];
        br.init(P.ecs + o0, (int64_t) (o1 - o0));
    }
    const bool dc_only = P.kind == 1;
    const int  al = P.al;

    int     pred0 = 0, pred1 = 0, pred2 = 0, pred3 = 0;
    int     mx = 0, blk = 0, z = 0;
    int64_t my = r0;
    bool    row_start = true;

    // per-block state
    int      comp = 0, dci = 0, aci = 0;
    int16_t *bptr = nullptr;
    bool     bvalid = false;

    if (my >= r1) goto finished;
    while (true) {
        if (z == 0) {
            if (row_start && blk == 0 && mx == 0) {
                row_start = false;
                if (P.extend) {  // decode.swift:3214-3220: stop silently at the end of the data
                    br.refill();
                    if (!(br.pos < br.count) || br.peek16() == 0xffffu) goto finished;
                }
            }
            // locate the block
            comp = P.blk_comp[blk];
            const int bx = mx * P.fx[comp] + P.blk_dx[blk];
            const int by = (int) my * P.fy[comp] + P.blk_dy[blk];
            bvalid = P.plane[comp] != nullptr && bx < P.ux[comp] && by < P.uy[comp];
            bptr = P.plane[comp] + (size_t) img * P.image_stride[comp] + 64 * ((size_t) P.ux[comp] * by + bx);
            dci = P.dc[comp];
            aci = P.ac[comp];
        }
        br.refill();
        if (!(br.pos < br.count)) FAIL_LANE(JPEG_SM100_ERR_TRUNCATED_ECS);
        const int      ti = (z == 0) ? dci : aci;
        const uint32_t ent = lut_lookup(entries, hdr->n[ti], hdr->zeta[ti], hdr->offset[ti], br.peek16());
        const int      sym = (int) (ent & 0xffu);
        br.consume((int) (ent >> 8));
        bool block_done = false;
        if (z == 0) {
            // decode.swift:2788-2820 DC composite; 3248-3254 prediction
            int diff = 0;
            if (sym > 0) {
                if (!(br.pos + sym <= br.count)) FAIL_LANE(JPEG_SM100_ERR_TRUNCATED_ECS);
                const uint32_t tail = br.peek16() >> ((16 - sym) & 15);
                diff = extend16(sym, tail);
                if (sym <= 16) br.consume(sym);
                else br.consume_long(sym);
            }
            int pr = comp == 0 ? pred0 : comp == 1 ? pred1 : comp == 2 ? pred2 : pred3;
            if (P.plane[comp] != nullptr) pr = (int) (short) (pr + diff);  // wrapping Int16, &+=
            if (comp == 0) pred0 = pr;
            else if (comp == 1) pred1 = pr;
            else if (comp == 2) pred2 = pr;
            else pred3 = pr;
            if (bvalid) bptr[0] = (int16_t) ((uint32_t) pr << al);
            z = 1;
            block_done = dc_only;
        } else {
            // decode.swift:2822-2872 AC composite; 3258-3286 block loop
            const int zeroes = sym >> 4, binade = sym & 15;
            if (binade == 0) {
                if (zeroes == 0) {
                    block_done = true;  // .eob(1)
                } else if (zeroes <= 14) {
                    if (!(br.pos + zeroes <= br.count)) FAIL_LANE(JPEG_SM100_ERR_TRUNCATED_ECS);
                    FAIL_LANE(JPEG_SM100_ERR_INVALID_BLOCK_RUN);  // .eob(n > 1) in a sequential scan
                } else {
                    z += 15;  // .run(15, value: 0)
                    if (z < 64) {
                        if (bvalid) bptr[z] = 0;
                        z += 1;
                    }
                    block_done = !(z < 64);
                }
            } else {
                if (!(br.pos + binade <= br.count)) FAIL_LANE(JPEG_SM100_ERR_TRUNCATED_ECS);
                const uint32_t tail = br.peek16() >> (16 - binade);
                const int      v = extend16(binade, tail);
                br.consume(binade);
                z += zeroes;
                if (z < 64) {
                    if (bvalid) bptr[z] = (int16_t) v;
                    z += 1;
                }
                block_done = !(z < 64);
            }
        }
        if (block_done) {
            z = 0;
            if (++blk == P.mcu_blocks) {
                blk = 0;
                if (++mx == P.W) {
                    mx = 0;
                    row_start = true;
                    if (++my >= r1) break;
                }
            }
        }
    }
finished:
    if (P.status) P.status[(size_t) img * P.n_ecs + e] = err;
}

// ---- latency-optimised decoder for sequential (kind 0) and DC-first (kind 1) scans ----------------------------------
// Every restart interval already has its own thread, so the kernel's duration is ONE thread's dependent-instruction
// chain times its symbol count.  The loop is therefore organised around that chain (ncu: a lone warp per scheduler
// issues one instruction every ~5 cycles, and exposed load latency was 20 % of the first-generation kernel):
//   * FAST PHASE (all but the last few bytes of an interval): no truncation bookkeeping at all -- while the next
//     word to load lies entirely inside the interval every consumed bit is a real bit, so the reference's
//     `i < count` / `i + n <= count` guards (decode.swift:2775-2862) cannot fire;
//   * the bitstream word for the NEXT refill is loaded one refill ahead (double buffering) and lines are prefetched
//     into L1 two lines ahead, so no load latency sits on the chain;
//   * codes of <= FAST_BITS bits (all DC codes, > 99 % of AC codes) resolve with one shared-memory load from a FAST_BITS-bit
//     table derived from the reference's two-level LUT; longer or invalid codes take the reference lookup;
//   * DC and AC symbols share one straight-line path; DC predictors live in shared memory (touched once per block);
//   * the successor block's geometry is recomputed every trip off the critical path and swapped in with selects;
//   * a CAREFUL PHASE with the full bookkeeping finishes the tail of the interval and raises the errors.
struct BlkInfo {  // one per block of the MCU, read with three 16-byte shared loads
    uint32_t base_blk, ux, uy, hasplane;  // base_blk: 128-byte block index of the plane's block (0,0) of this image
    int32_t  fx, fy, dx, dy;
    uint32_t dfast, afast;                // entry offsets (uint16 units) of the fast tables
    int32_t  tabs;                        // dci | aci << 8: LUT indices for the reference lookup
    int32_t  pred;                        // byte offset of the component's predictor row in shared memory
};
static_assert(sizeof(BlkInfo) == 48, "BlkInfo is three uint4");

__global__ void __launch_bounds__(WARP) k_decode_fast(const __grid_constant__ ScanParams P, int16_t *const plane0,
                                                      const uint32_t *const only_flagged)
{
    extern __shared__ __align__(16) uint8_t smem[];
    const uint32_t   img = blockIdx.y, lane = threadIdx.x;
    const uint32_t   e = blockIdx.x * WARP + lane;
    if (only_flagged) {  // fallback pass after k_decode_par: most warps have nothing to do
        const bool mine = e < P.n_ecs && only_flagged[(size_t) img * P.n_ecs + e] != 0u;
        if (!__any_sync(0xffffffffu, mine)) return;
    }
    const uint8_t   *lut_img = P.luts + (size_t) img * P.lut_stride;
    const LutHeader *hdr = reinterpret_cast<const LutHeader *>(smem);
    BlkInfo         *s_blk = reinterpret_cast<BlkInfo *>(smem + sizeof(LutHeader));
    constexpr uint32_t PRED0 = sizeof(LutHeader) + 12 * sizeof(BlkInfo);
    int             *s_pred = reinterpret_cast<int *>(smem + PRED0);
    constexpr uint32_t PRE = PRED0 + 4 * WARP * sizeof(int);
    const uint16_t  *entries = reinterpret_cast<const uint16_t *>(smem + PRE);
    {
        const LutHeader *gh = reinterpret_cast<const LutHeader *>(lut_img);
        const uint32_t   total = gh->total_all;
        uint32_t        *dst = reinterpret_cast<uint32_t *>(smem);
        const uint32_t  *src = reinterpret_cast<const uint32_t *>(lut_img);
        for (uint32_t i = lane; i < sizeof(LutHeader) / 4; i += WARP) dst[i] = src[i];
        uint32_t *d2 = reinterpret_cast<uint32_t *>(smem + PRE);
        for (uint32_t i = lane; i < (total + 1) / 2; i += WARP) d2[i] = src[sizeof(LutHeader) / 4 + i];
        if (lane < 12) {
            const int b = lane, c = P.blk_comp[b];
            BlkInfo   bi;
            bi.hasplane = P.plane[c] != nullptr;
            bi.base_blk = bi.hasplane ? (uint32_t) ((P.plane[c] + (size_t) img * P.image_stride[c] - plane0) / 64) : 0u;
            bi.ux = (uint32_t) P.ux[c], bi.uy = (uint32_t) P.uy[c];
            bi.fx = P.fx[c], bi.fy = P.fy[c], bi.dx = P.blk_dx[b], bi.dy = P.blk_dy[b];
            bi.dfast = gh->fast[P.dc[c]], bi.afast = gh->fast[P.ac[c]];
            bi.tabs = P.dc[c] | (P.ac[c] << 8);
            bi.pred = (int) (PRED0 + (uint32_t) c * WARP * sizeof(int));
            s_blk[b] = bi;
        }
        for (int c = 0; c < 4; ++c) s_pred[c * WARP + lane] = 0;
        __syncwarp();
    }
    if (e >= P.n_ecs) return;
    if (only_flagged && only_flagged[(size_t) img * P.n_ecs + e] == 0u) return;

    int     err = 0;
    int64_t r0, r1;
    if (P.interval == UINT64_MAX) {
        r0 = 0;
        r1 = P.H;
    } else {
        r0 = (int64_t) (((uint64_t) e * P.interval) / (uint32_t) P.W);
        r1 = (int64_t) (((uint64_t) (e + 1) * P.interval) / (uint32_t) P.W);
        if (r0 > P.H) r0 = P.H;
        if (r1 > P.H) r1 = P.H;
    }
    const uint64_t o0 = P.offsets[(size_t) img * P.n_ecs + e], o1 = P.offsets[(size_t) img * P.n_ecs + e + 1];
    const uint8_t *base = P.ecs + o0;
    if ((int64_t) (o1 - o0) > 0x0fffffff) {
        if (P.status) P.status[(size_t) img * P.n_ecs + e] = JPEG_SM100_ERR_UNSUPPORTED;
        return;
    }
    const int       nbytes = (int) (o1 - o0), count = 8 * nbytes;
    const int       lead = (int) (reinterpret_cast<uintptr_t>(base) & 3);
    const uint32_t *w0 = reinterpret_cast<const uint32_t *>(base - lead);
    const uint32_t  wlim = (uint32_t) (lead + nbytes) / 4;  // words [0, wlim) need no padding (word 0's lead bytes are shifted out)
    auto padded_word = [&](uint32_t i) -> uint32_t {
        const int first = (int) i * 4 - lead;
        uint32_t  be = 0xffffffffu;
        if (first < nbytes) {
            be = __byte_perm(__ldg(w0 + i), 0, 0x0123);
            const int valid = nbytes - first;
            if (valid < 4) be |= 0xffffffffu >> (8 * valid);
        }
        return be;
    };
    uint32_t wi = 0;  // next word to hand to the bit buffer
    uint64_t acc;
    int      navail;
    {
        const uint32_t b0 = nbytes > 0 ? padded_word(0) : 0xffffffffu;
        acc = (uint64_t) b0 << (32 + 8 * lead);
        navail = 32 - 8 * lead;
        const uint32_t b1 = padded_word(1);
        acc |= (uint64_t) b1 << (32 - navail);
        navail += 32;
        wi = 2;
    }
    const bool dc_only = P.kind == 1;
    const int  al = P.al, W = P.W, nblk = P.mcu_blocks, rend = (int) r1;

    int  z = 0, blk = 0, mx = 0, my = (int) r0;
    bool fresh_row = true;
    // current block: element pointer (nullptr: store suppressed), table offsets, predictor slot (bit 0: component has a plane)
    int16_t *cur_ptr = nullptr;
    uint32_t cur_dfo = 0, cur_afo = 0;
    int      cur_tabs = 0, cur_pred = 0;

#define LOAD_BLOCK(B, X, Y, O_PTR, O_DFO, O_AFO, O_TABS, O_PRED)                                                     \
    do {                                                                                                             \
        const uint4 *q_ = reinterpret_cast<const uint4 *>(&s_blk[(B)]);                                              \
        const uint4  a_ = q_[0], g_ = q_[1], t_ = q_[2];                                                             \
        const uint32_t bx_ = (uint32_t) (X) * g_.x + g_.z, by_ = (uint32_t) (Y) * g_.y + g_.w;                       \
        const bool     in_ = (bx_ < a_.y) & (by_ < a_.z) & (a_.w != 0u);                                             \
        const uint32_t idx_ = a_.x + a_.y * by_ + bx_;                                                               \
        O_PTR = in_ ? plane0 + (size_t) idx_ * 64 : nullptr;                                                         \
        O_DFO = t_.x, O_AFO = t_.y, O_TABS = (int) t_.z;                                                             \
        O_PRED = (int) t_.w + (int) lane * 4 + (a_.w != 0u ? 1 : 0);                                                 \
    } while (0)

    if (my >= rend) goto finished;
    LOAD_BLOCK(blk, mx, my, cur_ptr, cur_dfo, cur_afo, cur_tabs, cur_pred);

    // ================================ FAST PHASE ================================
    if (wi + 1 < wlim) {
        uint32_t nextw = __ldg(w0 + wi);  // raw word `wi`, consumed by the next refill
        for (;;) {
            // ---- successor block: depends only on (blk, mx, my); overlaps the symbol chain ----
            int       nb = blk + 1;
            const int wb = nb == nblk;
            nb = wb ? 0 : nb;
            int       nx = mx + wb;
            const int wx = nx == W;
            nx = wx ? 0 : nx;
            const int ny = my + wx;
            int16_t  *n_ptr;
            uint32_t  n_dfo, n_afo;
            int       n_tabs, n_pred;
            LOAD_BLOCK(nb, nx, ny, n_ptr, n_dfo, n_afo, n_tabs, n_pred);

            // ---- refill from the word loaded one refill ago; fetch the one after it ----
            if (navail <= 32) {
                const uint32_t be = __byte_perm(nextw, 0, 0x0123);
                wi += 1;
                nextw = __ldg(w0 + wi);
                asm volatile("prefetch.global.L1 [%0];" ::"l"(w0 + wi + 64));
                acc |= (uint64_t) be << (32 - navail);
                navail += 32;
            }
            if (P.extend && fresh_row && z == 0) {  // decode.swift:3214-3220 (`pos < count` always holds in this phase)
                if ((uint32_t) (acc >> 48) == 0xffffu) goto finished;
            }
            fresh_row = false;

            // ---- one symbol ----
            const bool     isdc = z == 0;
            const uint32_t cw = (uint32_t) (acc >> 48);
            const uint32_t *tab = reinterpret_cast<const uint32_t *>(entries + (isdc ? cur_dfo : cur_afo));
            uint32_t        ent = tab[cw >> (16 - FAST_BITS)];
            if (__builtin_expect(ent == 0u, 0)) {  // long / invalid / rejected code: the reference lookup
                const int ti = isdc ? (cur_tabs & 0xff) : (cur_tabs >> 8);
                ent = fast_entry(lut_lookup(entries, hdr->n[ti], hdr->zeta[ti], hdr->offset[ti], cw), isdc);
                if (ent == 0u) break;  // corrupt DC symbol or EOBn: finished (and diagnosed) by the careful phase
            }
            const int      len = (int) (ent & 0xffu), size = (int) __byte_perm(ent, 0, 0x4441);
            const int      adv = (int) __byte_perm(ent, 0, 0x4442), total = (int) (ent >> 24);
            const uint32_t after = (uint32_t) ((acc << len) >> 32);
            const uint32_t tail = size ? after >> (32 - size) : 0u;
            acc <<= total;
            navail -= total;
            const int v = size ? extend16(size, tail) : 0;
            // decode.swift:3248-3254: wrapping Int16 prediction; the predictor lives in shared memory, one slot per lane
            int *pslot = reinterpret_cast<int *>(smem + (cur_pred & ~3));
            const int pr0 = *pslot;
            const int pr1 = (isdc & (cur_pred & 1)) ? (int) (short) (pr0 + v) : pr0;
            *pslot = pr1;
            const int outv = isdc ? (int) ((uint32_t) pr1 << al) : v;
            const int zpos = z + adv - 1;  // z == 0 for a DC symbol (advance 1); an EOB lands beyond 63
            if ((cur_ptr != nullptr) & (zpos < 64)) cur_ptr[zpos] = (int16_t) outv;
            z = (isdc & dc_only) ? 64 : z + adv;

            // ---- block finished: swap in the successor ----
            const bool done = z >= 64;
            if (done & (ny >= rend)) goto finished;
            fresh_row = done & (nb == 0) & (nx == 0);
            z = done ? 0 : z;
            blk = done ? nb : blk, mx = done ? nx : mx, my = done ? ny : my;
            cur_ptr = done ? n_ptr : cur_ptr;
            cur_dfo = done ? n_dfo : cur_dfo, cur_afo = done ? n_afo : cur_afo;
            cur_tabs = done ? n_tabs : cur_tabs, cur_pred = done ? n_pred : cur_pred;
            if (!(wi + 1 < wlim)) break;
        }
    }

    // ================================ CAREFUL PHASE ================================
    {
        // bits handed to `acc` so far are all real: words [0, wi) minus the lead bytes
        int pos = 8 * ((int) wi * 4 - lead) - navail;
        for (;;) {
            if (navail <= 32) {
                const uint32_t be = padded_word(wi);
                wi += 1;
                acc |= (uint64_t) be << (32 - navail);
                navail += 32;
            }
            if (P.extend && fresh_row && z == 0) {
                if (!(pos < count) || (uint32_t) (acc >> 48) == 0xffffu) goto finished;
            }
            fresh_row = false;
            const bool     isdc = z == 0;
            const int      ti = isdc ? (cur_tabs & 0xff) : (cur_tabs >> 8);
            const uint32_t ent = lut_lookup(entries, hdr->n[ti], hdr->zeta[ti], hdr->offset[ti], (uint32_t) (acc >> 48));
            const int      len = (int) (ent >> 8), sym = (int) (ent & 0xffu);
            const int      size = isdc ? sym : (sym & 15);
            const int      run = isdc ? 0 : (sym >> 4);
            if (!(pos < count)) FAIL_LANE(JPEG_SM100_ERR_TRUNCATED_ECS);
            int v = 0;
            if (size > 16) {  // corrupt DC symbol: the reference's masked shifts (decode.swift:2742-2754, 2808-2818)
                acc <<= len;
                navail -= len;
                pos += len;
                if (!(pos + size <= count)) FAIL_LANE(JPEG_SM100_ERR_TRUNCATED_ECS);
                v = extend16(size, (uint32_t) (acc >> 48) >> ((16 - size) & 15));
                int left = size;
                while (left > 0) {
                    if (navail <= 32) {
                        const uint32_t be = padded_word(wi);
                        wi += 1;
                        acc |= (uint64_t) be << (32 - navail);
                        navail += 32;
                    }
                    const int k = left < 16 ? left : 16;
                    acc <<= k;
                    navail -= k;
                    pos += k;
                    left -= k;
                }
            } else {
                const bool eobn = !isdc && size == 0 && run != 0 && run != 15;
                const int  need = eobn ? run : size;
                if (need > 0 && !(pos + len + need <= count)) FAIL_LANE(JPEG_SM100_ERR_TRUNCATED_ECS);
                if (eobn) FAIL_LANE(JPEG_SM100_ERR_INVALID_BLOCK_RUN);
                const uint32_t after = (uint32_t) ((acc << len) >> 32);
                const uint32_t tail = size ? after >> (32 - size) : 0u;
                const int      total = len + size;
                acc <<= total;
                navail -= total;
                pos += total;
                v = size ? extend16(size, tail) : 0;
            }
            int outv = v;
            if (isdc) {
                int *pslot = reinterpret_cast<int *>(smem + (cur_pred & ~3));
                int  pr = *pslot;
                if (cur_pred & 1) pr = (int) (short) (pr + v);
                *pslot = pr;
                outv = (int) ((uint32_t) pr << al);
            }
            const bool eob = !isdc && sym == 0;
            const int  zpos = isdc ? 0 : z + run;
            if (cur_ptr != nullptr && !eob && zpos < 64) cur_ptr[zpos] = (int16_t) outv;
            z = eob ? 64 : zpos + 1;
            if (isdc && dc_only) z = 64;
            if (z >= 64) {
                int nb = blk + 1, nx = mx, ny = my;
                if (nb == nblk) {
                    nb = 0;
                    nx = mx + 1;
                }
                if (nx == W) {
                    nx = 0;
                    ny = my + 1;
                }
                if (ny >= rend) goto finished;
                fresh_row = (nb == 0) & (nx == 0);
                z = 0;
                blk = nb, mx = nx, my = ny;
                LOAD_BLOCK(blk, mx, my, cur_ptr, cur_dfo, cur_afo, cur_tabs, cur_pred);
            }
        }
    }
#undef LOAD_BLOCK
finished:
    if (P.status) P.status[(size_t) img * P.n_ecs + e] = err;
}

// ---- subsequence-parallel decoder for sequential scans (kind 0): intra-interval parallelism ------------------------------
// The restart interval is the only parallel axis the FORMAT gives, and a lone thread per interval is bound by its own
// dependency chain.  Huffman streams self-synchronise, so an interval can also be decoded speculatively in pieces
// (Klein & Wiseman; Weissenberger & Schmidt for JPEG on GPUs):
//   * one CTA per (image, interval); the interval's bits are cut into S <= 128 subsequences, one thread each;
//   * the parse state at a symbol boundary is (bit position, zig-zag position z, block-in-MCU b) -- MCU position and DC
//     predictors do not influence parsing;
//   * round 0: every thread parses its subsequence from a guessed state (its first bit, z = 0, b = 0) and records the state
//     it leaves with; round r: a thread whose predecessor's exit changed re-parses from that exit.  Thread 0 starts from the
//     true state, so after round r the first r + 1 exits are exact, and because streams re-synchronise after a few hundred
//     symbols almost every exit is already exact after one or two rounds.  No change anywhere => all entries exact;
//   * an exclusive scan of the per-subsequence block counts gives every subsequence its first block, then ONE decoding pass
//     writes the coefficients (AC values straight to their zig-zag slots, DC *differences* to a side array);
//   * k_dc_resolve turns the DC differences into predictions with a wrapping 16-bit prefix sum per (interval, component);
//   * anything irregular on the TRUE path (truncation, a symbol the sequential decoder rejects, a block count that does not
//     match) flags the interval; flagged intervals are zeroed and re-decoded by k_decode_fast, which also produces the
//     reference's error codes.  The speculative rounds never raise errors: garbage parses just end early.
constexpr int PAR_THREADS = 128;
constexpr int PAR_MIN_BITS = 1024;
#ifndef PAR_WARMUP_N
#define PAR_WARMUP_N 2
#endif
constexpr int PAR_WARMUP = PAR_WARMUP_N;  // subsequences of speculative warm-up in round 0

struct ParseState {
    uint32_t p;      // bit position of the next symbol
    uint16_t z, b;   // zig-zag position (0 = DC next), block index within the MCU
};
__device__ __forceinline__ uint64_t pack_state(uint32_t p, int z, int b) { return (uint64_t) p | ((uint64_t) z << 32) | ((uint64_t) b << 40); }

struct ParReader {
    const uint32_t *w0;
    uint32_t        wlim;   // words [0, wlim) need no padding
    int             lead, nbytes;
    uint64_t        acc;
    int             navail;
    uint32_t        wi;
    __device__ __forceinline__ uint32_t word(uint32_t i) const
    {
        if (i < wlim) return __byte_perm(__ldg(w0 + i), 0, 0x0123);
        const int first = (int) i * 4 - lead;
        uint32_t  be = 0xffffffffu;
        if (first < nbytes) {
            be = __byte_perm(__ldg(w0 + i), 0, 0x0123);
            const int valid = nbytes - first;
            if (valid < 4) be |= 0xffffffffu >> (8 * valid);
        }
        return be;
    }
    __device__ __forceinline__ void seek(uint32_t p)
    {
        const uint32_t ab = (uint32_t) lead * 8u + p;
        wi = ab >> 5;
        const int sh = (int) (ab & 31u);
        acc = (((uint64_t) word(wi) << 32) | word(wi + 1)) << sh;
        navail = 64 - sh;
        wi += 2;
    }
};

// Parses (FINAL = false) or decodes (FINAL = true) symbols from `st` until the bit position reaches `end_bit`.
// Returns the number of completed blocks; `st` is the exit state.  `bad` is set when the TRUE decoder would not simply
// carry on (truncation / rejected symbol); speculative callers ignore it.
// SAFE: the run (plus look-ahead) stays inside the words that need no padding, so every consumed bit is a real bit and the
// reference's truncation guards cannot fire -- no padding logic, no bit-count checks.  The loop is written branch-free
// (selects + predicated memory operations): lanes of a warp sit in different places of their blocks all the time.
template <bool FINAL, bool SAFE>
__device__ __forceinline__ uint32_t par_run(ParReader &rd, ParseState &st, const uint32_t end_bit, const uint32_t count_bits,
                                            const uint16_t *entries, const LutHeader *hdr, const BlkInfo *s_blk, const int nblk,
                                            bool &bad,
                                            // FINAL only:
                                            uint32_t N, const uint32_t N_total, const int W, const int my0, int16_t *plane0,
                                            int16_t *dcdiff)
{
    uint32_t p = st.p, done = 0;
    int      z = st.z, b = st.b;
    bad = false;
    if (p >= end_bit) return 0;
    if (FINAL && N >= N_total) return 0;
    rd.seek(p);
    uint64_t acc = rd.acc;
    int      navail = rd.navail;
    uint32_t wi = rd.wi;
    // FINAL: position and destination of the current block
    int      mx = 0, my = 0;
    int16_t *bptr = nullptr;
    if (FINAL) {
        const uint32_t mcu = N / (uint32_t) nblk;
        my = my0 + (int) (mcu / (uint32_t) W);
        mx = (int) (mcu - (mcu / (uint32_t) W) * (uint32_t) W);
        const BlkInfo &bi = s_blk[b];
        const uint32_t bx = (uint32_t) mx * bi.fx + bi.dx, by = (uint32_t) my * bi.fy + bi.dy;
        bptr = ((bx < bi.ux) & (by < bi.uy) & (bi.hasplane != 0u)) ? plane0 + (size_t) (bi.base_blk + bi.ux * by + bx) * 64 : nullptr;
    }
    while (p < end_bit) {
        if (navail <= 32) {
            const uint32_t be = SAFE ? __byte_perm(__ldg(rd.w0 + wi), 0, 0x0123) : rd.word(wi);
            acc |= (uint64_t) be << (32 - navail);
            wi += 1;
            navail += 32;
        }
        const uint4     t3 = reinterpret_cast<const uint4 *>(&s_blk[b])[2];  // dfast, afast, tabs, -
        const bool      isdc = z == 0;
        const uint32_t  cw = (uint32_t) (acc >> 48);
        const uint32_t *tab = reinterpret_cast<const uint32_t *>(entries + (isdc ? t3.x : t3.y));
        uint32_t        ent = tab[cw >> (16 - FAST_BITS)];
        if (__builtin_expect(ent == 0u, 0)) {
            const int ti = isdc ? ((int) t3.z & 0xff) : ((int) t3.z >> 8);
            ent = fast_entry(lut_lookup(entries, hdr->n[ti], hdr->zeta[ti], hdr->offset[ti], cw), isdc);
            if (ent == 0u) {
                bad = true;
                break;
            }
        }
        const int total = (int) (ent >> 24), adv = (int) __byte_perm(ent, 0, 0x4442);
        if (!SAFE) {
            if (__builtin_expect(p + (uint32_t) total > count_bits, 0)) {  // decode.swift:2808-2811, 2859-2863
                bad = true;
                break;
            }
        }
        if (FINAL) {
            const int      len = (int) (ent & 0xffu), size = (int) __byte_perm(ent, 0, 0x4441);
            const uint32_t after = (uint32_t) ((acc << len) >> 32);
            const uint32_t tail = size ? after >> (32 - size) : 0u;
            const int      v = size ? extend16(size, tail) : 0;
            const int      zpos = z + adv - 1;
            // DC differences go to the side array (resolved by k_dc_resolve); AC values to their zig-zag slot
            int16_t *dst = isdc ? dcdiff + N : bptr + zpos;
            if (isdc | ((bptr != nullptr) & (zpos < 64))) *dst = (int16_t) v;
        }
        acc <<= total;
        navail -= total;
        p += (uint32_t) total;
        z += adv;
        const bool fin = z >= 64;  // block complete
        z = fin ? 0 : z;
        done += fin ? 1u : 0u;
        const int b1 = (b + 1 == nblk) ? 0 : b + 1;
        if (FINAL) {
            if (fin) {
                N += 1;
                if (N >= N_total) {
                    b = b1;
                    break;
                }
                if (b1 == 0) {
                    mx += 1;
                    if (mx == W) {
                        mx = 0;
                        my += 1;
                    }
                }
                const BlkInfo &bi = s_blk[b1];
                const uint32_t bx = (uint32_t) mx * bi.fx + bi.dx, by = (uint32_t) my * bi.fy + bi.dy;
                bptr = ((bx < bi.ux) & (by < bi.uy) & (bi.hasplane != 0u)) ? plane0 + (size_t) (bi.base_blk + bi.ux * by + bx) * 64 : nullptr;
            }
        }
        b = fin ? b1 : b;
    }
    st.p = p;
    st.z = (uint16_t) z;
    st.b = (uint16_t) b;
    return done;
}

// picks the SAFE variant when the run, its overshoot (< 32 bits) and the 64-bit look-ahead stay in unpadded words
template <bool FINAL>
__device__ __forceinline__ uint32_t par_run_auto(ParReader &rd, ParseState &st, const uint32_t end_bit, const uint32_t count_bits,
                                                 const uint16_t *entries, const LutHeader *hdr, const BlkInfo *s_blk, const int nblk,
                                                 bool &bad, uint32_t N, const uint32_t N_total, const int W, const int my0,
                                                 int16_t *plane0, int16_t *dcdiff)
{
    // warp-uniform choice: a warp whose lanes disagree would execute both variants one after the other
    const uint32_t last_word = ((uint32_t) rd.lead * 8u + end_bit + 32u + 64u) / 32u + 1u;
    if (__all_sync(__activemask(), last_word < rd.wlim))
        return par_run<FINAL, true>(rd, st, end_bit, count_bits, entries, hdr, s_blk, nblk, bad, N, N_total, W, my0, plane0, dcdiff);
    return par_run<FINAL, false>(rd, st, end_bit, count_bits, entries, hdr, s_blk, nblk, bad, N, N_total, W, my0, plane0, dcdiff);
}

#ifndef PAR_MIN_CTAS
#define PAR_MIN_CTAS 1
#endif
__global__ void __launch_bounds__(PAR_THREADS, PAR_MIN_CTAS)
k_decode_par(const __grid_constant__ ScanParams P, int16_t *const plane0, int16_t *const dcdiff_all, const uint32_t dc_per_interval,
             uint32_t *const flagged, uint32_t *const stats)
{
    extern __shared__ __align__(16) uint8_t smem[];
    __shared__ uint64_t s_exit[PAR_THREADS], s_entry[PAR_THREADS];
    __shared__ uint32_t s_cnt[PAR_THREADS];
    __shared__ uint8_t  s_work[PAR_THREADS];
    __shared__ uint32_t s_nwork;
    __shared__ uint32_t s_warp[PAR_THREADS / 32];
    __shared__ uint32_t s_total, s_bad;
    const uint32_t   img = blockIdx.y, e = blockIdx.x, tid = threadIdx.x;
    const uint8_t   *lut_img = P.luts + (size_t) img * P.lut_stride;
    const LutHeader *hdr = reinterpret_cast<const LutHeader *>(smem);
    BlkInfo         *s_blk = reinterpret_cast<BlkInfo *>(smem + sizeof(LutHeader));
    constexpr uint32_t PRE = sizeof(LutHeader) + 12 * sizeof(BlkInfo);
    const uint16_t  *entries = reinterpret_cast<const uint16_t *>(smem + PRE);
    {
        const LutHeader *gh = reinterpret_cast<const LutHeader *>(lut_img);
        const uint32_t   total = gh->total_all;
        uint32_t        *dst = reinterpret_cast<uint32_t *>(smem);
        const uint32_t  *src = reinterpret_cast<const uint32_t *>(lut_img);
        for (uint32_t i = tid; i < sizeof(LutHeader) / 4; i += PAR_THREADS) dst[i] = src[i];
        uint32_t *d2 = reinterpret_cast<uint32_t *>(smem + PRE);
        for (uint32_t i = tid; i < (total + 1) / 2; i += PAR_THREADS) d2[i] = src[sizeof(LutHeader) / 4 + i];
        if (tid < 12) {
            const int b = tid, c = P.blk_comp[b];
            BlkInfo   bi;
            bi.hasplane = P.plane[c] != nullptr;
            bi.base_blk = bi.hasplane ? (uint32_t) ((P.plane[c] + (size_t) img * P.image_stride[c] - plane0) / 64) : 0u;
            bi.ux = (uint32_t) P.ux[c], bi.uy = (uint32_t) P.uy[c];
            bi.fx = P.fx[c], bi.fy = P.fy[c], bi.dx = P.blk_dx[b], bi.dy = P.blk_dy[b];
            bi.dfast = gh->fast[P.dc[c]], bi.afast = gh->fast[P.ac[c]];
            bi.tabs = P.dc[c] | (P.ac[c] << 8);
            bi.pred = 0;
            s_blk[b] = bi;
        }
        if (tid == 0) s_total = 0, s_bad = 0;
    }
    __syncthreads();

    int64_t r0, r1;
    if (P.interval == UINT64_MAX) {
        r0 = 0;
        r1 = P.H;
    } else {
        r0 = (int64_t) (((uint64_t) e * P.interval) / (uint32_t) P.W);
        r1 = (int64_t) (((uint64_t) (e + 1) * P.interval) / (uint32_t) P.W);
        if (r0 > P.H) r0 = P.H;
        if (r1 > P.H) r1 = P.H;
    }
    const int      W = P.W, nblk = P.mcu_blocks;
    const uint32_t N_total = (uint32_t) (r1 - r0) * (uint32_t) W * (uint32_t) nblk;
    const size_t   slot = (size_t) img * P.n_ecs + e;
    if (N_total == 0) {  // nothing to decode: the reference's row loop does not run
        if (tid == 0) {
            flagged[slot] = 0;
            if (P.status) P.status[slot] = 0;
        }
        return;
    }
    const uint64_t o0 = P.offsets[slot], o1 = P.offsets[slot + 1];
    const uint8_t *base = P.ecs + o0;
    const bool     oversize = (o1 - o0) > 0x07ffffffull || N_total > dc_per_interval;
    if (oversize) {  // 32-bit bit positions / side-array capacity: leave it to the sequential kernel
        if (tid == 0) flagged[slot] = 1;
        return;
    }
    ParReader rd;
    rd.nbytes = (int) (o1 - o0);
    rd.lead = (int) (reinterpret_cast<uintptr_t>(base) & 3);
    rd.w0 = reinterpret_cast<const uint32_t *>(base - rd.lead);
    rd.wlim = (uint32_t) (rd.lead + rd.nbytes) / 4;
    const uint32_t count = 8u * (uint32_t) rd.nbytes;
    uint32_t       B = (count + PAR_THREADS - 1) / PAR_THREADS;
    B = (B + 31u) & ~31u;
    if (B < (uint32_t) PAR_MIN_BITS) B = PAR_MIN_BITS;
    const uint32_t S = count ? (count + B - 1) / B : 1u;  // <= PAR_THREADS
    const bool     active = tid < S;
    const uint32_t start_bit = tid * B, end_bit = (tid + 1 == S) ? count : (tid + 1) * B;
    int16_t *const dcdiff = dcdiff_all + slot * dc_per_interval;

    // ---- round 0: warm up over the preceding PAR_WARMUP subsequences from a guessed state, then parse the own one ----
    // (by the time the speculative parse reaches its own first bit it has usually re-synchronised with the true parse, so
    //  most subsequences never need a second look; thread 0 -- and every thread whose warm-up starts at bit 0 -- is exact)
    ParseState st;
    bool       bad;
    if (active) {
        const uint32_t warm = tid >= (uint32_t) PAR_WARMUP ? (tid - PAR_WARMUP) * B : 0u;
        st.p = warm, st.z = 0, st.b = 0;
        if (tid > 0) par_run_auto<false>(rd, st, start_bit, count, entries, hdr, s_blk, nblk, bad, 0, 0, W, 0, nullptr, nullptr);
        s_entry[tid] = pack_state(st.p, st.z, st.b);
        s_cnt[tid] = par_run_auto<false>(rd, st, end_bit, count, entries, hdr, s_blk, nblk, bad, 0, 0, W, 0, nullptr, nullptr);
        s_exit[tid] = pack_state(st.p, st.z, st.b);
    } else {
        s_entry[tid] = 0, s_exit[tid] = 0, s_cnt[tid] = 0;
    }
    __syncthreads();
    // ---- synchronisation rounds: re-parse wherever the entry that was used differs from the predecessor's exit.
    // The subsequences to redo are compacted into a work list so that they occupy the lanes of as few warps as possible.
    uint32_t n_redo = 0, n_rounds = 0;
    for (uint32_t round = 1; round <= S + 1; ++round) {
        const bool redo = active && tid >= 1 && s_exit[tid - 1] != s_entry[tid];
        n_redo += redo ? 1u : 0u;
        n_rounds = round;
        if (tid == 0) s_nwork = 0;
        __syncthreads();
        if (redo) s_work[atomicAdd(&s_nwork, 1u)] = (uint8_t) tid;
        __syncthreads();
        const uint32_t nwork = s_nwork;
        if (nwork == 0) break;
        if (tid < nwork) {
            const uint32_t sid = s_work[tid];
            const uint64_t entry = s_exit[sid - 1];
            const uint32_t e_bit = (sid + 1 == S) ? count : (sid + 1) * B;
            st.p = (uint32_t) entry, st.z = (uint16_t) ((entry >> 32) & 0xff), st.b = (uint16_t) ((entry >> 40) & 0xff);
            const uint32_t c = par_run_auto<false>(rd, st, e_bit, count, entries, hdr, s_blk, nblk, bad, 0, 0, W, 0, nullptr, nullptr);
            const uint64_t x = pack_state(st.p, st.z, st.b);
            // every work item reads exit[sid - 1] before any item writes exit[sid]: the write is deferred past a barrier
            s_cnt[sid] = c;
            s_entry[sid] = entry;
            st.p = (uint32_t) x, st.z = (uint16_t) ((x >> 32) & 0xff), st.b = (uint16_t) ((x >> 40) & 0xff);
        }
        __syncthreads();
        if (tid < nwork) s_exit[s_work[tid]] = pack_state(st.p, st.z, st.b);
        __syncthreads();
    }
    const uint32_t my_cnt = s_cnt[tid];
    const uint64_t my_entry = s_entry[tid];
    // ---- first block of every subsequence: exclusive scan of the block counts -----------------------------------------
    uint32_t incl = my_cnt;
    const int lane = tid & 31, wid = tid >> 5;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += y;
    }
    if (lane == 31) s_warp[wid] = incl;
    __syncthreads();
    uint32_t before = incl - my_cnt;
    for (int w = 0; w < wid; ++w) before += s_warp[w];
    // ---- the one real decoding pass ---------------------------------------------------------------------------------------
    if (active) {
        st.p = (uint32_t) my_entry, st.z = (uint16_t) ((my_entry >> 32) & 0xff), st.b = (uint16_t) ((my_entry >> 40) & 0xff);
        uint32_t done = 0;
        bad = false;
        if (before < N_total)
            done = par_run_auto<true>(rd, st, end_bit, count, entries, hdr, s_blk, nblk, bad, before, N_total, W, (int) r0, plane0, dcdiff);
        if (bad) atomicOr(&s_bad, 1u);
        atomicAdd(&s_total, done);
    }
    __syncthreads();
    if (tid == 0) {
        // every expected block must have been completed (a short stream is a truncation in the reference)
        const uint32_t f = (s_bad != 0u || s_total != N_total) ? 1u : 0u;
        flagged[slot] = f;
        if (!f && P.status) P.status[slot] = 0;
    }
    if (stats) {  // JPEG_SM100_PAR_STATS=1: rounds and re-parses per interval
        atomicAdd(&stats[0], n_redo);
        if (tid == 0) {
            atomicAdd(&stats[1], n_rounds);
            atomicAdd(&stats[2], S);
            atomicAdd(&stats[3], 1u);
            atomicMax(&stats[4], n_rounds);
        }
    }
}

// DC differences -> DC coefficients: decode.swift:3248-3254 (wrapping Int16 prediction, reset per interval), one warp per
// (interval, component).  Out-of-plane blocks take part in the prediction but are not stored (decode.swift:1470-1475).
__global__ void __launch_bounds__(WARP)
k_dc_resolve(const __grid_constant__ ScanParams P, const int16_t *const dcdiff_all, const uint32_t dc_per_interval,
             const uint32_t *const flagged)
{
    const uint32_t e = blockIdx.x, img = blockIdx.y, c = blockIdx.z, lane = threadIdx.x;
    const size_t   slot = (size_t) img * P.n_ecs + e;
    if (flagged[slot]) return;  // re-decoded sequentially, DC included
    int64_t r0, r1;
    if (P.interval == UINT64_MAX) {
        r0 = 0;
        r1 = P.H;
    } else {
        r0 = (int64_t) (((uint64_t) e * P.interval) / (uint32_t) P.W);
        r1 = (int64_t) (((uint64_t) (e + 1) * P.interval) / (uint32_t) P.W);
        if (r0 > P.H) r0 = P.H;
        if (r1 > P.H) r1 = P.H;
    }
    const uint32_t W = (uint32_t) P.W, nblk = (uint32_t) P.mcu_blocks;
    const uint32_t nc = (uint32_t) (P.fx[c] * P.fy[c]);
    uint32_t       fb = 0;
    for (uint32_t b = 0; b < nblk; ++b)
        if (P.blk_comp[b] == c) {
            fb = b;
            break;
        }
    const uint32_t K = (uint32_t) (r1 - r0) * W * nc;
    const int16_t *dc = dcdiff_all + slot * dc_per_interval;
    int16_t       *pl = P.plane[c] ? P.plane[c] + (size_t) img * P.image_stride[c] : nullptr;
    int            carry = 0;
    for (uint32_t base = 0; base < K; base += WARP) {
        const uint32_t k = base + lane;
        const uint32_t mcu = k / nc, j = k - mcu * nc;
        int            v = k < K ? (int) dc[mcu * nblk + fb + j] : 0;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, v, d);
            if (lane >= (uint32_t) d) v += y;
        }
        const int pred = (int) (short) (carry + v);
        if (k < K && pl) {
            const uint32_t my = (uint32_t) r0 + mcu / W, mx = mcu - (mcu / W) * W;
            const uint32_t bx = mx * P.fx[c] + P.blk_dx[fb + j], by = my * P.fy[c] + P.blk_dy[fb + j];
            if (bx < (uint32_t) P.ux[c] && by < (uint32_t) P.uy[c])
                pl[64 * ((size_t) P.ux[c] * by + bx)] = (int16_t) ((uint32_t) pred << P.al);
        }
        carry = (int) (short) __shfl_sync(0xffffffffu, pred, 31);
    }
}

// zero the blocks of flagged intervals before the sequential kernel re-decodes them
__global__ void __launch_bounds__(128) k_zero_flagged(const __grid_constant__ ScanParams P, const uint32_t *const flagged)
{
    const uint32_t e = blockIdx.x, img = blockIdx.y;
    if (!flagged[(size_t) img * P.n_ecs + e]) return;
    int64_t r0, r1;
    if (P.interval == UINT64_MAX) {
        r0 = 0;
        r1 = P.H;
    } else {
        r0 = (int64_t) (((uint64_t) e * P.interval) / (uint32_t) P.W);
        r1 = (int64_t) (((uint64_t) (e + 1) * P.interval) / (uint32_t) P.W);
        if (r0 > P.H) r0 = P.H;
        if (r1 > P.H) r1 = P.H;
    }
    for (int c = 0; c < P.n_comp; ++c) {
        if (!P.plane[c]) continue;
        int16_t  *pl = P.plane[c] + (size_t) img * P.image_stride[c];
        const int y0 = min((int) r0 * P.fy[c], P.uy[c]), y1 = min((int) r1 * P.fy[c], P.uy[c]);
        uint4    *q = reinterpret_cast<uint4 *>(pl + 64 * (size_t) P.ux[c] * y0);
        const size_t n16 = (size_t) 8 * P.ux[c] * (y1 - y0);
        for (size_t i = threadIdx.x; i < n16; i += blockDim.x) q[i] = make_uint4(0, 0, 0, 0);
    }
}

// ---- straightforward per-thread decoders for the refinement / AC progressive scans ---------------------------------
__device__ __forceinline__ int16_t coef_get(const int16_t *pl, int ux, int uy, int x, int y, int z)
{
    if (!(x >= 0 && x < ux && y >= 0 && y < uy)) return 0;  // decode.swift:1459-1464
    return pl[64 * ((size_t) ux * y + x) + z];
}
__device__ __forceinline__ void coef_set(int16_t *pl, int ux, int uy, int x, int y, int z, int16_t v)
{
    if (!(x >= 0 && x < ux && y >= 0 && y < uy)) return;  // decode.swift:1470-1475
    pl[64 * ((size_t) ux * y + x) + z] = v;
}

template <bool LUT_SMEM>
__global__ void __launch_bounds__(WARP) k_decode_progressive(const __grid_constant__ ScanParams P)
{
    extern __shared__ __align__(16) uint8_t smem[];
    const uint32_t img = blockIdx.y;
    const uint32_t e = blockIdx.x * WARP + threadIdx.x;
    const uint8_t *lut_img = P.luts + (size_t) img * P.lut_stride;
    // the table header always lives in shared memory; the entries too when they fit
    const LutHeader *hdr = reinterpret_cast<const LutHeader *>(smem);
    const uint16_t  *entries = reinterpret_cast<const uint16_t *>(lut_img + sizeof(LutHeader));
    {
        const uint32_t total = reinterpret_cast<const LutHeader *>(lut_img)->total_entries;
        const uint32_t words = (uint32_t) sizeof(LutHeader) / 4 + (LUT_SMEM ? (total + 1) / 2 : 0);
        uint32_t      *dst = reinterpret_cast<uint32_t *>(smem);
        const uint32_t *src = reinterpret_cast<const uint32_t *>(lut_img);
        for (uint32_t i = threadIdx.x; i < words; i += WARP) dst[i] = src[i];
        __syncwarp();
        if (LUT_SMEM) entries = reinterpret_cast<const uint16_t *>(smem + sizeof(LutHeader));
    }
    if (e >= P.n_ecs) return;
    int err = 0;

    // rows: lo / w ..< min(hi / w, limit), iterated through General.Range2 (common.swift:383-409): an empty y range
    // still yields ONE row, lower > upper traps in the reference (decode.swift:3009-3013, 3031-3036, 3421-3425)
    int64_t r0, r1;
    if (P.interval == UINT64_MAX) {
        r0 = 0;
        r1 = P.H;
    } else {
        r0 = (int64_t) (((uint64_t) e * P.interval) / (uint32_t) P.W);
        r1 = (int64_t) (((uint64_t) (e + 1) * P.interval) / (uint32_t) P.W);
        if (r1 > P.H) r1 = P.H;
    }
    int64_t nrows = r1 - r0;
    BitReader br;
    const int al = P.al;
    if (r0 > r1) FAIL_LANE(JPEG_SM100_ERR_PRECONDITION);
    if (nrows == 0) nrows = 1;
    {
        const uint64_t o0 = P.offsets[(size_t) img * P.n_ecs + e], o1 = P.offsets[(size_t) img * P.n_ecs + e + 1];
        br.init(P.ecs + o0, (int64_t) (o1 - o0));
    }
    if (P.kind == 2) {
        // decode.swift:3007-3018, 3395-3445 DC refinement: one bit per block
        for (int64_t my = r0; my < r0 + nrows; ++my)
            for (int mx = 0; mx < P.W; ++mx)
                for (int b = 0; b < P.mcu_blocks; ++b) {
                    const int c = P.blk_comp[b];
                    br.refill();
                    if (!(br.pos < br.count)) FAIL_LANE(JPEG_SM100_ERR_TRUNCATED_ECS);
                    const int bit = (int) (br.peek16() >> 15);
                    br.consume(1);
                    if (P.plane[c] == nullptr) continue;
                    int16_t  *pl = P.plane[c] + (size_t) img * P.image_stride[c];
                    const int x = mx * P.fx[c] + P.blk_dx[b], y = (int) my * P.fy[c] + P.blk_dy[b];
                    coef_set(pl, P.ux[c], P.uy[c], x, y, 0,
                             (int16_t) (coef_get(pl, P.ux[c], P.uy[c], x, y, 0) | (int16_t) ((uint32_t) bit << al)));
                }
    } else {
        int16_t  *pl = P.plane[0] + (size_t) img * P.image_stride[0];
        const int ux = P.ux[0], uy = P.uy[0];
        const int ti = P.ac[0];
        const int tn = hdr->n[ti], tz = hdr->zeta[ti];
        const uint32_t toff = hdr->offset[ti];
        int       skip = 0;
        for (int64_t y64 = r0; y64 < r0 + nrows; ++y64)
            for (int x = 0; x < P.W; ++x) {
                const int y = (int) y64;
                int       z = P.band_lo;
                while (z < P.band_hi) {
                    int     zeroes = 0, kind, run = 0, v = 0;
                    int16_t delta = 0;
                    if (P.kind == 3) {
                        // decode.swift:3038-3067 AC first scan
                        if (skip != 0) {
                            skip -= 1;
                            break;
                        }
                    }
                    if (P.kind == 4 && skip > 0) {
                        zeroes = 64;
                        delta = 0;
                        skip -= 1;
                        kind = 2;
                    } else {
                        // decode.swift:2822-2872 AC composite
                        br.refill();
                        if (!(br.pos < br.count)) FAIL_LANE(JPEG_SM100_ERR_TRUNCATED_ECS);
                        const uint32_t ent = lut_lookup(entries, tn, tz, toff, br.peek16());
                        const int      sym = (int) (ent & 0xffu);
                        br.consume((int) (ent >> 8));
                        const int sz = sym >> 4, binade = sym & 15;
                        if (binade == 0) {
                            if (sz == 0) {
                                kind = 1;
                                run = 1;
                            } else if (sz <= 14) {
                                br.refill();
                                if (!(br.pos + sz <= br.count)) FAIL_LANE(JPEG_SM100_ERR_TRUNCATED_ECS);
                                kind = 1;
                                run = (1 << sz) | (int) (br.peek16() >> (16 - sz));
                                br.consume(sz);
                            } else {
                                kind = 0;
                                run = 15;
                                v = 0;
                            }
                        } else {
                            br.refill();
                            if (!(br.pos + binade <= br.count)) FAIL_LANE(JPEG_SM100_ERR_TRUNCATED_ECS);
                            kind = 0;
                            run = sz;
                            v = extend16(binade, br.peek16() >> (16 - binade));
                            br.consume(binade);
                        }
                    }
                    if (P.kind == 3) {
                        if (kind == 0) {
                            z += run;
                            if (!(z < P.band_hi)) break;
                            coef_set(pl, ux, uy, x, y, z, (int16_t) ((uint32_t) v << al));
                            z += 1;
                        } else {
                            skip = run - 1;
                            break;
                        }
                        continue;
                    }
                    // decode.swift:3093-3149 AC refinement
                    if (kind == 0) {
                        if (!(v >= -1 && v <= 1)) FAIL_LANE(JPEG_SM100_ERR_INVALID_COMPOSITE_VALUE);
                        zeroes = run;
                        delta = (int16_t) v;
                    } else if (kind == 1) {
                        zeroes = 64;
                        delta = 0;
                        skip = run - 1;
                    }
                    int  skipped = 0;
                    bool placed = false;
                    do {
                        const int16_t unrefined = coef_get(pl, ux, uy, x, y, z);
                        if (unrefined == 0) {
                            if (!(skipped < zeroes)) {
                                coef_set(pl, ux, uy, x, y, z, (int16_t) ((uint32_t) (int) delta << al));
                                z += 1;
                                placed = true;
                                break;
                            }
                            skipped += 1;
                        } else {
                            br.refill();
                            if (!(br.pos < br.count)) FAIL_LANE(JPEG_SM100_ERR_TRUNCATED_ECS);
                            const int bit = (int) (br.peek16() >> 15);
                            br.consume(1);
                            const int d = (unrefined < 0 ? -1 : 1) * bit;
                            coef_set(pl, ux, uy, x, y, z, (int16_t) (unrefined + (int16_t) ((uint32_t) d << al)));
                        }
                        z += 1;
                    } while (z < P.band_hi);
                    if (!placed) break;
                }
            }
    }
finished:
    if (P.status) P.status[(size_t) img * P.n_ecs + e] = err;
}

// per image: first non-zero status in interval order (the reference throws at the first failing interval)
__global__ void k_reduce_status(const int32_t *__restrict__ per_ecs, uint32_t n_ecs, int32_t *__restrict__ per_image)
{
    const uint32_t img = blockIdx.x;
    __shared__ uint32_t first;
    if (threadIdx.x == 0) first = 0xffffffffu;
    __syncthreads();
    for (uint32_t e = threadIdx.x; e < n_ecs; e += blockDim.x)
        if (per_ecs[(size_t) img * n_ecs + e] != 0) atomicMin(&first, e);
    __syncthreads();
    if (threadIdx.x == 0) per_image[img] = first == 0xffffffffu ? 0 : per_ecs[(size_t) img * n_ecs + first];
}

}  // namespace

// Layer-B implementation.  scratch slots 8 (LUTs) and 9 (per-ECS status) belong to this file.
int jpeg_huffman_decode_scan(jpeg_sm100_ctx *ctx, const jpeg_sm100_scan_desc *scan, const uint8_t *d_ecs,
                             const uint64_t *d_offsets, uint32_t n_ecs, uint64_t interval, int extend,
                             const jpeg_sm100_huff_table *tables, int tables_shared,
                             const jpeg_sm100_dev_spectral *sp, int32_t *d_status)
{
    if (!scan || !sp || !tables) return JPEG_SM100_ERR_INVALID_ARGUMENT;
    if (scan->n_comp < 1 || scan->n_comp > 4) return JPEG_SM100_ERR_INVALID_ARGUMENT;
    if (!(scan->band_lo >= 0 && scan->band_lo < scan->band_hi && scan->band_hi <= 64)) return JPEG_SM100_ERR_INVALID_ARGUMENT;
    const bool initial = scan->bit_hi < 0;
    ScanParams P;
    memset(&P, 0, sizeof P);
    if (scan->band_lo == 0 && scan->band_hi == 64) {
        if (!initial) return JPEG_SM100_ERR_PRECONDITION;  // fatalError("unreachable") decode.swift:3512
        P.kind = 0;
    } else if (scan->band_lo == 0 && scan->band_hi == 1)
        P.kind = initial ? 1 : 2;
    else
        P.kind = initial ? 3 : 4;
    if (P.kind >= 3 && scan->n_comp != 1) return JPEG_SM100_ERR_PRECONDITION;
    if (scan->bit_lo < 0 || scan->bit_lo > 15) return JPEG_SM100_ERR_INVALID_ARGUMENT;
    const uint32_t n_images = sp->n_images;
    if (n_images == 0 || n_ecs == 0) {
        if (d_status && n_images) CU_TRY(ctx, cudaMemsetAsync(d_status, 0, sizeof(int32_t) * n_images, ctx->stream));
        return JPEG_SM100_OK;
    }
    if (interval == 0 || (interval != JPEG_SM100_INTERVAL_NONE && interval > 0xffffffffull))
        return JPEG_SM100_ERR_INVALID_ARGUMENT;
    // stride(from: 0, to: .max, by: .max) yields a single element: without DRI only the first ECS is decoded
    if (interval == JPEG_SM100_INTERVAL_NONE && n_ecs > 1) return JPEG_SM100_ERR_INVALID_ARGUMENT;

    P.band_lo = scan->band_lo;
    P.band_hi = scan->band_hi;
    P.al = scan->bit_lo;
    P.n_comp = scan->n_comp;
    P.extend = (extend && P.kind <= 1) ? 1 : 0;
    P.n_ecs = n_ecs;
    P.interval = interval;
    P.ecs = d_ecs;
    P.offsets = d_offsets;
    const bool interleaved = scan->n_comp > 1;
    int        volume = 0;
    for (int c = 0; c < scan->n_comp; ++c) {
        const int p = scan->comp[c].plane;
        if (p >= (int) sp->n_planes) return JPEG_SM100_ERR_INVALID_ARGUMENT;
        if (p < 0 && !interleaved) {  // decode.swift:3169-3173: nothing to do
            if (d_status) CU_TRY(ctx, cudaMemsetAsync(d_status, 0, sizeof(int32_t) * n_images, ctx->stream));
            return JPEG_SM100_OK;
        }
        if (scan->comp[c].dc < 0 || scan->comp[c].dc > 3 || scan->comp[c].ac < 0 || scan->comp[c].ac > 3)
            return JPEG_SM100_ERR_INVALID_ARGUMENT;
        P.dc[c] = scan->comp[c].dc;
        P.ac[c] = 4 + scan->comp[c].ac;
        if (p >= 0) {
            P.plane[c] = sp->plane[p].coef;
            P.image_stride[c] = sp->plane[p].image_stride;
            P.ux[c] = sp->plane[p].units_x;
            P.uy[c] = sp->plane[p].units_y;
        }
        P.fx[c] = interleaved ? scan->comp[c].factor_x : 1;
        P.fy[c] = interleaved ? scan->comp[c].factor_y : 1;
        if (P.fx[c] < 1 || P.fy[c] < 1 || P.fx[c] > 4 || P.fy[c] > 4) return JPEG_SM100_ERR_INVALID_ARGUMENT;
        for (int dy = 0; dy < P.fy[c]; ++dy)
            for (int dx = 0; dx < P.fx[c]; ++dx) {
                if (volume >= 12) return JPEG_SM100_ERR_INVALID_ARGUMENT;
                P.blk_comp[volume] = (uint8_t) c;
                P.blk_dx[volume] = (uint8_t) dx;
                P.blk_dy[volume] = (uint8_t) dy;
                ++volume;
            }
    }
    P.mcu_blocks = volume;
    P.W = interleaved ? scan->blocks_x : P.ux[0];
    P.H = interleaved ? scan->blocks_y : P.uy[0];
    if (P.W <= 0 || P.H < 0) return JPEG_SM100_ERR_INVALID_ARGUMENT;

    // Tables: validated and sized on the host (16 additions per table), expanded into the reference's two-level LUT
    // by k_build_luts on the device.  decode.swift:2884-2895, 3186-3203: a missing table is an error of the scan.
    const uint32_t n_sets = tables_shared ? 1 : n_images;
    const bool need_dc = P.kind == 0 || P.kind == 1, need_ac = P.kind == 0 || P.kind == 3 || P.kind == 4;
    void *staging = nullptr;
    int   slot = 0;
    J_TRY(pinned_acquire(ctx, sizeof(RawSet) * (size_t) n_sets, &staging, &slot));
    RawSet *raw = reinterpret_cast<RawSet *>(staging);
    size_t  max_entries = 0;
    for (uint32_t s = 0; s < n_sets; ++s) {
        LutHeader &h = raw[s].header;
        memset(&h, 0, sizeof h);
        uint32_t total = 0;
        for (int c = 0; c < scan->n_comp; ++c) {
            const int slots[2] = {need_dc ? P.dc[c] : -1, need_ac ? P.ac[c] : -1};
            for (int k = 0; k < 2; ++k) {
                const int ti = slots[k];
                if (ti < 0 || h.present[ti]) continue;
                const jpeg_sm100_huff_table &t = tables[(size_t) s * 8 + ti];
                if (!t.present) return k == 0 ? JPEG_SM100_ERR_UNDEFINED_DC : JPEG_SM100_ERR_UNDEFINED_AC;
                int n, z, leaves = 0;
                if (!huff_size(t.counts, n, z)) return JPEG_SM100_ERR_INVALID_HUFFMAN;
                for (int l = 0; l < 16; ++l) leaves += t.counts[l];
                if (leaves > 256) return JPEG_SM100_ERR_INVALID_HUFFMAN;
                h.n[ti] = n;
                h.zeta[ti] = z + n * 255;
                h.offset[ti] = total;
                h.present[ti] = 1;
                total += (uint32_t) z;
                total = (total + 1u) & ~1u;  // keep every table 4-byte aligned
                memcpy(raw[s].counts[ti], t.counts, 16);
                memcpy(raw[s].values[ti], t.values, 256);
            }
        }
        h.total_entries = total;
        for (int ti = 0; ti < 8; ++ti)
            if (h.present[ti]) {
                h.fast[ti] = total;
                total += 2 * FAST_ENTRIES;  // FAST_ENTRIES x 32-bit entries
            }
        h.total_all = total;
        if (total > max_entries) max_entries = total;
    }
    const size_t entry_bytes = (max_entries * 2 + 15) & ~size_t(15);
    const size_t stride = sizeof(LutHeader) + entry_bytes + 16;
    void *d_raw = nullptr, *d_luts = nullptr;
    J_TRY(scratch_reserve(ctx, 7, sizeof(RawSet) * (size_t) n_sets, &d_raw));
    J_TRY(scratch_reserve(ctx, 8, stride * n_sets, &d_luts));
    CU_TRY(ctx, cudaMemcpyAsync(d_raw, raw, sizeof(RawSet) * (size_t) n_sets, cudaMemcpyHostToDevice, ctx->stream));
    J_TRY(pinned_release(ctx, slot));
    k_build_luts<<<dim3(8, n_sets), 128, 0, ctx->stream>>>(reinterpret_cast<const RawSet *>(d_raw),
                                                           reinterpret_cast<uint8_t *>(d_luts), stride);
    LAUNCH_CHECK(ctx);
    P.luts = reinterpret_cast<const uint8_t *>(d_luts);
    P.lut_stride = tables_shared ? 0 : stride;
    P.lut_smem = entry_bytes <= (size_t) MAX_LUT_SMEM ? 1 : 0;

    void *d_per_ecs = nullptr;
    J_TRY(scratch_reserve(ctx, 9, sizeof(int32_t) * (size_t) n_images * n_ecs, &d_per_ecs));
    P.status = reinterpret_cast<int32_t *>(d_per_ecs);

    const dim3   grid((n_ecs + WARP - 1) / WARP, n_images);
    const size_t smem = sizeof(LutHeader) + (P.lut_smem ? entry_bytes : 0);
    static const bool use_flat = [] {
        const char *e = getenv("JPEG_SM100_HUFF");
        return e && strcmp(e, "flat") == 0;  // the first-generation kernel, kept for A/B validation
    }();
    // the fast kernel addresses all planes as 128-byte blocks relative to the lowest plane pointer (32-bit indices)
    int16_t *plane0 = nullptr;
    bool     fast_ok = P.lut_smem != 0;
    for (int c = 0; c < scan->n_comp; ++c)
        if (P.plane[c] && (!plane0 || P.plane[c] < plane0)) plane0 = P.plane[c];
    for (int c = 0; c < scan->n_comp && plane0; ++c) {
        if (!P.plane[c]) continue;
        const uint64_t span = (uint64_t) (P.plane[c] - plane0) + P.image_stride[c] * (uint64_t) n_images;
        if (((P.plane[c] - plane0) & 63) || (P.image_stride[c] & 63) || span / 64 > 0xfffffff0ull) fast_ok = false;
    }
    if (!plane0) fast_ok = false;
    if (P.kind <= 1 && !use_flat && fast_ok) {
        const size_t smem2 = sizeof(LutHeader) + 12 * sizeof(BlkInfo) + 4 * WARP * sizeof(int) + entry_bytes;
        static const bool no_par = [] {
            const char *e = getenv("JPEG_SM100_HUFF");
            return e && strcmp(e, "seq") == 0;  // one thread per interval only (A/B validation of the parallel decoder)
        }();
        if (P.kind == 0 && !P.extend && !no_par) {
            // subsequence-parallel decode, then the sequential kernel for whatever it flagged, then the DC prefix sums
            const uint64_t rows_max = (interval == JPEG_SM100_INTERVAL_NONE) ? (uint64_t) P.H : (interval + P.W - 1) / P.W + 1;
            const uint64_t dc_per_interval = rows_max * (uint64_t) P.W * (uint64_t) volume;
            const uint64_t slots = (uint64_t) n_images * n_ecs;
            void          *d_dc = nullptr, *d_flag = nullptr;
            if (dc_per_interval <= 0x7fffffffull && slots * dc_per_interval * 2 <= (4ull << 30)) {
                J_TRY(scratch_reserve(ctx, 12, (size_t) (slots * dc_per_interval * 2 + 256), &d_dc));
                J_TRY(scratch_reserve(ctx, 13, (size_t) (slots * 4 + 256), &d_flag));
                const size_t smem_par = sizeof(LutHeader) + 12 * sizeof(BlkInfo) + entry_bytes;
                const dim3   grid_par(n_ecs, n_images);
                static const bool want_stats = getenv("JPEG_SM100_PAR_STATS") != nullptr;
                uint32_t         *d_stats = nullptr;
                if (want_stats) {
                    void *p = nullptr;
                    J_TRY(scratch_reserve(ctx, 14, 64, &p));
                    d_stats = reinterpret_cast<uint32_t *>(p);
                    CU_TRY(ctx, cudaMemsetAsync(d_stats, 0, 64, ctx->stream));
                }
                k_decode_par<<<grid_par, PAR_THREADS, smem_par, ctx->stream>>>(P, plane0, reinterpret_cast<int16_t *>(d_dc),
                                                                                (uint32_t) dc_per_interval,
                                                                                reinterpret_cast<uint32_t *>(d_flag), d_stats);
                LAUNCH_CHECK(ctx);
                if (want_stats) {
                    uint32_t h[5];
                    CU_TRY(ctx, cudaMemcpyAsync(h, d_stats, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
                    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
                    fprintf(stderr, "[k_decode_par] intervals %u, subsequences/interval %.1f, rounds avg %.2f max %u, re-parses per subsequence %.2f\n",
                            h[3], (double) h[2] / h[3], (double) h[1] / h[3], h[4], (double) h[0] / h[2]);
                }
                k_zero_flagged<<<grid_par, 128, 0, ctx->stream>>>(P, reinterpret_cast<const uint32_t *>(d_flag));
                LAUNCH_CHECK(ctx);
                k_decode_fast<<<grid, WARP, smem2, ctx->stream>>>(P, plane0, reinterpret_cast<const uint32_t *>(d_flag));
                LAUNCH_CHECK(ctx);
                k_dc_resolve<<<dim3(n_ecs, n_images, scan->n_comp), WARP, 0, ctx->stream>>>(
                    P, reinterpret_cast<const int16_t *>(d_dc), (uint32_t) dc_per_interval, reinterpret_cast<const uint32_t *>(d_flag));
            } else
                k_decode_fast<<<grid, WARP, smem2, ctx->stream>>>(P, plane0, nullptr);
        } else
            k_decode_fast<<<grid, WARP, smem2, ctx->stream>>>(P, plane0, nullptr);
    } else if (P.kind <= 1) {
        if (P.lut_smem) k_decode_flat<true><<<grid, WARP, smem, ctx->stream>>>(P);
        else k_decode_flat<false><<<grid, WARP, smem, ctx->stream>>>(P);
    } else {
        if (P.lut_smem) k_decode_progressive<true><<<grid, WARP, smem, ctx->stream>>>(P);
        else k_decode_progressive<false><<<grid, WARP, smem, ctx->stream>>>(P);
    }
    LAUNCH_CHECK(ctx);
    if (d_status) {
        k_reduce_status<<<n_images, 128, 0, ctx->stream>>>(P.status, n_ecs, d_status);
        LAUNCH_CHECK(ctx);
    }
    return JPEG_SM100_OK;
}

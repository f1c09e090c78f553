// huffman_decode.cu -- K3: restart-interval-parallel entropy decode, all five scan kinds.
//
// Replaces Spectral.decode(ecss:interval:scan:tables:extend:) (reference decode.swift:3476-3551) and everything
// below it: the ten per-kind decoders (decode.swift:2880-3445), composites (2773-2872), EXTEND (2742-2754), the
// two-level Huffman LUT (310-351, 1037-1265) and the 1-padded bitstream (jpeg.swift:1873-1916).
//
// B200 design (DESIGN.md section 3 has the measurements behind each choice)
//   * The format's own parallel axis is the restart interval (the reference decoder walks them in a serial loop,
//     decode.swift:3500).  Sequential scans and progressive AC-first scans are additionally cut INSIDE every interval:
//     Huffman streams self-synchronise, so k_decode_par parses subsequences of >= 1 Kbit speculatively, repairs the ones whose
//     entry state was wrong (checkpointed re-parses that stop when they merge with the recorded parse), scans the block counts
//     and then decodes every subsequence for real.  A CTA of 128 threads holds 1..8 intervals; an interval without DRI
//     (one segment per image: everything the reference's own encoder writes) gets a thread-block cluster whose CTAs exchange
//     neighbour state through distributed shared memory (k_decode_par_cluster).
//   * The kernels are bound by the latency of every lane's dependent chain of table look-ups, then by instruction issue -- not
//     by HBM.  So: as many resident warps as the SM holds (8 CTAs x 4 warps: 62 registers, 8-bit shared-memory LUTs with
//     sub-tables for longer codes, synchronisation records aliased onto the block buffers); branch-light flat state machines
//     (one Huffman symbol per trip whatever block the lane is in); stream words loaded one refill ahead; and a warp-synchronous
//     decoding pass in which lanes that complete a block wait for each other, do their block-end work together and have the
//     whole warp flush the finished blocks from shared memory as 128-byte lines.
//   * Anything irregular only flags its interval; flagged intervals are redone by the one-thread-per-interval kernels below
//     (k_decode_fast / k_decode_progressive), which evaluate every guard of the reference and produce its error codes.
//     Those kernels also serve DC-first, refinement and `extend` scans.
//   * The reference's two-level (8 + 8 bit) LUT is reproduced entry for entry on the GPU (k_build_luts).
//   * Coefficients land in the Spectral.Plane layout, 64 * (units_x * y + x) + z.
#include <cooperative_groups.h>
#include <time.h>

#include "common.cuh"

namespace {

constexpr int WARP = 32;
#ifndef FAST_BITS_N
#define FAST_BITS_N 8  // 8 (with PAR_MIN_CTAS 8) rather than 9: 4 KB less shared memory per CTA buys an eighth CTA per SM
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

constexpr uint32_t FAST_LINK = 0x80u;
// A codeword the sequential-scan decoders cannot take in their stride (invalid; DC category 16; EOBn): byte 0 carries this flag and
// the entry otherwise reads as a 16-bit symbol that ends its block with a zig-zag position no real symbol can reach: a SPECULATIVE
// parse just moves on (its result is garbage either way), the decoding pass finds z > 127 when the block ends and hands the interval
// to the sequential kernel, and k_decode_fast tests the flag and takes the careful path.
constexpr uint32_t FAST_SPECIAL = 0x40u;
constexpr uint32_t FAST_SPECIAL_ENTRY = FAST_SPECIAL | 16u | (200u << 16) | (16u << 24);  // 16 bits; advance 200: ends the block with z > 127
// (symbol | length << 8) of the reference LUT -> the fast decoders' entry:
//     byte0 = code length, byte1 = extra bits, byte2 = zig-zag advance (run + 1; 64 for EOB), byte3 = length + extra bits
// FAST_SPECIAL_ENTRY if the sequential decoders need the careful path for it (DC magnitude category > 15, or EOBn with n > 0).
// With z the zig-zag position before the symbol (0 for DC): the coefficient lands at z + advance - 1 (an EOB lands beyond
// 63, i.e. nowhere) and the next position is z + advance.
// A fast-table slot whose FAST_BITS-bit prefix is shared by longer codes holds a LINK instead: bit 7 set, bits 0..4 = 32 - d
// (so that (top << FAST_BITS) >> entry, a wrapping shift, is the sub-table index; its low three bits are 16 - FAST_BITS - d),
// bits 8.. = byte offset (from the table base) of a 2^d-entry sub-table (same entry format) addressed by the next d bits of the
// codeword.
__host__ __device__ inline uint32_t fast_entry(uint32_t ref, bool dc)
{
    const uint32_t len = ref >> 8, sym = ref & 0xffu;
    // category 16 takes the reference's masked-shift EXTEND (decode.swift:2742-2754), which is not the textbook one
    if (dc) return sym > 15u ? FAST_SPECIAL_ENTRY : (len | (sym << 8) | (1u << 16) | ((len + sym) << 24));
    const uint32_t size = sym & 15u, run = sym >> 4;
    if (size == 0u && run != 0u && run != 15u) return FAST_SPECIAL_ENTRY;
    const uint32_t adv = sym == 0u ? 64u : run + 1u;
    return len | (size << 8) | (adv << 16) | ((len + size) << 24);
}
// AC table of a progressive AC-first scan (decode.swift:2822-2872, 3038-3067): EOBn is a regular symbol there -- byte2 = 0x80 marks
// it, byte1 = n = the number of extra bits; the run is (1 << n) + those bits (n = 0: the plain end-of-block)
__host__ __device__ inline uint32_t fast_entry_prog(uint32_t ref)
{
    const uint32_t len = ref >> 8, sym = ref & 0xffu, size = sym & 15u, run = sym >> 4;
    if (size == 0u && run != 15u) return len | (run << 8) | (0x80u << 16) | ((len + run) << 24);
    return len | (size << 8) | ((run + 1u) << 16) | ((len + size) << 24);
}
constexpr uint32_t SUB_MAX = 1536;  // sub-table entries per Huffman table (6 KB); larger sets keep the reference lookup

// Depth (bits beyond FAST_BITS) of the longest canonical code under every FAST_BITS-bit prefix (T.81 Annex C code
// assignment == the leaf order of the reference's LUT, decode.swift:1037-1240).  Returns the number of sub-table entries.
// Host (layout) and device (construction) run this same function, so both agree on every offset.
__host__ __device__ inline uint32_t sub_depths(const uint8_t counts[16], uint8_t depth[FAST_ENTRIES])
{
    for (int i = 0; i < FAST_ENTRIES; ++i) depth[i] = 0;
    uint32_t code = 0;
    for (int l = 1; l <= 16; ++l) {
        if (l > FAST_BITS && counts[l - 1]) {
            const int      d = l - FAST_BITS;
            const uint32_t p0 = code >> d, p1 = (code + counts[l - 1] - 1u) >> d;
            for (uint32_t p = p0; p <= p1 && p < (uint32_t) FAST_ENTRIES; ++p)
                if (depth[p] < d) depth[p] = (uint8_t) d;
        }
        code = (code + counts[l - 1]) << 1;
    }
    uint32_t total = 0;
    for (int i = 0; i < FAST_ENTRIES; ++i)
        if (depth[i]) total += 1u << depth[i];
    if (total > SUB_MAX) {
        for (int i = 0; i < FAST_ENTRIES; ++i) depth[i] = 0;
        total = 0;
    }
    return total;
}

// decode.swift:1037-1240 Table.Huffman.decoder(): level l (codes of l+1 bits) contributes `0x8080 >> l & 0xff` clones
// of (symbol, l+1) per leaf: 128, 64, ... 1 in the level-0 table, then 128 ... 1 again in the 256-entry sub-tables.
__global__ void __launch_bounds__(128) k_build_luts(const RawSet *__restrict__ raw, uint8_t *__restrict__ luts, size_t stride, int prog)
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
        const uint32_t clones = (0x8080u >> l) & 0xffu, shift = 7u - ((uint32_t) l & 7u);  // clones == 1 << shift
        // one thread per ENTRY, not per leaf (a 1-bit code alone is 128 clones)
        for (uint32_t k = threadIdx.x; k < (uint32_t) counts[l] * clones; k += blockDim.x)
            entries[level_start + k] = (uint16_t) (values[leaf_base + (k >> shift)] | ((l + 1) << 8));
        level_start += counts[l] * clones;
        leaf_base += counts[l];
    }
    // FAST_BITS-bit fast table, one 32-bit entry per prefix (format: fast_entry), followed by the sub-tables of the
    // prefixes that longer codes share.  Everything is derived from the reference LUT just written, entry for entry.
    __shared__ uint8_t  s_depth[FAST_ENTRIES];
    __shared__ uint16_t s_off[FAST_ENTRIES];
    __shared__ uint32_t s_total;
    // the same layout sub_depths() computes on the host, spread over the CTA: clear, mark the (few) prefixes under long codes,
    // prefix-sum the sub-table sizes (warp 0, FAST_ENTRIES / 32 slots per lane)
    for (uint32_t i = threadIdx.x; i < (uint32_t) FAST_ENTRIES; i += blockDim.x) s_depth[i] = 0;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t code = 0;
        for (int l = 1; l <= 16; ++l) {
            if (l > FAST_BITS && counts[l - 1]) {
                const int      d = l - FAST_BITS;
                const uint32_t p0 = code >> d, p1 = (code + counts[l - 1] - 1u) >> d;
                for (uint32_t p = p0; p <= p1 && p < (uint32_t) FAST_ENTRIES; ++p)
                    if (s_depth[p] < d) s_depth[p] = (uint8_t) d;
            }
            code = (code + counts[l - 1]) << 1;
        }
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        constexpr int PER = FAST_ENTRIES / 32;
        uint32_t      sum = 0;
        for (int k = 0; k < PER; ++k) {
            const uint32_t d = s_depth[threadIdx.x * PER + k];
            sum += d ? 1u << d : 0u;
        }
        uint32_t inc = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, inc, d);
            if (threadIdx.x >= (uint32_t) d) inc += y;
        }
        if (threadIdx.x == 31) s_total = inc;
        uint32_t o = FAST_ENTRIES + inc - sum;
        for (int k = 0; k < PER; ++k) {
            const uint32_t d = s_depth[threadIdx.x * PER + k];
            s_off[threadIdx.x * PER + k] = (uint16_t) o;
            o += d ? 1u << d : 0u;
        }
    }
    __syncthreads();
    if (s_total > SUB_MAX) {  // too many sub-table entries: this table keeps the reference lookup for its long codes
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < (uint32_t) FAST_ENTRIES; i += blockDim.x) s_depth[i] = 0;
    }
    __syncthreads();
    uint32_t *fast = reinterpret_cast<uint32_t *>(reinterpret_cast<uint16_t *>(dst + sizeof(LutHeader)) + h.fast[ti]);
    const int n = h.n[ti], zeta = h.zeta[ti];
    auto ref_lookup = [&](uint32_t cw, bool &valid) -> uint32_t {
        const int hi = (int) (cw >> 8);
        valid = true;
        if (hi < n) return entries[hi];
        if ((int) cw < zeta) return entries[(int) cw - 255 * n];
        valid = false;
        return 0x1000u;
    };
    for (uint32_t i = threadIdx.x; i < (uint32_t) FAST_ENTRIES; i += blockDim.x) {
        const uint32_t cw = i << (16 - FAST_BITS);
        const uint32_t d = s_depth[i];
        if (d == 0u) {
            bool           valid;
            const uint32_t e = ref_lookup(cw, valid);
            const bool     ok = valid && (e >> 8) <= (uint32_t) FAST_BITS;
            fast[i] = (prog && ti >= 4) ? (ok ? fast_entry_prog(e) : 0u) : (ok ? fast_entry(e, ti < 4) : FAST_SPECIAL_ENTRY);
            continue;
        }
        fast[i] = FAST_LINK | (32u - d) | ((uint32_t) s_off[i] << 10);  // shift, byte offset
    }
    // the sub-tables (up to 256 entries under one prefix): a warp per prefix, its lanes over the entries
    for (uint32_t i = threadIdx.x >> 5; i < (uint32_t) FAST_ENTRIES; i += blockDim.x >> 5) {
        const uint32_t d = s_depth[i];
        if (d == 0u) continue;
        const uint32_t cw = i << (16 - FAST_BITS);
        for (uint32_t j = threadIdx.x & 31u; j < (1u << d); j += 32u) {
            bool           valid;
            const uint32_t e = ref_lookup(cw | (j << (16 - FAST_BITS - d)), valid);
            const bool     ok = valid && (e >> 8) <= (uint32_t) FAST_BITS + d;
            fast[s_off[i] + j] = (prog && ti >= 4) ? (ok ? fast_entry_prog(e) : 0u) : (ok ? fast_entry(e, ti < 4) : FAST_SPECIAL_ENTRY);
        }
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

// fast table + sub-table lookup: one shared-memory load for codes of up to FAST_BITS bits, a second (predicated) one for
// longer codes.  0: not resolvable here (invalid codeword, or a symbol the sequential decoders treat specially).
__device__ __forceinline__ uint32_t fast_lookup(const uint32_t *tab, uint32_t cw)
{
    uint32_t ent = tab[cw >> (16 - FAST_BITS)];
    if (ent & FAST_LINK) {
        const uint32_t rest = cw & ((1u << (16 - FAST_BITS)) - 1u);
        ent = tab[(ent >> 10) + (rest >> (ent & 7u))];
    }
    return ent;
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
            uint32_t        ent = fast_lookup(tab, cw);
            if (__builtin_expect((ent & FAST_SPECIAL) != 0u, 0)) {  // invalid / rejected code (or oversized table set): the reference lookup
                const int ti = isdc ? (cur_tabs & 0xff) : (cur_tabs >> 8);
                ent = fast_entry(lut_lookup(entries, hdr->n[ti], hdr->zeta[ti], hdr->offset[ti], cw), isdc);
                if (ent & FAST_SPECIAL) break;  // corrupt DC symbol or EOBn: finished (and diagnosed) by the careful phase
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
//   * a CTA holds 128 / T intervals of one image, T = 16..128 threads each (or a cluster of CTAs holds one big interval); the
//     interval's bits are cut into S <= T subsequences, one thread each;
//   * the parse state at a symbol boundary is (bit position, zig-zag position z, block-in-MCU b) -- MCU position and DC
//     predictors do not influence parsing;
//   * round 0: every thread parses its subsequence from a guessed state (its first bit, z = 0, b = 0) and records the state
//     it leaves with; round r: a thread whose predecessor's exit changed re-parses from that exit.  Thread 0 starts from the
//     true state, so after round r the first r + 1 exits are exact, and because streams re-synchronise after a few hundred
//     symbols almost every exit is already exact after one or two rounds.  No change anywhere => all entries exact;
//   * an exclusive scan of the per-subsequence block counts gives every subsequence its first block, then ONE decoding pass
//     (par_run_final) assembles every block in its owner's shared-memory buffer (DC *differences* go to a side array) and the
//     warp stores completed blocks as whole 128-byte lines;
//   * the CTA's warps then turn the DC differences into predictions with a wrapping 16-bit prefix sum per (interval, component);
//   * anything irregular on the TRUE path (truncation, a symbol the sequential decoder rejects, a block count that does not
//     match) flags the interval; flagged intervals are zeroed and re-decoded by k_decode_fast, which also produces the
//     reference's error codes.  The speculative rounds never raise errors: garbage parses just end early.
#ifndef PAR_THREADS_N
#define PAR_THREADS_N 128
#endif
constexpr int PAR_THREADS = PAR_THREADS_N;  // 128: four 32-thread intervals share a CTA (and one copy of the tables)
constexpr int PAR_BIG_THREADS = 512;  // CTA size for scans with few, large intervals (no DRI: one entropy-coded segment per image)
constexpr int PAR_MIN_BITS = 1024;
#ifndef PAR_NSEG_N
#define PAR_NSEG_N 8
#endif
constexpr int PAR_NSEG = PAR_NSEG_N;  // checkpoints per subsequence
#ifndef PAR_BUF_STRIDE_N
#define PAR_BUF_STRIDE_N 128
#endif
// bytes between the threads' block buffers.  144: 16-byte aligned, 4 banks apart.  128 (2 KB less per CTA: with 56 registers a ninth
// CTA fits on the SM): buffers are 128-byte aligned and the eight 16-byte chunks of lane L's buffer sit at chunk ^ (L & 7), so that
// lanes writing the same zig-zag slot spread over eight bank groups instead of hitting one.
constexpr uint32_t PAR_BUF_STRIDE = PAR_BUF_STRIDE_N;
constexpr bool     PAR_SWIZZLE = PAR_BUF_STRIDE == 128;
static_assert(PAR_BUF_STRIDE == 128 || PAR_BUF_STRIDE == 144, "block buffer stride");

struct ParseState {
    uint32_t p;      // bit position of the next symbol
    uint16_t z, b;   // zig-zag position (0 = DC next), block index within the MCU
};
__device__ __forceinline__ uint64_t pack_state(uint32_t p, int z, int b) { return (uint64_t) p | ((uint64_t) z << 32) | ((uint64_t) b << 40); }
__device__ __forceinline__ ParseState unpack_state(uint64_t x)
{
    ParseState s;
    s.p = (uint32_t) x, s.z = (uint16_t) ((x >> 32) & 0xff), s.b = (uint16_t) ((x >> 40) & 0xff);
    return s;
}

struct ParIO {             // where the bytes of one restart interval live
    const uint32_t *w0;    // the aligned word that holds its first byte
    uint32_t        wlim;  // words [0, wlim) need no padding
    int32_t         lead, nbytes;
    uint32_t        clamp;  // the interval ends within 16 bytes of the end of the buffer: loads are clamped to wlast
    uint32_t        wlast;  // index of the last word that holds a byte of the interval
    // jpeg.swift:1881-1887: bytes past the end of the interval read as 1-bits; `be` = word i, already big-endian
    __device__ __forceinline__ uint32_t pad(uint32_t be, uint32_t i) const
    {
        const int first = (int) i * 4 - lead;
        if (first >= nbytes) return 0xffffffffu;
        const int valid = nbytes - first;
        return valid < 4 ? be | (0xffffffffu >> (8 * valid)) : be;
    }
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
};
struct ParGroup {  // one restart interval of the CTA (the CTA decodes PAR_THREADS / T of them, T threads each)
    ParIO    io;
    uint32_t count, B, S;   // bits, bits per subsequence, subsequences (0: nothing to do in this kernel)
    uint32_t N_total;       // blocks the interval must produce
    int32_t  r0, r1;        // MCU rows
    uint32_t slot, valid;   // img * n_ecs + e; interval exists
    uint32_t total, bad;
};

// per block of the MCU: where its coefficients go and which tables decode it (two 16-byte shared loads)
struct ParBlk {
    uint32_t C, fx, R, lim;           // block index (128-byte units from plane0) = C + mx * fx + my * R; lim = LX | LY << 16:
                                      // the block lies inside its plane iff mx < LX and my < LY (0 / 0: component without plane)
    uint32_t dtab, atab, tabs, next;  // shared-memory addresses of the DC / AC fast tables; dc | ac << 8 | b << 16 | (b == 0) << 24;
                                      // shared-memory address of the successor block's second quad
};
static_assert(sizeof(ParBlk) == 32, "two uint4 per block of the MCU");
constexpr uint32_t PAR_PRE = sizeof(LutHeader) + 12 * sizeof(ParBlk);  // shared-memory offset of the fast tables in the parallel kernels

__device__ __forceinline__ uint32_t lds32(uint32_t a)
{
    uint32_t v;
    asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ uint4 lds128(uint32_t a)
{
    uint4 v;
    asm("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}

// `extend` scans (the first scan of a file, decode.swift:3214-3220, 2906-2912) test `bits[b, count: 16] != 0xffff` before every
// MCU row.  top: the next 32 stream bits as the parallel decoders hold them (NOT 1-padded past the end of the data);
// remaining: bits from the cursor to the end of the data (> 0).
__device__ __forceinline__ bool row_start_all_ones(const uint32_t top, const int remaining)
{
    uint32_t t16 = top >> 16;
    if (remaining < 16) t16 |= 0xffffu >> remaining;  // jpeg.swift:1881-1887: the stream is padded with 1-bits
    return t16 == 0xffffu;
}

// ---- generation 6: the two hot loops, rebuilt around their instruction count ---------------------------------------------------
// The kernel is bound by instruction issue (ncu, round 1: 66 % issue-active at 32 resident warps; the synchronisation parse
// cost 38 and the decoding pass ~105 warp instructions per symbol step).  Both loops now keep ONE absolute bit cursor
//     a = 8 * lead + p        (a & 31: shift of the funnel shift, a >> 5: stream word; compared against `bound`, `end`, `count`)
// instead of (cnt, left, slack); a table entry's byte 3 (code length + extra bits) is added to it with one LEA.HI; the table to
// look a symbol up in is a state variable (DC table after a block end, AC table after any symbol) instead of a select on z == 0;
// sub-table links carry their shift in the low five bits (consumed by a wrapping shift, no masking); codewords the sequential
// decoders single out are ordinary entries with FAST_SPECIAL set, so the synchronisation parse never tests for them.
// Stream words are read without padding or clamping: the bytes that follow an interval are the next interval's (readable), and a
// symbol that reaches past `count` ends the run.  Only an interval within 16 bytes of the end of the buffer clamps (CLAMP).
__device__ __forceinline__ uint32_t ldg_word(const uint32_t *p) { return __ldg(p); }
// shared-memory address of the entry a symbol's first FAST_BITS bits select (top byte of `top` times four, plus the table)
__device__ __forceinline__ uint32_t fast_slot(const uint32_t tab, const uint32_t top)
{
    uint32_t idx, addr;
    asm volatile("shr.u32 %0, %1, %2;" : "=r"(idx) : "r"(top), "n"(32 - FAST_BITS));  // (volatile: keeps SHF + LEA, not SHF + LOP3 + IADD)
    asm("mad.lo.u32 %0, %1, 4, %2;" : "=r"(addr) : "r"(idx), "r"(tab));
    return addr;
}

// ---- a lane's window on its part of the entropy-coded segment ----------------------------------------------------------------------
// The symbol loops run under predication: a lane refills its window when IT has consumed 32 bits (every ~6 symbols), but the warp
// issues the (predicated) refill instructions at every step, and an instruction that names the destination register of a global
// load in flight waits for it whether its predicate is on or not: the word loaded "one refill ahead" is really needed one STEP
// later (ncu: a quarter of all stall samples sit on its byte swap, long scoreboard).  Two remedies, both measured:
//   PAR_RING_N = 0 (default)  plain loads.  (PAR_L2_POLICY = 1 additionally loads the stream with an L2 evict-last policy and
//               stores the coefficient lines with evict-first, so that the three passes over a subsequence find it in L2: no gain.)
//   PAR_RING_N = 8, 16        the words travel global -> shared memory with cp.async (LDGSTS: no destination register, no
//               scoreboard) into a private ring per lane, and a refill takes its word from the ring.  Removes the long-scoreboard
//               stalls (26 % -> 5 % of samples) but costs 20 % more instructions and a CTA per SM: 2.27 ms against 1.96 ms.
#ifndef PAR_DEFER_STEPS
#define PAR_DEFER_STEPS 6  // measured (64 x 4K, 16 threads per interval): 4: 1.93 ms, 6: 1.89, 8: 1.94, 12: 1.99
#endif
#ifndef PAR_RING_N
#define PAR_RING_N 0
#endif
constexpr uint32_t PAR_RING = PAR_RING_N;  // words per lane (0: no ring; else a power of two >= 2 * PAR_DEFER_STEPS, see win_refill)
static_assert(PAR_RING == 0 || ((PAR_RING & (PAR_RING - 1)) == 0 && PAR_RING >= 2 * PAR_DEFER_STEPS), "ring depth");
__device__ __forceinline__ void cp_async4(const uint32_t saddr, const uint32_t *g)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(saddr), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async_commit()
{
    if (PAR_RING) asm volatile("cp.async.commit_group;" ::: "memory");
}
template <int N>
__device__ __forceinline__ void cp_async_wait()
{
    if (PAR_RING) asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
// L2 eviction policies (createpolicy): the entropy-coded bytes stay, the coefficient lines go
__device__ __forceinline__ uint64_t l2_policy_evict_last()
{
    uint64_t p;
    asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first()
{
    uint64_t p;
    asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
#ifndef PAR_DC_INLINE
#define PAR_DC_INLINE 1  // 1: DC predictions resolved BEFORE the flush (per-subsequence DC sums from the synchronisation parse, scanned
                         // with the block counts): no side array, no fix-up pass over the flushed lines -- see k_decode_par
#endif
#ifndef PAR_L2_POLICY
#define PAR_L2_POLICY 0  // 1: L2 eviction policies on the stream loads / coefficient stores (measured: no gain, 2.12 vs 2.10 ms)
#endif
__device__ __forceinline__ uint32_t ldg_stream(const uint32_t *p, const uint64_t pol)
{
    if (!PAR_L2_POLICY) return __ldg(p);
    uint32_t v;
    asm volatile("ld.global.nc.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
    return v;
}

struct Window {
    uint32_t hi, lo;  // the 64 stream bits from word (a >> 5) on, big-endian
    uint32_t nxt;     // the word after lo, still raw (swapped at the point of use)
    uint32_t bound;   // bit index of the word after lo: refill when a >= bound
    uint32_t ringp;   // ring: shared address of the slot that holds the word after nxt
    uint32_t wnext;   // index (from io.w0) of the word the next request fetches
    uint64_t pol;     // L2 policy of the stream loads
};
template <bool CLAMP>
__device__ __forceinline__ void win_open(Window &w, const ParIO &io, const uint32_t a, const uint32_t ring)
{
    const uint32_t cur = a >> 5;
    w.pol = l2_policy_evict_last();
    if (PAR_RING) {
#pragma unroll
        for (uint32_t k = 0; k < PAR_RING; ++k) cp_async4(ring + 4u * k, io.w0 + (CLAMP ? min(cur + 3u + k, io.wlast) : cur + 3u + k));
        cp_async_commit();
    }
    w.hi = __byte_perm(ldg_stream(io.w0 + (CLAMP ? min(cur, io.wlast) : cur), w.pol), 0, 0x0123);
    w.lo = __byte_perm(ldg_stream(io.w0 + (CLAMP ? min(cur + 1u, io.wlast) : cur + 1u), w.pol), 0, 0x0123);
    w.nxt = ldg_stream(io.w0 + (CLAMP ? min(cur + 2u, io.wlast) : cur + 2u), w.pol);
    w.bound = (a & ~31u) + 32u;
    w.ringp = ring;
    w.wnext = cur + 3u + PAR_RING;
    cp_async_wait<0>();
}
// Ring: one commit + wait_group 1 per PAR_DEFER_STEPS steps proves the slot about to be read complete -- a word is read PAR_RING
// refills after it was requested, a step refills at most once, so with PAR_RING >= 2 * PAR_DEFER_STEPS the request lies at least
// one whole group back.
template <bool CLAMP>
__device__ __forceinline__ void win_refill(Window &w, const ParIO &io)
{
    w.hi = w.lo;
    asm volatile("prmt.b32 %0, %1, 0, 0x0123;" : "=r"(w.lo) : "r"(w.nxt));  // the swap stays at the point of use (see k_decode_fast)
    if (PAR_RING) {
        w.nxt = lds32(w.ringp);
        cp_async4(w.ringp, io.w0 + (CLAMP ? min(w.wnext, io.wlast) : w.wnext));  // the slot just read takes the word PAR_RING further on
        w.ringp = (w.ringp & ~(4u * PAR_RING - 1u)) | ((w.ringp + 4u) & (4u * PAR_RING - 1u));
    } else
        w.nxt = ldg_stream(io.w0 + (CLAMP ? min(w.wnext, io.wlast) : w.wnext), w.pol);
    w.wnext += 1u;
    w.bound += 32u;
}

// four wrapping 16-bit sums (one per component of the scan) in two registers
__device__ __forceinline__ uint32_t add16x2(uint32_t a, uint32_t b) { return ((a & 0x7fff7fffu) + (b & 0x7fff7fffu)) ^ ((a ^ b) & 0x80008000u); }
__device__ __forceinline__ uint32_t sub16x2(uint32_t a, uint32_t b) { return add16x2(a, add16x2(~b, 0x00010001u)); }
__device__ __forceinline__ void     acc16(uint2 &s, const uint32_t comp, const uint32_t v)  // s[comp] += v (mod 2^16)
{
    const uint32_t add = (comp & 1u) ? v << 16 : v & 0xffffu;
    if (comp & 2u) s.y = add16x2(s.y, add);
    else s.x = add16x2(s.x, add);
}
__device__ __forceinline__ uint32_t get16(const uint2 s, const uint32_t comp) { return ((comp & 2u) ? s.y : s.x) >> (16u * (comp & 1u)) & 0xffffu; }

// DCS: also accumulate the DC differences of the blocks that START in [st.p, end_bit), per component, into dcs (PAR_DC_INLINE)
template <bool CLAMP, bool DCS = false>
__device__ __forceinline__ uint32_t par_parse(const ParIO &io, ParseState &st, const uint32_t end_bit, const uint32_t count_bits,
                                              const uint32_t blk0, const uint32_t ring, uint2 *dcs = nullptr)
{
    if (st.p >= end_bit) return 0;
    const uint32_t base = (uint32_t) io.lead * 8u;
    uint32_t       a = base + st.p;
    const uint32_t end_a = base + end_bit;
    int            z = st.z;
    uint32_t       done = 0;
    Window         w;
    win_open<CLAMP>(w, io, a, ring);
    // the tables of the block in progress: (DC table, AC table, selectors, link to the successor's quad), reloaded as a whole
    uint4 q = lds128(blk0 + (uint32_t) st.b * (uint32_t) sizeof(ParBlk) + 16u);
    do {
#pragma unroll
        for (int step = 0; step < PAR_DEFER_STEPS; ++step) {
            if (a < end_a) {
                if (a >= w.bound) win_refill<CLAMP>(w, io);
                const uint32_t top = __funnelshift_l(w.lo, w.hi, a);  // the next 32 bits of the stream
                const uint32_t tab = z == 0 ? q.x : q.y;
                uint32_t       ent = lds32(fast_slot(tab, top));
                if (ent & FAST_LINK) ent = lds32(tab + (ent >> 8) + ((top << FAST_BITS) >> (ent & 31u)) * 4u);
                if (DCS) {
                    if (z == 0) {  // the DC symbol of a block that starts in this subsequence: its difference, T.81 EXTEND as in par_decode
                        const uint32_t top2 = __funnelshift_l(0u, top, ent), size = ent >> 8;
                        const uint32_t tail = __funnelshift_l(top2, 0u, size), mask = __funnelshift_l(0xffffffffu, 0u, size);
                        acc16(*dcs, (q.z >> 25) & 3u, (int) top2 >= 0 ? tail - mask : tail);
                    }
                }
                // (no end-of-data test in the loop: see below)
                a += ent >> 24;
                z += (int) __byte_perm(ent, 0, 0x4442);
                if (z >= 64) {  // block complete: the successor's tables
                    z = 0;
                    done += 1;
                    q = lds128(q.w);
                }
            }
        }
        cp_async_commit();
        cp_async_wait<1>();
    } while (a < end_a);
    cp_async_wait<0>();  // (the ring is reused by the next call)
    // A symbol that reaches past the data does not exist (decode.swift:2808-2811, 2859-2863).  On a healthy stream only one such
    // "symbol" can be met: the <= 7 padding 1-bits after the interval's last block read as a DC codeword -- no real one (all-ones
    // prefixes are never complete codes), i.e. a FAST_SPECIAL entry, 16 bits long, which ends a block that is not there.
    if (a > base + count_bits && z == 0 && done) done -= 1;
    st.p = a - base;
    st.z = (uint16_t) z;
    st.b = (uint16_t) ((q.z >> 16) & 0xffu);
    return done;
}

// warp-uniform choice of the variant (a warp whose lanes disagree would execute both one after the other)
template <bool DCS = false>
__device__ __forceinline__ uint32_t par_parse_auto(const ParIO &io, ParseState &st, const uint32_t end_bit, const uint32_t count_bits,
                                                   const uint32_t blk0, const uint32_t ring, uint2 *dcs = nullptr)
{
    if (__any_sync(__activemask(), io.clamp != 0u)) return par_parse<true, DCS>(io, st, end_bit, count_bits, blk0, ring, dcs);
    return par_parse<false, DCS>(io, st, end_bit, count_bits, blk0, ring, dcs);
}

// ---- the decoding pass, warp-synchronous ------------------------------------------------------------------------------------------
// Every lane of the warp calls it (`go`: the lane has a subsequence to decode) and the lanes stay converged.
// A block is decoded by the thread in whose subsequence it STARTS: that thread runs past the end of its subsequence until the block
// is complete, and the thread that finds a block in progress at its entry skips to the block's end.  So every block is assembled
// by one thread, in that thread's 128-byte buffer in shared memory (zero-initialised, coefficient z at buf[z]; the DC DIFFERENCE
// sits in slot 0 like any other coefficient), and leaves as eight 16-byte stores: whole lines.
// Block ends are deferred: about two of 32 lanes complete a block at every step, and the ~80 instructions of block-end work
// (successor tables, geometry, DC side array, flush) at 2 lanes cost the warp more than a symbol step.  A lane that completes a
// block stops (z >= 64 marks it as waiting, `lim` = 0 keeps it from stepping) until the warp's next block-end section, which runs
// after every PAR_DEFER_STEPS symbol steps: the waiting lanes do their bookkeeping together and queue (buffer | in-plane, block
// index); the whole warp then flushes the queued blocks, 8 lanes x 16 bytes per block, 4 blocks per LDS.128 / STG.128 / STS.128.
// The symbol step is predicated on ONE compare, a < lim  (lim = `end` while the lane skips the block in progress at its entry,
// unbounded while it owns the block in progress, 0 while it waits or is finished), and contains no test that can wait for the
// block end: a FAST_SPECIAL entry advances z beyond 127 and a symbol that reaches past the data leaves a > count, both seen by the
// block-end section; the lane that skips a block stores into its own buffer like everybody else and clears it afterwards.
__device__ __forceinline__ void sts128_zero(uint32_t a)
{
    asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(a), "r"(0u) : "memory");
}
// DCI (PAR_DC_INLINE): `pred` holds, per component, the DC prediction at the lane's first own block (the sum of the differences of
// all blocks that start before its subsequence); every own block's difference is turned into its DC value before the flush.
template <bool CLAMP, bool EXT, bool DCI = false>
__device__ __forceinline__ uint32_t par_decode(const bool go, const ParIO &io, const ParseState st, const uint32_t end_bit,
                                               const uint32_t count_bits, const uint32_t blk0, const int nblk, bool &bad_out, uint32_t N,
                                               const uint32_t N_total, const int W, const int my0, int16_t *plane0, int16_t *dcdiff,
                                               const uint32_t buf /* shared address of the lane's block buffer */,
                                               const uint32_t fq /* shared address of the warp's 32-entry flush queue */,
                                               const uint32_t ring /* shared address of the lane's stream ring */, uint2 pred = make_uint2(0u, 0u))
{
    constexpr uint32_t FULL = 0xffffffffu;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t base = (uint32_t) io.lead * 8u;
    uint32_t       a = base + st.p;
    const uint32_t end_a = base + end_bit, count_a = base + count_bits;
    int            z = st.z;
    const bool     run = go && st.p < end_bit && N < N_total;
    uint32_t       done = 0, lim = 0, ent = 0, bad = 0;
    Window         w = {0, 0, 0, 0, ring, 0, 0};
    uint4          q = make_uint4(0, 0, 0, 0);
    int            mx = 0, my = 0;
    uint32_t       bidx = 0;   // destination of the block in progress, in 128-byte units from plane0
    uint32_t       owned = 0;  // the block in progress started in this subsequence (bit 0); it lies inside its plane (bit 1)
    bool           rowstart = false;  // EXT (`extend` scans): the next symbol is the DC symbol that opens an MCU row
#pragma unroll
    for (int i = 0; i < 8; ++i) sts128_zero(buf + 16u * (uint32_t) i);
    if (run) {
        win_open<CLAMP>(w, io, a, ring);
        q = lds128(blk0 + (uint32_t) st.b * (uint32_t) sizeof(ParBlk) + 16u);
        const uint32_t mcu = N / (uint32_t) nblk;
        my = my0 + (int) (mcu / (uint32_t) W);
        mx = (int) (mcu - (mcu / (uint32_t) W) * (uint32_t) W);
        const uint4 g = lds128(blk0 + (uint32_t) st.b * (uint32_t) sizeof(ParBlk));
        const bool  inp = ((uint32_t) mx < (g.w & 0xffffu)) & ((uint32_t) my < (g.w >> 16));
        bidx = g.x + (uint32_t) mx * g.y + (uint32_t) my * g.z;
        owned = (z == 0 ? 1u : 0u) | (inp ? 2u : 0u);  // entry at a block boundary: the block that starts here is this thread's
        lim = z == 0 ? 0xffffffffu : end_a;
        if (EXT) rowstart = z == 0 && st.b == 0 && mx == 0;
    }
    __syncwarp();
    const uint32_t swz = PAR_SWIZZLE ? (lane & 7u) << 4 : 0u;  // chunk c of this lane's buffer sits at chunk c ^ (lane & 7)
    const uint64_t pol_out = l2_policy_evict_first();
    const uint32_t bufm2 = buf - 2u;
    for (;;) {
#pragma unroll
        for (int step = 0; step < PAR_DEFER_STEPS; ++step) {
            if (a < lim) {
                if (a >= w.bound) win_refill<CLAMP>(w, io);
                const uint32_t top = __funnelshift_l(w.lo, w.hi, a);
                if (EXT) {
                    if (rowstart) {  // sixteen 1-bits at the start of a row end an `extend` scan silently (decode.swift:3214-3220):
                        rowstart = false;  // never on a healthy stream, so the interval just goes to the sequential decoder
                        if (row_start_all_ones(top, (int) (count_a - a))) bad = 1u, lim = 0u;
                    }
                }
                const uint32_t tab = z == 0 ? q.x : q.y;
                ent = lds32(fast_slot(tab, top));
                if (ent & FAST_LINK) ent = lds32(tab + (ent >> 8) + ((top << FAST_BITS) >> (ent & 31u)) * 4u);
                const uint32_t top2 = __funnelshift_l(0u, top, ent);            // top << code length (the low five bits of ent)
                const uint32_t size = ent >> 8;                                 // (its low five bits: the number of extra bits)
                const uint32_t tail = __funnelshift_l(top2, 0u, size);          // the extra bits: top2 >> (32 - size), 0 if none
                const uint32_t mask = __funnelshift_l(0xffffffffu, 0u, size);   // 2^size - 1
                const uint32_t v = (int) top2 >= 0 ? tail - mask : tail;        // T.81 EXTEND (first extra bit 0: negative)
                a += ent >> 24;
                z += (int) __byte_perm(ent, 0, 0x4442);
                if (z <= 64) {  // the coefficient lands at z_before + advance - 1 (an EOB lands beyond 63: nowhere)
                    const uint32_t slot = PAR_SWIZZLE ? (((2u * (uint32_t) z - 2u) ^ swz) | buf) : bufm2 + 2u * (uint32_t) z;
                    asm volatile("st.shared.u16 [%0], %1;" ::"r"(slot), "h"((uint16_t) v) : "memory");
                }
                if (z >= 64) lim = 0u;  // block complete: the lane waits for the block-end section
            }
        }
        cp_async_commit();
        cp_async_wait<1>();
        const uint32_t mw = __ballot_sync(FULL, z >= 64);
        if (mw == 0u) {
            if (!__any_sync(FULL, a < lim)) break;
            continue;
        }
        uint32_t fin_buf = 0, fin_bidx = 0;  // fin_buf != 0: this lane hands a block to the flush
        if (z >= 64) {
            // a FAST_SPECIAL entry (invalid codeword / a symbol the sequential decoders single out: z > 127) or a symbol that
            // reached past the data (decode.swift:2808-2811, 2859-2863): the interval is left to the sequential kernel
            const bool broken = z > 127 || a > count_a;
            done += (!broken && (a - (ent >> 24)) < end_a) ? 1u : 0u;  // counted where its last symbol STARTS (still in `ent`)
            const uint32_t cur = q.w, comp_done = (q.z >> 25) & 3u;
            q = lds128(cur);
            if (owned & 1u) {
                fin_buf = buf | ((owned >> 1) & (broken ? 0u : 1u)) | 2u | (swz >> 2), fin_bidx = bidx;  // (bits 2..4: lane & 7)
                uint16_t d;
                asm volatile("ld.shared.u16 %0, [%1];" : "=h"(d) : "r"(buf + swz) : "memory");
                if (DCI) {  // difference -> DC value, in the buffer, before the block leaves (decode.swift:3248-3254: wrapping Int16)
                    acc16(pred, comp_done, (uint32_t) d);
                    asm volatile("st.shared.u16 [%0], %1;" ::"r"(buf + swz), "h"((uint16_t) get16(pred, comp_done)) : "memory");
                } else
                    dcdiff[N] = (int16_t) d;  // DC differences also go to the side array (resolved into predictions after the pass)
            } else {  // the block this lane skipped: its symbols were stored like any others
#pragma unroll
                for (int i = 0; i < 8; ++i) sts128_zero(buf + 16u * (uint32_t) i);
            }
            N += 1;
            z = 0;
            bad |= broken ? 1u : 0u;
            if (N < N_total && a < end_a && bad == 0u) {  // the block that starts at `a` is this thread's
                lim = 0xffffffffu;
                mx += (int) ((q.z >> 24) & 1u);
                if (mx == W) {
                    mx = 0;
                    my += 1;
                }
                if (EXT) rowstart = ((q.z >> 24) & 1u) != 0u && mx == 0;
                const uint4 g = lds128(cur - 16u);
                const bool  inp = ((uint32_t) mx < (g.w & 0xffffu)) & ((uint32_t) my < (g.w >> 16));
                bidx = g.x + (uint32_t) mx * g.y + (uint32_t) my * g.z;
                owned = 1u | (inp ? 2u : 0u);
            }  // else: lim stays 0 -- finished
        }
        const uint32_t m = __ballot_sync(FULL, fin_buf != 0u);
        if (m) {
            if (fin_buf) {
                const uint32_t rank = __popc(m & ((1u << lane) - 1u));
                asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(fq + 8u * rank), "r"(fin_buf), "r"(fin_bidx) : "memory");
            }
            __syncwarp();
            const uint32_t n = __popc(m), chunk = (lane & 7u) * 16u;
            for (uint32_t k = lane >> 3; k < n; k += 4u) {
                uint32_t src, dst;
                asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(src), "=r"(dst) : "r"(fq + 8u * k) : "memory");
                const uint32_t sa = PAR_SWIZZLE ? (src & ~127u) + (chunk ^ ((src << 2) & 0x70u)) : (src & ~3u) + chunk;
                uint4          v;
                asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(sa) : "memory");
                if (!PAR_L2_POLICY) {
                    if (src & 1u) *reinterpret_cast<uint4 *>(reinterpret_cast<char *>(plane0) + (int64_t) (int32_t) dst * 128 + chunk) = v;
                } else if (src & 1u)  // (written once, read by the next kernel from DRAM: the lines should not push the stream out of L2)
                    asm volatile("st.global.L2::cache_hint.v4.u32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(reinterpret_cast<char *>(plane0) + (int64_t) (int32_t) dst * 128 + chunk),
                                 "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "l"(pol_out)
                                 : "memory");
                sts128_zero(sa);
            }
            __syncwarp();
        }
    }
    cp_async_wait<0>();
    bad_out = bad != 0u;
    return done;
}

__device__ __forceinline__ uint32_t par_decode_auto(const bool go, const ParIO &io, const ParseState st, const uint32_t end_bit,
                                                    const uint32_t count_bits, const uint32_t blk0, const int nblk, bool &bad,
                                                    const uint32_t N, const uint32_t N_total, const int W, const int my0, int16_t *plane0,
                                                    int16_t *dcdiff, const uint32_t buf, const uint32_t fq, const uint32_t ring, const bool ext,
                                                    const bool dci = false, const uint2 pred = make_uint2(0u, 0u))
{
    const bool clamp = __any_sync(0xffffffffu, go && io.clamp != 0u);
#if PAR_DC_INLINE
    if (dci) {  // (kernel-uniform)
        if (ext) {
            if (clamp) return par_decode<true, true, true>(go, io, st, end_bit, count_bits, blk0, nblk, bad, N, N_total, W, my0, plane0, dcdiff, buf, fq, ring, pred);
            return par_decode<false, true, true>(go, io, st, end_bit, count_bits, blk0, nblk, bad, N, N_total, W, my0, plane0, dcdiff, buf, fq, ring, pred);
        }
        if (clamp) return par_decode<true, false, true>(go, io, st, end_bit, count_bits, blk0, nblk, bad, N, N_total, W, my0, plane0, dcdiff, buf, fq, ring, pred);
        return par_decode<false, false, true>(go, io, st, end_bit, count_bits, blk0, nblk, bad, N, N_total, W, my0, plane0, dcdiff, buf, fq, ring, pred);
    }
#endif
    if (ext) {  // (kernel-uniform)
        if (clamp) return par_decode<true, true>(go, io, st, end_bit, count_bits, blk0, nblk, bad, N, N_total, W, my0, plane0, dcdiff, buf, fq, ring);
        return par_decode<false, true>(go, io, st, end_bit, count_bits, blk0, nblk, bad, N, N_total, W, my0, plane0, dcdiff, buf, fq, ring);
    }
    if (clamp) return par_decode<true, false>(go, io, st, end_bit, count_bits, blk0, nblk, bad, N, N_total, W, my0, plane0, dcdiff, buf, fq, ring);
    return par_decode<false, false>(go, io, st, end_bit, count_bits, blk0, nblk, bad, N, N_total, W, my0, plane0, dcdiff, buf, fq, ring);
}

// ---- progressive AC-first scans (kind 3) on the same machinery ------------------------------------------------------------------
// One component, one table, blocks in raster order of the plane; the parse state is (bit position, z) alone -- no place in an MCU
// to agree on, so streams re-synchronise quickly.  An EOBn symbol ends its block AND the n-1 blocks after it without consuming
// another bit, so the whole run counts as completed blocks at that symbol and never needs to be carried in the state.
// Coefficients are scattered 2-byte stores (the block's other bands belong to other scans); an invalid codeword or a read past the
// end flags the interval, which k_decode_progressive then redoes with the reference's error semantics.
template <bool FINAL, bool SAFE>
__device__ __forceinline__ uint32_t par_run_ac(const ParIO &io, ParseState &st, const uint32_t end_bit, const uint32_t count_bits,
                                               const uint32_t tab, const int band_lo, const int band_hi, const int al, bool &bad,
                                               uint32_t N, const uint32_t N_total, int16_t *base /* block 0 of the interval */)
{
    int z = st.z;
    bad = false;
    if (st.p >= end_bit) return 0;
    if (FINAL && N >= N_total) return 0;
    uint32_t  done = 0;
    int       left = (int) (end_bit - st.p);
    const int slack = (int) (count_bits - end_bit);
    uint32_t  wi, cnt, hi, lo, nxt;
    {
        const uint32_t ab = (uint32_t) io.lead * 8u + st.p;
        wi = ab >> 5;
        cnt = ab & 31u;
        hi = io.word(wi), lo = io.word(wi + 1);
        nxt = __ldg(io.w0 + (SAFE ? wi + 2 : min(wi + 2, io.wlast)));
        wi += 3;
    }
    while (left > 0) {
        if (cnt >= 32u) {
            hi = lo;
            asm volatile("prmt.b32 %0, %1, 0, 0x0123;" : "=r"(lo) : "r"(nxt));
            nxt = __ldg(io.w0 + (SAFE ? wi : min(wi, io.wlast)));
            wi += 1;
            cnt -= 32u;
        }
        const uint32_t top = __funnelshift_l(lo, hi, cnt);
        uint32_t       ent = lds32(tab + ((top >> (32 - FAST_BITS)) << 2));
        if (ent & FAST_LINK) {
            const uint32_t rest = (top >> 16) & ((1u << (16 - FAST_BITS)) - 1u);
            ent = lds32(tab + (ent >> 8) + ((rest >> (ent & 7u)) << 2));
        }
        if (__builtin_expect(ent == 0u, 0)) {  // not a codeword of this table (or a table too large for sub-tables)
            bad = true;
            break;
        }
        const int total = (int) (ent >> 24);
        if (!SAFE) {
            if (__builtin_expect(total > left + slack, 0)) {  // decode.swift:2843-2846, 2859-2863
                bad = true;
                break;
            }
        }
        const int      len = (int) (ent & 0x7fu), size = (int) __byte_perm(ent, 0, 0x4441), adv = (int) __byte_perm(ent, 0, 0x4442);
        const uint32_t top2 = top << len;
        const uint32_t tail = __funnelshift_rc(top2, 0u, 32 - size);
        cnt += (uint32_t) total;
        left -= total;
        if (adv & 0x80) {  // EOBn: this block and (1 << n) + tail - 1 more are done (decode.swift:3059-3064)
            const uint32_t nb = (1u << size) + tail;
            done += nb;
            N += nb;
            z = band_lo;
        } else {
            const int v = (int) top2 >= 0 ? (int) (tail + (0xffffffffu << size) + 1u) : (int) tail;
            const int zpos = z + adv - 1;  // decode.swift:3047-3055: skip `run` coefficients, place the value (ZRL places a zero)
            if (FINAL && zpos < band_hi) base[(size_t) N * 64 + zpos] = (int16_t) ((uint32_t) v << al);
            z += adv;
            if (z >= band_hi) {
                done += 1;
                N += 1;
                z = band_lo;
            }
        }
        if (FINAL && N >= N_total) break;
    }
    st.p = end_bit - (uint32_t) left;
    st.z = (uint16_t) z;
    st.b = 0;
    return done;
}

template <bool FINAL>
__device__ __forceinline__ uint32_t par_run_ac_auto(const ParIO &io, ParseState &st, const uint32_t end_bit, const uint32_t count_bits,
                                                    const uint32_t tab, const int band_lo, const int band_hi, const int al, bool &bad,
                                                    uint32_t N, const uint32_t N_total, int16_t *base)
{
    const uint32_t mask = __activemask();
    const uint32_t last_word = ((uint32_t) io.lead * 8u + end_bit + 32u + 64u) / 32u + 2u;
    if (__all_sync(mask, last_word < io.wlim)) return par_run_ac<FINAL, true>(io, st, end_bit, count_bits, tab, band_lo, band_hi, al, bad, N, N_total, base);
    return par_run_ac<FINAL, false>(io, st, end_bit, count_bits, tab, band_lo, band_hi, al, bad, N, N_total, base);
}

// ---- progressive DC-first scans (kind 1) on the same machinery -------------------------------------------------------------------
// One DC symbol per block (decode.swift:2960-3004 single component, 3295-3392 interleaved), so the parse state is (bit position,
// block-in-MCU b) -- the tables may differ between the components of an MCU.  The decoding pass only fills the side array of DC
// differences; the prefix sums that follow it (shared with the sequential-scan path) store `prediction << al` into coefficient 0
// and touch nothing else of the block.  DC category 16 and invalid codewords flag the interval for the sequential kernel.
template <bool FINAL, bool SAFE>
__device__ __forceinline__ uint32_t par_run_dc(const ParIO &io, ParseState &st, const uint32_t end_bit, const uint32_t count_bits,
                                               const uint32_t blk0, bool &bad, uint32_t N, const uint32_t N_total, const uint32_t row_blocks,
                                               int16_t *dcdiff, const bool ext)
{
    bad = false;
    if (st.p >= end_bit) return 0;
    if (FINAL && N >= N_total) return 0;
    uint32_t  done = 0;
    int       left = (int) (end_bit - st.p);
    const int slack = (int) (count_bits - end_bit);
    uint32_t  wi, cnt, hi, lo, nxt;
    {
        const uint32_t ab = (uint32_t) io.lead * 8u + st.p;
        wi = ab >> 5;
        cnt = ab & 31u;
        hi = io.word(wi), lo = io.word(wi + 1);
        nxt = __ldg(io.w0 + (SAFE ? wi + 2 : min(wi + 2, io.wlast)));
        wi += 3;
    }
    uint32_t dtab, tabs, next;
    {
        const uint4 q = lds128(blk0 + (uint32_t) st.b * (uint32_t) sizeof(ParBlk) + 16u);
        dtab = q.x, tabs = q.z, next = q.w;
    }
    uint32_t rowpos = (FINAL && ext) ? N % row_blocks : 1u;  // blocks since the start of the MCU row (`extend` scans only)
    while (left > 0) {
        if (cnt >= 32u) {
            hi = lo;
            asm volatile("prmt.b32 %0, %1, 0, 0x0123;" : "=r"(lo) : "r"(nxt));
            nxt = __ldg(io.w0 + (SAFE ? wi : min(wi, io.wlast)));
            wi += 1;
            cnt -= 32u;
        }
        const uint32_t top = __funnelshift_l(lo, hi, cnt);
        if (FINAL && ext) {
            if (rowpos == 0u && row_start_all_ones(top, left + slack)) {  // decode.swift:2906-2912, 3214-3220
                bad = true;
                break;
            }
            if (++rowpos == row_blocks) rowpos = 0u;
        }
        uint32_t ent = lds32(dtab + ((top >> (32 - FAST_BITS)) << 2));
        if (ent & FAST_LINK) {
            const uint32_t rest = (top >> 16) & ((1u << (16 - FAST_BITS)) - 1u);
            ent = lds32(dtab + (ent >> 8) + ((rest >> (ent & 7u)) << 2));
        }
        if (__builtin_expect((ent & FAST_SPECIAL) != 0u, 0)) {  // invalid codeword (the reference reads it as (0, 16)) or category 16
            bad = true;
            break;
        }
        const int total = (int) (ent >> 24);
        if (!SAFE) {
            if (__builtin_expect(total > left + slack, 0)) {  // decode.swift:2791-2794, 2808-2811
                bad = true;
                break;
            }
        }
        if (FINAL) {
            const int      len = (int) (ent & 0x7fu), size = (int) __byte_perm(ent, 0, 0x4441);
            const uint32_t top2 = top << len;
            const uint32_t tail = __funnelshift_rc(top2, 0u, 32 - size);
            dcdiff[N] = (int16_t) ((int) top2 >= 0 ? (int) (tail + (0xffffffffu << size) + 1u) : (int) tail);  // T.81 EXTEND
        }
        cnt += (uint32_t) total;
        left -= total;
        done += 1;
        N += 1;
        {
            const uint4 q = lds128(next);
            dtab = q.x, tabs = q.z, next = q.w;
        }
        if (FINAL && N >= N_total) break;
    }
    st.p = end_bit - (uint32_t) left;
    st.z = 0;
    st.b = (uint16_t) ((tabs >> 16) & 0xffu);
    return done;
}

template <bool FINAL>
__device__ __forceinline__ uint32_t par_run_dc_auto(const ParIO &io, ParseState &st, const uint32_t end_bit, const uint32_t count_bits,
                                                    const uint32_t blk0, bool &bad, uint32_t N, const uint32_t N_total,
                                                    const uint32_t row_blocks, int16_t *dcdiff, const bool ext)
{
    const uint32_t mask = __activemask();
    const uint32_t last_word = ((uint32_t) io.lead * 8u + end_bit + 32u + 64u) / 32u + 2u;
    if (__all_sync(mask, last_word < io.wlim)) return par_run_dc<FINAL, true>(io, st, end_bit, count_bits, blk0, bad, N, N_total, row_blocks, dcdiff, ext);
    return par_run_dc<FINAL, false>(io, st, end_bit, count_bits, blk0, bad, N, N_total, row_blocks, dcdiff, ext);
}

// ---- prologue shared by the subsequence-parallel kernels ---------------------------------------------------------------------
// shared-memory image of one table set: LutHeader | ParBlk[12] (in the BlkInfo area) | fast tables + sub-tables.  The reference
// LUT behind them stays in global memory (only invalid codewords, EOBn and DC category 16 ever look at it).
template <int NT>
__device__ __forceinline__ void par_stage_tables(const ScanParams &P, const int16_t *plane0, const uint32_t img, const uint32_t tid,
                                                 uint8_t *smem, const uint8_t *lut_img)
{
    constexpr uint32_t PRE = PAR_PRE;
    const uint32_t   sbase = smem_u32(smem), blk0 = sbase + (uint32_t) sizeof(LutHeader);
    ParBlk          *s_blk = reinterpret_cast<ParBlk *>(smem + sizeof(LutHeader));
    const int        nblk = P.mcu_blocks;
    const LutHeader *gh = reinterpret_cast<const LutHeader *>(lut_img);
    const uint32_t   total = gh->total_all;
    uint32_t        *dst = reinterpret_cast<uint32_t *>(smem);
    const uint32_t  *src = reinterpret_cast<const uint32_t *>(lut_img);
    for (uint32_t i = tid; i < sizeof(LutHeader) / 4; i += NT) dst[i] = src[i];
    const uint32_t ref_total = gh->total_entries;
    uint32_t      *d2 = reinterpret_cast<uint32_t *>(smem + PRE);
    for (uint32_t i = tid; i < (total - ref_total + 1) / 2; i += NT) d2[i] = src[(sizeof(LutHeader) + 2 * ref_total) / 4 + i];
    if (tid < 12) {
        const int      b = tid, c = P.blk_comp[b];
        const bool     has = P.plane[c] != nullptr;
        const uint32_t ux = (uint32_t) P.ux[c], uy = (uint32_t) P.uy[c], fx = (uint32_t) P.fx[c], fy = (uint32_t) P.fy[c];
        const uint32_t dx = P.blk_dx[b], dy = P.blk_dy[b];
        const uint32_t base_blk = has ? (uint32_t) ((P.plane[c] + (size_t) img * P.image_stride[c] - plane0) / 64) : 0u;
        ParBlk         pb;
        pb.C = base_blk + dx + dy * ux, pb.fx = fx, pb.R = fy * ux;
        const uint32_t LX = (has && ux > dx) ? min((ux - dx + fx - 1u) / fx, 0xffffu) : 0u;
        const uint32_t LY = (has && uy > dy) ? min((uy - dy + fy - 1u) / fy, 0xffffu) : 0u;
        pb.lim = LX | (LY << 16);
        pb.dtab = sbase + PRE + 2u * (gh->fast[P.dc[c]] - ref_total), pb.atab = sbase + PRE + 2u * (gh->fast[P.ac[c]] - ref_total);
        pb.tabs = (uint32_t) P.dc[c] | ((uint32_t) P.ac[c] << 8) | ((uint32_t) b << 16) | (b == 0 ? 1u << 24 : 0u) | ((uint32_t) c << 25);
        pb.next = blk0 + (uint32_t) ((b + 1 == nblk) ? 0 : b + 1) * (uint32_t) sizeof(ParBlk) + 16u;
        s_blk[b] = pb;
    }
}

// interval e of image img, cut into at most T subsequences
__device__ __forceinline__ ParGroup par_setup_group(const ScanParams &P, const uint32_t img, const uint32_t e, const uint32_t T,
                                                    const uint32_t dc_per_interval, uint32_t *flagged, const uint32_t n_images)
{
    ParGroup q;
    memset(&q, 0, sizeof q);
    q.valid = e < P.n_ecs;
    if (!q.valid) return q;
    const int W = P.W, nblk = P.mcu_blocks;
    int64_t   r0, r1;
    if (P.interval == UINT64_MAX) {
        r0 = 0;
        r1 = P.H;
    } else {
        r0 = (int64_t) (((uint64_t) e * P.interval) / (uint32_t) W);
        r1 = (int64_t) (((uint64_t) (e + 1) * P.interval) / (uint32_t) W);
        if (r0 > P.H) r0 = P.H;
        if (r1 > P.H) r1 = P.H;
    }
    q.r0 = (int32_t) r0, q.r1 = (int32_t) r1;
    q.slot = img * P.n_ecs + e;
    const uint64_t n_total = (uint64_t) (r1 - r0) * (uint32_t) W * (uint32_t) nblk;
    const uint64_t o0 = P.offsets[q.slot], o1 = P.offsets[q.slot + 1];
    if (n_total == 0) {
        // sequential scans: nothing to decode, the reference's row loop does not run.  Progressive scans iterate through
        // General.Range2, which yields one row even when empty (decode.swift:3031-3036): left to the sequential kernel.
        if (flagged) {
            flagged[q.slot] = P.kind == 3 ? 1u : 0u;
            if (P.kind != 3 && P.status) P.status[q.slot] = 0;
        }
    } else if ((o1 - o0) > 0x07ffffffull || n_total > dc_per_interval) {
        if (flagged) flagged[q.slot] = 1;  // 32-bit bit positions / side-array capacity: left to the sequential kernel
    } else {
        const uint8_t *base = P.ecs + o0;
        q.N_total = (uint32_t) n_total;
        q.io.nbytes = (int32_t) (o1 - o0);
        q.io.lead = (int32_t) (reinterpret_cast<uintptr_t>(base) & 3);
        q.io.w0 = reinterpret_cast<const uint32_t *>(base - q.io.lead);
        q.io.wlim = (uint32_t) (q.io.lead + q.io.nbytes) / 4;
        q.io.wlast = q.io.nbytes ? (uint32_t) (q.io.lead + q.io.nbytes - 1) / 4 : 0u;
        q.count = 8u * (uint32_t) q.io.nbytes;
        uint32_t B = (q.count + T - 1) / T;
        B = (B + 31u) & ~31u;
        if (B < (uint32_t) PAR_MIN_BITS) B = PAR_MIN_BITS;
        q.B = B;
        q.S = q.count ? (q.count + B - 1) / B : 1u;  // <= T
        // The decoders read ahead of the cursor without looking (four words; a lane that finishes its block on a corrupt stream may
        // run ~2 Kbit past the interval's end): fine wherever more data follows, clamped to the interval's last word for the
        // intervals within 512 bytes of the end of the buffer.
        q.io.clamp = (o1 + 512 > P.offsets[(size_t) n_images * P.n_ecs]) ? 1u : 0u;
    }
    return q;
}

// The kernel is latency-bound (every lane walks its own dependent chain of table look-ups), so resident warps are what
// counts: 8 CTAs per SM = 32 warps needs <= 64 registers (62 used, no spills) and <= 27.5 KB of shared memory per CTA
// (8-bit first-level tables + PAR_ALIAS).  Measured on the bench workload: 6 CTAs 2.79 ms, 7 CTAs 2.71 ms, 8 CTAs 2.39 ms.
#ifndef PAR_MIN_CTAS
#define PAR_MIN_CTAS 8
#endif
#ifndef PAR_ALIAS
#define PAR_ALIAS 1
#endif
// tshift: log2 of the threads per interval (4 .. 7); warm_bits: speculative warm-up before a subsequence's first bit;
// buf_off: the threads' block buffers (PAR_BUF_STRIDE bytes each) in the dynamic shared memory
constexpr int MODE_SEQ = 0, MODE_AC = 1, MODE_DC = 2;  // sequential scan | progressive AC-first | progressive DC-first
template <int NT, int MIN_CTAS, int MODE>
__global__ void __launch_bounds__(NT, MIN_CTAS)
k_decode_par(const __grid_constant__ ScanParams P, int16_t *const plane0, int16_t *const dcdiff_all, const uint32_t dc_per_interval,
             uint32_t *const flagged, uint32_t *const stats, const int tshift, const uint32_t warm_bits, const uint32_t buf_off)
{
    extern __shared__ __align__(16) uint8_t smem[];
    // The records of the synchronisation (exit / entry states, block counts, checkpoints, work list) are dead once the decoding
    // pass starts, and the threads' block buffers are unused until then: with PAR_ALIAS the records live in the buffer area
    // (4.9 KB less shared memory per CTA: 7 CTAs per SM instead of 6).  AC scans have no block buffers and keep static arrays.
    constexpr bool AC = MODE == MODE_AC, DC = MODE == MODE_DC;
    constexpr bool ALIAS = PAR_ALIAS && MODE == MODE_SEQ;
    static_assert(NT * (8 + 8 + 4 + 2 + 4 * PAR_NSEG) <= NT * (int) PAR_BUF_STRIDE, "records fit in the block buffers");
    __shared__ uint64_t s_exit_st[ALIAS ? 1 : NT], s_entry_st[ALIAS ? 1 : NT];
    __shared__ uint32_t s_cnt_st[ALIAS ? 1 : NT];
    __shared__ uint16_t s_work_st[ALIAS ? 1 : NT];
    __shared__ uint32_t s_nwork;
    __shared__ uint32_t s_warp[NT / 32];
    __shared__ ParGroup s_grp[NT / 16];
    __shared__ __align__(16) uint32_t s_ck_st[ALIAS ? 1 : PAR_NSEG][ALIAS ? 4 : NT];  // checkpoints of every subsequence's recorded parse (packed, see parse_sub)
    __shared__ __align__(16) uint2 s_fq[ALIAS ? NT : 1];  // the warps' flush queues (without ALIAS they reuse the checkpoint array)
    // (the block buffers start at the first 128-byte boundary at or after buf_off when they are swizzled)
    const uint32_t bufs_off = PAR_SWIZZLE ? ((smem_u32(smem) + buf_off + 127u) & ~127u) - smem_u32(smem) : buf_off;
    // the lanes' stream rings follow the block buffers (sequential scans only)
    const uint32_t ring = ((smem_u32(smem) + bufs_off + (uint32_t) NT * PAR_BUF_STRIDE + 31u) & ~31u) + threadIdx.x * 4u * PAR_RING;
    uint64_t *const s_exit = ALIAS ? reinterpret_cast<uint64_t *>(smem + bufs_off) : s_exit_st;
    uint64_t *const s_entry = ALIAS ? s_exit + NT : s_entry_st;
    uint32_t (*const s_ck)[NT] = ALIAS ? reinterpret_cast<uint32_t (*)[NT]>(s_entry + NT) : reinterpret_cast<uint32_t (*)[NT]>(&s_ck_st[0][0]);
    uint32_t *const s_cnt = ALIAS ? &s_ck[PAR_NSEG][0] : s_cnt_st;
    uint16_t *const s_work = ALIAS ? reinterpret_cast<uint16_t *>(s_cnt + NT) : s_work_st;
    // PAR_DC_INLINE: cumulative DC sums (four 16-bit sums in a uint2) at every checkpoint of every subsequence's recorded parse, in the
    // same dead buffer area (54 + 64 of its 128 bytes per thread); short intervals only (the scan of the sums stays inside a warp)
    constexpr bool DCI_OK = PAR_DC_INLINE && ALIAS && (22 + 12 * PAR_NSEG) <= (int) PAR_BUF_STRIDE && (NT % 4) == 0;
    const bool     dci = DCI_OK && tshift <= 5;
    uint2 (*const s_dcs)[NT] = reinterpret_cast<uint2 (*)[NT]>(reinterpret_cast<uint8_t *>(s_work) + 2 * NT);
    const uint32_t   img = blockIdx.y, tid = threadIdx.x;
    const uint32_t   T = 1u << tshift, G = NT >> tshift;
#ifdef PAR_INSTRUMENT  // -DPAR_INSTRUMENT builds: with JPEG_SM100_PAR_STATS, cycles per phase (thread 0 of the CTA)
    long long        t_phase = stats ? clock64() : 0;
#define PAR_PHASE(k)                                                                                                 \
    do {                                                                                                             \
        if (stats && tid == 0) {                                                                                     \
            const long long now_ = clock64();                                                                        \
            atomicAdd(reinterpret_cast<unsigned long long *>(stats) + 4 + (k), (unsigned long long) (now_ - t_phase)); \
            t_phase = now_;                                                                                          \
        }                                                                                                            \
    } while (0)
#else
#define PAR_PHASE(k)
#endif
    const uint32_t   g = tid >> tshift, l = tid & (T - 1u);
    const uint8_t   *lut_img = P.luts + (size_t) img * P.lut_stride;
    const LutHeader *hdr = reinterpret_cast<const LutHeader *>(smem);
    ParBlk          *s_blk = reinterpret_cast<ParBlk *>(smem + sizeof(LutHeader));
    const uint32_t   sbase = smem_u32(smem), blk0 = sbase + (uint32_t) sizeof(LutHeader);
    const uint16_t  *ref_entries = reinterpret_cast<const uint16_t *>(lut_img + sizeof(LutHeader));
    const int        W = P.W, nblk = P.mcu_blocks;
    par_stage_tables<NT>(P, plane0, img, tid, smem, lut_img);
    if (tid >= 32 && tid < 32 + NT / 16) {  // one thread per interval of the CTA: where its bytes are, how it is cut
        if (tid - 32 < G)
            s_grp[tid - 32] = par_setup_group(P, img, blockIdx.x * G + (tid - 32), T, dc_per_interval, flagged, gridDim.y);
        else  // NT is not a multiple of T: the threads past the last whole interval idle (S = 0)
            memset(&s_grp[tid - 32], 0, sizeof(ParGroup));
    }
    __syncthreads();
    PAR_PHASE(0);
    const ParIO    io = s_grp[g].io;
    const uint32_t count = s_grp[g].count, B = s_grp[g].B, S = s_grp[g].S;
    const bool     active = l < S;
    const uint32_t start_bit = l * B, end_bit = (l + 1 == S) ? count : (l + 1) * B;

    // ---- synchronisation.  Every subsequence is parsed once from a guessed state (optionally after a speculative warm-up over
    // the warm_bits before it) and leaves checkpoints -- parse state and block count at PAR_NSEG positions.  Then, round by
    // round, every subsequence whose entry differs from its predecessor's exit is re-parsed from that exit, but only until
    // the new parse MERGES with the recorded one (same state at a checkpoint: Huffman streams self-synchronise after a few
    // hundred bits); the rest of the record stays valid, so a correction costs a fraction of a subsequence.  Thread 0 of an
    // interval starts from the true state; once no entry differs from its predecessor's exit every record is exact.
    // The subsequences to redo are compacted into a work list so that they occupy the lanes of as few warps as possible.
    ParseState st;
    bool       bad;
    // parses subsequence `sid` = [s_bit, e_bit) of `qio` from `st`, checkpoint by checkpoint; returns its block count and exit
    auto parse_sub = [&](const ParIO &qio, const uint32_t qcount, const uint32_t s_bit, const uint32_t e_bit, const uint32_t sid,
                         const bool first_time, uint64_t &exit_out) -> uint32_t {
        const uint32_t seglen = (e_bit - s_bit) / PAR_NSEG;
        uint32_t       cum = 0;
        uint2          dcs = make_uint2(0u, 0u);  // (PAR_DC_INLINE) DC differences of the blocks that start in the subsequence so far
#pragma unroll 1
        for (uint32_t k = 0; k < (uint32_t) PAR_NSEG; ++k) {
            const uint32_t seg_end = (k + 1 == (uint32_t) PAR_NSEG) ? e_bit : s_bit + (k + 1) * seglen;
            if (AC) cum += par_run_ac_auto<false>(qio, st, seg_end, qcount, s_blk[0].atab, P.band_lo, P.band_hi, P.al, bad, 0, 0, nullptr);
            else if (DC) cum += par_run_dc_auto<false>(qio, st, seg_end, qcount, blk0, bad, 0, 0, 1u, nullptr, false);
            else if (DCI_OK && dci) cum += par_parse_auto<true>(qio, st, seg_end, qcount, blk0, ring, &dcs);
            else cum += par_parse_auto(qio, st, seg_end, qcount, blk0, ring);
            // checkpoint = (overshoot past seg_end (< 32), z, b) in 16 bits + blocks so far in 16 bits; 0xffff....: unusable
            const uint32_t over = st.p - seg_end;
            const uint32_t code = (over < 32u && cum < 0xffffu) ? (over | ((uint32_t) st.z << 5) | ((uint32_t) st.b << 11) | (cum << 16)) : 0xffffffffu;
            const uint32_t old = s_ck[k][sid];
            if (!first_time && code != 0xffffffffu && old != 0xffffffffu && (code & 0xffffu) == (old & 0xffffu)) {
                // merged: from here on the recorded parse is this parse; the later checkpoints keep their states, their counts shift
                const uint32_t delta = cum - (old >> 16);
                for (uint32_t kk = k; kk < (uint32_t) PAR_NSEG; ++kk) {
                    const uint32_t o = s_ck[kk][sid], c2 = (o >> 16) + delta;
                    s_ck[kk][sid] = (o == 0xffffffffu || c2 >= 0xffffu) ? 0xffffffffu : ((o & 0xffffu) | (c2 << 16));
                }
                if (DCI_OK && dci) {  // the recorded sums of the tail shift by what this parse found differently up to here
                    const uint2 od = s_dcs[k][sid];
                    const uint2 dd = make_uint2(sub16x2(dcs.x, od.x), sub16x2(dcs.y, od.y));
                    for (uint32_t kk = k; kk < (uint32_t) PAR_NSEG; ++kk) {
                        const uint2 o2 = s_dcs[kk][sid];
                        s_dcs[kk][sid] = make_uint2(add16x2(o2.x, dd.x), add16x2(o2.y, dd.y));
                    }
                }
                exit_out = s_exit[sid];
                return s_cnt[sid] + delta;
            }
            s_ck[k][sid] = code;
            if (DCI_OK && dci) s_dcs[k][sid] = dcs;
        }
        exit_out = pack_state(st.p, st.z, st.b);
        return cum;
    };
    // DC-first scans of several components (kind 1, MODE_DC): one symbol per block and nothing else in the stream, so a parse that
    // starts with the wrong block-in-MCU index b keeps the wrong table sequence for good -- nothing ever re-synchronises b, and the
    // rounds below degenerate into a serial walk (measured: 511 rounds for a 4K scan without DRI, 25 ms).  So round 0 parses every
    // subsequence once per possible b at the start of its warm-up (once in all when every component uses the same DC table: then
    // the parse does not depend on b and the exit index follows from the symbol count) and keeps (entry, exit, count) of each
    // hypothesis; before every round one thread per interval walks the chain and adopts, subsequence by subsequence, the
    // hypothesis whose entry is its predecessor's exit.  Only a subsequence whose warm-up did not synchronise is left to a round.
    bool dc_uniform = true;
    for (int c = 1; c < P.n_comp; ++c) dc_uniform = dc_uniform && P.dc[c] == P.dc[0];
    // (short intervals of 16 or 32 subsequences converge in a few rounds anyway: there the extra parses cost more than they save)
    const uint32_t  H = (DC && tshift >= 6) ? (dc_uniform ? 1u : (uint32_t) nblk) : 0u;
    uint64_t *const s_hyp = reinterpret_cast<uint64_t *>(smem + ((buf_off + 15u) & ~15u));  // [NT][H]: entry over | b << 6 | exit over << 10 | b << 16 | count << 32
    if (DC && H != 0u && active && l > 0) {
        uint64_t first_entry = 0, first_exit = 0;
        uint32_t first_cnt = 0;
        for (uint32_t h = 0; h < H; ++h) {
            st.p = start_bit > warm_bits ? start_bit - warm_bits : 0u, st.z = 0, st.b = (uint16_t) h;
            if (warm_bits) par_run_dc_auto<false>(io, st, start_bit, count, blk0, bad, 0, 0, 1u, nullptr, false);
            else st.p = start_bit;
            const uint32_t e_over = st.p - start_bit, e_b = st.b;
            const uint64_t e_state = pack_state(st.p, 0, st.b);
            const uint32_t c = par_run_dc_auto<false>(io, st, end_bit, count, blk0, bad, 0, 0, 1u, nullptr, false);
            const uint32_t x_over = st.p - end_bit;
            s_hyp[(size_t) tid * H + h] = (uint64_t) (e_over < 32u ? e_over : 63u) | ((uint64_t) e_b << 6) | ((uint64_t) (x_over < 32u ? x_over : 63u) << 10) |
                                          ((uint64_t) st.b << 16) | ((uint64_t) c << 32);
            if (h == 0) first_entry = e_state, first_exit = pack_state(st.p, 0, st.b), first_cnt = c;
        }
        s_entry[tid] = first_entry, s_exit[tid] = first_exit, s_cnt[tid] = first_cnt;
        for (uint32_t k = 0; k < (uint32_t) PAR_NSEG; ++k) s_ck[k][tid] = 0xffffffffu;  // (no checkpoints: a re-parse never merges)
    } else if (active) {
        st.p = start_bit > warm_bits ? start_bit - warm_bits : 0u, st.z = AC ? (uint16_t) P.band_lo : 0, st.b = 0;
        if (l > 0 && warm_bits) {
            if (AC) par_run_ac_auto<false>(io, st, start_bit, count, s_blk[0].atab, P.band_lo, P.band_hi, P.al, bad, 0, 0, nullptr);
            else if (DC) par_run_dc_auto<false>(io, st, start_bit, count, blk0, bad, 0, 0, 1u, nullptr, false);
            else par_parse_auto(io, st, start_bit, count, blk0, ring);
        }
        if (l > 0 && !warm_bits) st.p = start_bit;
        s_entry[tid] = pack_state(st.p, st.z, st.b);
        uint64_t       x;
        const uint32_t c = parse_sub(io, count, start_bit, end_bit, tid, true, x);
        s_cnt[tid] = c;
        s_exit[tid] = x;
    } else {
        s_entry[tid] = 0, s_exit[tid] = 0, s_cnt[tid] = 0;
    }
    __syncthreads();
    PAR_PHASE(1);
    uint32_t n_redo = 0, n_rounds = 0;
    for (uint32_t round = 1; round <= T + 1; ++round) {
        if (DC && H != 0u) {
            if (active && l == 0)
                for (uint32_t j = 1; j < S; ++j) {
                    const uint64_t prev = s_exit[tid + j - 1];
                    if (prev == s_entry[tid + j]) continue;
                    const ParseState ps = unpack_state(prev);
                    const uint32_t   over = ps.p - j * B;
                    bool             found = false;
                    for (uint32_t h = 0; h < H && !found && over < 32u; ++h) {
                        const uint64_t e = s_hyp[(size_t) (tid + j) * H + h];
                        if ((uint32_t) (e & 63u) != over || (!dc_uniform && (uint32_t) ((e >> 6) & 15u) != ps.b)) continue;
                        const uint32_t x_over = (uint32_t) (e >> 10) & 63u, cnt = (uint32_t) (e >> 32);
                        if (x_over == 63u && j + 1 < S) continue;  // (only the last subsequence may end without a usable exit)
                        const uint32_t x_b = dc_uniform ? (ps.b + cnt) % (uint32_t) nblk : (uint32_t) ((e >> 16) & 15u);
                        const uint32_t e_bit = (j + 1 == S) ? count : (j + 1) * B;
                        s_entry[tid + j] = prev;
                        s_exit[tid + j] = pack_state(e_bit + (x_over & 31u), 0, (int) x_b);
                        s_cnt[tid + j] = cnt;
                        found = true;
                    }
                    if (!found) break;
                }
            __syncthreads();
        }
        const bool redo = active && l >= 1 && s_exit[tid - 1] != s_entry[tid];
        n_redo += redo ? 1u : 0u;
        n_rounds = round;
        if (tid == 0) s_nwork = 0;
        __syncthreads();
        if (redo) s_work[atomicAdd(&s_nwork, 1u)] = (uint16_t) tid;
        __syncthreads();
        const uint32_t nwork = s_nwork;
        if (nwork == 0) break;
        uint64_t x = 0;
        if (tid < nwork) {
            const uint32_t  sid = s_work[tid];
            const ParGroup &q = s_grp[sid >> tshift];
            const uint32_t  ll = sid & (T - 1u);
            const ParIO     qio = q.io;
            const uint64_t  entry = s_exit[sid - 1];
            const uint32_t  e_bit = (ll + 1 == q.S) ? q.count : (ll + 1) * q.B;
            st = unpack_state(entry);
            const uint32_t c = parse_sub(qio, q.count, ll * q.B, e_bit, sid, false, x);
            // every work item reads exit[sid - 1] before any item writes exit[sid]: the write is deferred past a barrier
            s_cnt[sid] = c;
            s_entry[sid] = entry;
        }
        __syncthreads();
        if (tid < nwork) s_exit[s_work[tid]] = x;
        __syncthreads();
    }
    PAR_PHASE(2);
    const uint32_t my_cnt = s_cnt[tid];
    const uint64_t my_entry = s_entry[tid];
    const uint2    my_dcs = (DCI_OK && dci && active) ? s_dcs[PAR_NSEG - 1][tid] : make_uint2(0u, 0u);
    // ---- first block of every subsequence: exclusive scan of the block counts within the interval -------------------------
    uint32_t incl = my_cnt;
    const int      lane = tid & 31, wid = tid >> 5;
    const uint32_t seg = (T < 32u ? T : 32u) - 1u;  // intervals of 16 threads share a warp: segmented scan
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, incl, d);
        if ((lane & seg) >= (uint32_t) d) incl += y;
    }
    // (PAR_DC_INLINE) the same exclusive scan over the per-subsequence DC sums: the prediction at every thread's first own block
    uint2 pred = my_dcs;
    if (DCI_OK && dci) {
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t yx = __shfl_up_sync(0xffffffffu, pred.x, d), yy = __shfl_up_sync(0xffffffffu, pred.y, d);
            if ((lane & seg) >= (uint32_t) d) pred.x = add16x2(pred.x, yx), pred.y = add16x2(pred.y, yy);
        }
        pred.x = sub16x2(pred.x, my_dcs.x), pred.y = sub16x2(pred.y, my_dcs.y);
    }
    if (lane == 31) s_warp[wid] = incl;
    __syncthreads();
    PAR_PHASE(3);
    uint32_t before = incl - my_cnt;
    if (tshift > 5)
        for (int w = (int) (g << (tshift - 5)); w < wid; ++w) before += s_warp[w];
    // ---- the one real decoding pass ---------------------------------------------------------------------------------------
    const uint32_t N_total = s_grp[g].N_total, slot = s_grp[g].slot;
    int16_t *const dcdiff = dcdiff_all + (size_t) slot * dc_per_interval;
    if (MODE == MODE_SEQ) {  // every lane of every warp takes part in the cooperative flush
        static_assert(PAR_NSEG * 4 >= 8, "the flush queues fit in the checkpoint array");
        const bool     go = active && before < N_total;
        const uint32_t done = par_decode_auto(go, io, unpack_state(my_entry), end_bit, count, blk0, nblk, bad, before, N_total, W,
                                              s_grp[g].r0, plane0, dcdiff, sbase + bufs_off + tid * PAR_BUF_STRIDE,
                                              (ALIAS ? smem_u32(&s_fq[0]) : smem_u32(&s_ck_st[0][0])) + (tid & ~31u) * 8u, ring, P.extend != 0,
                                              DCI_OK && dci, pred);
        if (active) {
            if (bad || (go && done != my_cnt)) atomicOr(&s_grp[g].bad, 1u);
            atomicAdd(&s_grp[g].total, done);
            if (stats) {  // JPEG_SM100_PAR_STATS: why intervals get flagged
                if (bad) atomicAdd(&stats[20], 1u);
                if (go && done != my_cnt) atomicAdd(&stats[21], 1u);
            }
        }
    } else if (active) {
        st = unpack_state(my_entry);
        uint32_t done = 0;
        bad = false;
        if (before < N_total) {
            if (AC)  // single component, blocks in raster order: block N of the interval is N blocks after its first
                done = par_run_ac_auto<true>(io, st, end_bit, count, s_blk[0].atab, P.band_lo, P.band_hi, P.al, bad, before, N_total,
                                             P.plane[0] + (size_t) img * P.image_stride[0] + (size_t) s_grp[g].r0 * (size_t) W * 64);
            else
                done = par_run_dc_auto<true>(io, st, end_bit, count, blk0, bad, before, N_total, (uint32_t) W * (uint32_t) nblk, dcdiff, P.extend != 0);
        }
        // a subsequence that does not produce the blocks the synchronisation counted for it (or that was cut short because the
        // interval is complete while data remains) leaves the interval to the sequential decoder
        if (bad || (before < N_total && done != my_cnt)) atomicOr(&s_grp[g].bad, 1u);
        atomicAdd(&s_grp[g].total, done);
    }
    __syncthreads();
    PAR_PHASE(4);
    // every expected block must have been completed (a short stream is a truncation in the reference)
    const bool mine = S != 0u;  // this interval was decoded here
    if (mine && l == 0) {
        const bool f = s_grp[g].bad != 0u || s_grp[g].total != N_total;
        if (stats) {
            if (f) atomicAdd(&stats[22], 1u);
            if (s_grp[g].total != N_total) atomicAdd(&stats[23], 1u);
        }
        s_grp[g].bad = f ? 1u : 0u;
        flagged[slot] = f ? 1u : 0u;
        if (!f && P.status) P.status[slot] = 0;
    }
    __syncthreads();
    // ---- DC differences -> DC coefficients: decode.swift:3248-3254 (wrapping Int16 prediction, reset per interval); one warp
    // per interval runs a 16-bit prefix sum per component over the side array the pass above filled (L2-resident).
    // Out-of-plane blocks take part in the prediction but are not stored (decode.swift:1470-1475).
    // Work items (interval, component) are dealt to the warps; a lane owns a contiguous run of the component's blocks: it sums its
    // differences, the warp scans the lane sums, the lane walks its run again and stores the predictions.
    for (uint32_t item = (uint32_t) wid; !AC && !(DCI_OK && dci) && item < G * (uint32_t) P.n_comp; item += NT / 32) {  // (sequential and DC-first scans)
        const uint32_t  gg = item / (uint32_t) P.n_comp;
        const int       c = (int) (item - gg * (uint32_t) P.n_comp);
        const ParGroup &q = s_grp[gg];
        if (q.S == 0u || q.bad != 0u || !P.plane[c]) continue;
        uint32_t fb = 0;
        for (int cc = 0; cc < c; ++cc) fb += (uint32_t) (P.fx[cc] * P.fy[cc]);
        const uint32_t uW = (uint32_t) W, unblk = (uint32_t) nblk;
        const uint32_t nc = (uint32_t) (P.fx[c] * P.fy[c]);
        const uint32_t K = (uint32_t) (q.r1 - q.r0) * uW * nc;
        const int16_t *dcd = dcdiff_all + (size_t) q.slot * dc_per_interval;
        const uint32_t R = (K + 31u) / 32u, k0 = min((uint32_t) lane * R, K), k1 = min(k0 + R, K);
        int            sum = 0;
        {
            uint32_t mcu = k0 / nc, j = k0 - mcu * nc;
#pragma unroll 4
            for (uint32_t k = k0; k < k1; ++k) {
                sum += (int) dcd[mcu * unblk + fb + j];
                if (++j == nc) j = 0, ++mcu;
            }
        }
        int incl2 = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, incl2, d);
            if (lane >= d) incl2 += y;
        }
        int      run = incl2 - sum;
        uint32_t mcu = k0 / nc, j = k0 - mcu * nc;
        uint32_t my = (uint32_t) q.r0 + mcu / uW, mx = mcu - (mcu / uW) * uW;
#pragma unroll 4
        for (uint32_t k = k0; k < k1; ++k) {
            run += (int) dcd[mcu * unblk + fb + j];
            const ParBlk &pb = s_blk[fb + j];
            if (mx < (pb.lim & 0xffffu) && my < (pb.lim >> 16))
                plane0[(int64_t) (int32_t) (pb.C + mx * pb.fx + my * pb.R) * 64] = (int16_t) ((uint32_t) (int) (short) run << P.al);
            if (++j == nc) {
                j = 0, ++mcu;
                if (++mx == uW) mx = 0, ++my;
            }
        }
    }
    __syncthreads();
    PAR_PHASE(5);
    if (stats) {  // JPEG_SM100_PAR_STATS=1: rounds and re-parses per interval
        atomicAdd(&stats[0], n_redo);
        if (mine && l == 0) {
            atomicAdd(&stats[1], n_rounds);
            atomicAdd(&stats[2], S);
            atomicAdd(&stats[3], 1u);
            atomicMax(&stats[4], n_rounds);
        }
    }
}

// ---- the same decoder for FEW, LARGE intervals: a thread-block cluster per interval -------------------------------------------
// A file without DRI is one entropy-coded segment per scan (the reference's own encoder never emits DRI), i.e. ONE interval per
// image.  A cluster of up to 8 CTAs x 128 threads shares it: subsequence l = cluster rank * 128 + thread.  Everything a thread
// needs from a neighbour -- the predecessor's exit state across a CTA border, the other CTAs' work counts, block totals, DC
// sums -- is read from the owner's shared memory through DSMEM (mapa / ld.shared::cluster) between barrier.cluster syncs;
// nothing goes through global memory except the coefficients and the DC side array.  Same records, same rounds, same
// decoding pass as k_decode_par (see there); grid = (cluster size * n_ecs, n_images), cluster = (cluster size, 1, 1).
constexpr int CL_THREADS = 128;

__global__ void __launch_bounds__(CL_THREADS, 4)
k_decode_par_cluster(const __grid_constant__ ScanParams P, int16_t *const plane0, int16_t *const dcdiff_all, const uint32_t dc_per_interval,
                     uint32_t *const flagged, const uint32_t warm_bits, const uint32_t buf_off)
{
    namespace cg = cooperative_groups;
    constexpr int NT = CL_THREADS;
    cg::cluster_group cluster = cg::this_cluster();
    const uint32_t    crank = cluster.block_rank(), csize = cluster.num_blocks();
    extern __shared__ __align__(16) uint8_t smem[];
    __shared__ uint64_t s_exit[NT], s_entry[NT];
    __shared__ uint32_t s_cnt[NT];
    __shared__ uint16_t s_work[NT];
    __shared__ uint32_t s_nwork, s_ctatotal, s_bad, s_total;
    __shared__ uint32_t s_warp[NT / 32];
    __shared__ int      s_dcw[4][NT / 32];
    __shared__ ParGroup s_q;
    __shared__ __align__(16) uint32_t s_ck[PAR_NSEG][NT];
    // PAR_DC_INLINE (see k_decode_par): cumulative DC sums per checkpoint, warp / CTA totals of the per-subsequence sums
    constexpr bool CL_DCI = PAR_DC_INLINE != 0;
    __shared__ uint2    s_dcs[CL_DCI ? PAR_NSEG : 1][CL_DCI ? NT : 1];
    __shared__ uint64_t s_warpd[NT / 32], s_ctad;
    const uint32_t  img = blockIdx.y, tid = threadIdx.x, e = blockIdx.x / csize;
    const uint32_t  T = NT * csize, l = crank * NT + tid;
    const int       lane = tid & 31, wid = tid >> 5;
    const uint8_t  *lut_img = P.luts + (size_t) img * P.lut_stride;
    ParBlk         *s_blk = reinterpret_cast<ParBlk *>(smem + sizeof(LutHeader));
    const uint32_t  sbase = smem_u32(smem), blk0 = sbase + (uint32_t) sizeof(LutHeader);
    const int       W = P.W, nblk = P.mcu_blocks;
    const uint32_t  bufs = PAR_SWIZZLE ? (sbase + buf_off + 127u) & ~127u : sbase + buf_off;  // block buffers, then the stream rings
    const uint32_t  ring = ((bufs + (uint32_t) NT * PAR_BUF_STRIDE + 31u) & ~31u) + tid * 4u * PAR_RING;
    // values other CTAs of the cluster hold in their shared memory
    auto remote64 = [&](uint64_t *p_, uint32_t r) -> uint64_t { return *cluster.map_shared_rank(p_, r); };
    auto remote32 = [&](uint32_t *p_, uint32_t r) -> uint32_t { return *cluster.map_shared_rank(p_, r); };
    par_stage_tables<NT>(P, plane0, img, tid, smem, lut_img);
    if (tid == 32) s_q = par_setup_group(P, img, e, T, dc_per_interval, crank == 0 ? flagged : nullptr, gridDim.y);
    if (tid == 0) s_nwork = 0, s_ctatotal = 0, s_bad = 0, s_total = 0;
    __syncthreads();
    const ParIO    io = s_q.io;
    const uint32_t count = s_q.count, B = s_q.B, S = s_q.S, N_total = s_q.N_total, slot = s_q.slot;
    const bool     active = l < S, mine = S != 0u;
    const uint32_t start_bit = l * B, end_bit = (l + 1 == S) ? count : (l + 1) * B;
    ParseState     st;
    bool           bad;
    auto parse_sub = [&](const uint32_t s_bit, const uint32_t e_bit, const uint32_t sid, const bool first_time, uint64_t &exit_out) -> uint32_t {
        const uint32_t seglen = (e_bit - s_bit) / PAR_NSEG;
        uint32_t       cum = 0;
        uint2          dcs = make_uint2(0u, 0u);
#pragma unroll 1
        for (uint32_t k = 0; k < (uint32_t) PAR_NSEG; ++k) {
            const uint32_t seg_end = (k + 1 == (uint32_t) PAR_NSEG) ? e_bit : s_bit + (k + 1) * seglen;
            if (CL_DCI) cum += par_parse_auto<true>(io, st, seg_end, count, blk0, ring, &dcs);
            else cum += par_parse_auto(io, st, seg_end, count, blk0, ring);
            const uint32_t over = st.p - seg_end;
            const uint32_t code = (over < 32u && cum < 0xffffu) ? (over | ((uint32_t) st.z << 5) | ((uint32_t) st.b << 11) | (cum << 16)) : 0xffffffffu;
            const uint32_t old = s_ck[k][sid];
            if (!first_time && code != 0xffffffffu && old != 0xffffffffu && (code & 0xffffu) == (old & 0xffffu)) {
                const uint32_t delta = cum - (old >> 16);
                for (uint32_t kk = k; kk < (uint32_t) PAR_NSEG; ++kk) {
                    const uint32_t o = s_ck[kk][sid], c2 = (o >> 16) + delta;
                    s_ck[kk][sid] = (o == 0xffffffffu || c2 >= 0xffffu) ? 0xffffffffu : ((o & 0xffffu) | (c2 << 16));
                }
                if (CL_DCI) {
                    const uint2 od = s_dcs[k][sid];
                    const uint2 dd = make_uint2(sub16x2(dcs.x, od.x), sub16x2(dcs.y, od.y));
                    for (uint32_t kk = k; kk < (uint32_t) PAR_NSEG; ++kk) {
                        const uint2 o2 = s_dcs[kk][sid];
                        s_dcs[kk][sid] = make_uint2(add16x2(o2.x, dd.x), add16x2(o2.y, dd.y));
                    }
                }
                exit_out = s_exit[sid];
                return s_cnt[sid] + delta;
            }
            s_ck[k][sid] = code;
            if (CL_DCI) s_dcs[k][sid] = dcs;
        }
        exit_out = pack_state(st.p, st.z, st.b);
        return cum;
    };
    // ---- round 0 ----
    if (active) {
        st.p = start_bit > warm_bits ? start_bit - warm_bits : 0u, st.z = 0, st.b = 0;
        if (l > 0 && warm_bits) par_parse_auto(io, st, start_bit, count, blk0, ring);
        if (l > 0 && !warm_bits) st.p = start_bit;
        s_entry[tid] = pack_state(st.p, st.z, st.b);
        uint64_t       x;
        const uint32_t c = parse_sub(start_bit, end_bit, tid, true, x);
        s_cnt[tid] = c;
        s_exit[tid] = x;
    } else {
        s_entry[tid] = 0, s_exit[tid] = 0, s_cnt[tid] = 0;
    }
    cluster.sync();
    // ---- synchronisation rounds (the predecessor of a CTA's first subsequence lives in the previous CTA) ----
    for (uint32_t round = 1; round <= T + 1; ++round) {
        uint64_t prev = 0;
        if (active && l >= 1) prev = tid > 0 ? s_exit[tid - 1] : remote64(&s_exit[NT - 1], crank - 1);
        const bool redo = active && l >= 1 && prev != s_entry[tid];
        if (tid == 0) s_nwork = 0;
        __syncthreads();
        if (redo) s_work[atomicAdd(&s_nwork, 1u)] = (uint16_t) tid;
        cluster.sync();
        uint32_t pending = 0;
        for (uint32_t r = 0; r < csize; ++r) pending += remote32(&s_nwork, r);
        if (pending == 0) break;  // (the same sum in every thread of the cluster)
        const uint32_t nwork = s_nwork;
        uint64_t       x = 0;
        if (tid < nwork) {
            const uint32_t sid = s_work[tid], ll = crank * NT + sid;
            const uint64_t entry = sid > 0 ? s_exit[sid - 1] : remote64(&s_exit[NT - 1], crank - 1);
            const uint32_t e_bit = (ll + 1 == S) ? count : (ll + 1) * B;
            st = unpack_state(entry);
            const uint32_t c = parse_sub(ll * B, e_bit, sid, false, x);
            s_cnt[sid] = c;
            s_entry[sid] = entry;
        }
        cluster.sync();  // every exit has been read ...
        if (tid < nwork) s_exit[s_work[tid]] = x;
        cluster.sync();  // ... before any is replaced
    }
    // ---- first block of every subsequence: scan inside the CTA, then the totals of the CTAs before this one ----
    const uint32_t my_cnt = s_cnt[tid];
    uint32_t       incl = my_cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += y;
    }
    if (lane == 31) s_warp[wid] = incl;
    __syncthreads();
    uint32_t before = incl - my_cnt;
    for (int w = 0; w < wid; ++w) before += s_warp[w];
    if (tid == NT - 1) s_ctatotal = before + my_cnt;
    // (PAR_DC_INLINE) the same two-level exclusive scan over the per-subsequence DC sums: warp, CTA, the CTAs before this one
    uint2 pred = make_uint2(0u, 0u);
    if (CL_DCI) {
        const uint2 my_dcs = active ? s_dcs[PAR_NSEG - 1][tid] : make_uint2(0u, 0u);
        uint2       inc = my_dcs;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t yx = __shfl_up_sync(0xffffffffu, inc.x, d), yy = __shfl_up_sync(0xffffffffu, inc.y, d);
            if (lane >= d) inc.x = add16x2(inc.x, yx), inc.y = add16x2(inc.y, yy);
        }
        if (lane == 31) s_warpd[wid] = (uint64_t) inc.x | ((uint64_t) inc.y << 32);
        __syncthreads();
        pred = make_uint2(sub16x2(inc.x, my_dcs.x), sub16x2(inc.y, my_dcs.y));
        uint2 cta = make_uint2(0u, 0u);
        for (int w = 0; w < NT / 32; ++w) {
            const uint64_t t = s_warpd[w];
            if (w < wid) pred.x = add16x2(pred.x, (uint32_t) t), pred.y = add16x2(pred.y, (uint32_t) (t >> 32));
            cta.x = add16x2(cta.x, (uint32_t) t), cta.y = add16x2(cta.y, (uint32_t) (t >> 32));
        }
        if (tid == 0) s_ctad = (uint64_t) cta.x | ((uint64_t) cta.y << 32);
    }
    cluster.sync();
    for (uint32_t r = 0; r < crank; ++r) before += remote32(&s_ctatotal, r);
    if (CL_DCI)
        for (uint32_t r = 0; r < crank; ++r) {
            const uint64_t t = remote64(&s_ctad, r);
            pred.x = add16x2(pred.x, (uint32_t) t), pred.y = add16x2(pred.y, (uint32_t) (t >> 32));
        }
    // ---- the decoding pass ----
    int16_t *const dcdiff = dcdiff_all + (size_t) slot * dc_per_interval;
    {
        static_assert(sizeof(s_ck) >= NT * 8, "the flush queues live in the checkpoint array");
        const bool     go = active && before < N_total;
        const uint32_t done = par_decode_auto(go, io, unpack_state(s_entry[tid]), end_bit, count, blk0, nblk, bad, before, N_total, W, s_q.r0,
                                              plane0, dcdiff, bufs + tid * PAR_BUF_STRIDE,
                                              smem_u32(&s_ck[0][0]) + (tid & ~31u) * 8u, ring, P.extend != 0, CL_DCI, pred);
        if (active) {
            if (bad || (go && done != my_cnt)) atomicOr(&s_bad, 1u);
            atomicAdd(&s_total, done);
        }
    }
    cluster.sync();
    uint32_t bad_all = 0, total_all = 0;
    for (uint32_t r = 0; r < csize; ++r) bad_all |= remote32(&s_bad, r), total_all += remote32(&s_total, r);
    const bool f = mine && (bad_all != 0u || total_all != N_total);
    if (mine && crank == 0 && tid == 0) {
        flagged[slot] = f ? 1u : 0u;
        if (!f && P.status) P.status[slot] = 0;
    }
    // ---- DC differences -> DC coefficients.  Every warp of the cluster takes a contiguous range of each component's blocks:
    // lane runs are summed, warps publish their totals, everybody adds up the totals before its own (DSMEM), second walk stores.
    // The side array was written by other SMs: L2 loads (ld.global.cg), ordered by the cluster barrier above.
    const bool     do_dc = mine && !f && !CL_DCI;  // (PAR_DC_INLINE: the DC values left with their blocks)
    const uint32_t warps_cta = NT / 32, gw = crank * warps_cta + (uint32_t) wid, n_gw = csize * warps_cta;
    int            lane_ex[4] = {0, 0, 0, 0};
    uint32_t       rk0[4] = {0, 0, 0, 0}, rk1[4] = {0, 0, 0, 0};
    {
        uint32_t fb = 0;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            if (c >= P.n_comp) break;
            const uint32_t nc = (uint32_t) (P.fx[c] * P.fy[c]);
            if (do_dc && P.plane[c]) {
                const uint32_t K = (uint32_t) (s_q.r1 - s_q.r0) * (uint32_t) W * nc;
                const uint32_t per = (K + n_gw - 1) / n_gw, w0 = min(gw * per, K), w1 = min(w0 + per, K);
                const uint32_t R = (w1 - w0 + 31u) / 32u, k0 = min(w0 + (uint32_t) lane * R, w1), k1 = min(k0 + R, w1);
                int            sum = 0;
                uint32_t       mcu = k0 / nc, j = k0 - mcu * nc;
                for (uint32_t k = k0; k < k1; ++k) {
                    sum += (int) __ldcg(dcdiff + mcu * (uint32_t) nblk + fb + j);
                    if (++j == nc) j = 0, ++mcu;
                }
                int inc2 = sum;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const int y = __shfl_up_sync(0xffffffffu, inc2, d);
                    if (lane >= d) inc2 += y;
                }
                lane_ex[c] = inc2 - sum, rk0[c] = k0, rk1[c] = k1;
                if (lane == 31) s_dcw[c][wid] = inc2;
            } else if (lane == 31)
                s_dcw[c][wid] = 0;
            fb += nc;
        }
    }
    cluster.sync();
    if (do_dc) {
        uint32_t fb = 0;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            if (c >= P.n_comp) break;
            const uint32_t nc = (uint32_t) (P.fx[c] * P.fy[c]);
            if (P.plane[c]) {
                // totals of the warps before this one: lane j fetches warp j's (there are at most 32 warps in a cluster of 8)
                int part = 0;
                for (uint32_t j = (uint32_t) lane; j < gw; j += 32) part += *cluster.map_shared_rank(&s_dcw[c][j % warps_cta], j / warps_cta);
#pragma unroll
                for (int d = 16; d >= 1; d >>= 1) part += __shfl_xor_sync(0xffffffffu, part, d);
                int            run = part + lane_ex[c];
                const uint32_t uW = (uint32_t) W;
                uint32_t       mcu = rk0[c] / nc, j = rk0[c] - mcu * nc;
                uint32_t       my = (uint32_t) s_q.r0 + mcu / uW, mx = mcu - (mcu / uW) * uW;
                for (uint32_t k = rk0[c]; k < rk1[c]; ++k) {
                    run += (int) __ldcg(dcdiff + mcu * (uint32_t) nblk + fb + j);
                    const ParBlk &pb = s_blk[fb + j];
                    if (mx < (pb.lim & 0xffffu) && my < (pb.lim >> 16))
                        plane0[(int64_t) (int32_t) (pb.C + mx * pb.fx + my * pb.R) * 64] = (int16_t) ((uint32_t) (int) (short) run << P.al);
                    if (++j == nc) {
                        j = 0, ++mcu;
                        if (++mx == uW) mx = 0, ++my;
                    }
                }
            }
            fb += nc;
        }
    }
    cluster.sync();  // no CTA leaves while another may still read its shared memory
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

// the band of a progressive AC-first scan in the blocks of flagged intervals, before the sequential kernel re-decodes them (the
// band is all zero before its first scan; the parallel decoder may have written part of it)
__global__ void __launch_bounds__(128) k_zero_band_flagged(const __grid_constant__ ScanParams P, const uint32_t *const flagged)
{
    const uint32_t e = blockIdx.x, img = blockIdx.y;
    if (!flagged[(size_t) img * P.n_ecs + e] || !P.plane[0]) return;
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
    int16_t       *pl = P.plane[0] + (size_t) img * P.image_stride[0] + 64 * (size_t) P.ux[0] * (size_t) r0;
    const uint32_t nb = (uint32_t) P.ux[0] * (uint32_t) (r1 - r0), band = (uint32_t) (P.band_hi - P.band_lo);
    for (uint32_t i = threadIdx.x; i < nb * band; i += blockDim.x) pl[(size_t) (i / band) * 64 + P.band_lo + i % band] = 0;
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
__global__ void __launch_bounds__(WARP) k_decode_progressive(const __grid_constant__ ScanParams P, const uint32_t *const only_flagged)
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
    if (only_flagged && only_flagged[(size_t) img * P.n_ecs + e] == 0u) return;  // fallback pass after the parallel decoder
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

// ---- progressive DC refinement scans (kind 2): no Huffman codes at all ---------------------------------------------------------
// decode.swift:3007-3018 (single component), 3395-3445 (interleaved): block n of a restart interval owns bit n of its entropy-coded
// segment and ORs it into coefficient 0 at bit `al` -- a pure gather, one thread per block.  Out-of-plane blocks of partial MCUs
// and components without a plane consume their bit and store nothing.  An interval with fewer bits than blocks is truncated
// (decode.swift:2775-2778).  Intervals whose row range is empty or inverted (General.Range2's quirks) are flagged for
// k_decode_progressive, which reproduces them.
__global__ void __launch_bounds__(256) k_decode_dc_refine(const __grid_constant__ ScanParams P, const uint32_t n_images, uint32_t *const flagged)
{
    const uint32_t nblk = (uint32_t) P.mcu_blocks, W = (uint32_t) P.W;
    for (uint32_t img = blockIdx.z; img < n_images; img += gridDim.z) {
        for (uint32_t e = blockIdx.y; e < P.n_ecs; e += gridDim.y) {
            const size_t slot = (size_t) img * P.n_ecs + e;
            int64_t      r0 = 0, r1 = P.H;
            if (P.interval != UINT64_MAX) {
                r0 = (int64_t) (((uint64_t) e * P.interval) / W);
                r1 = (int64_t) (((uint64_t) (e + 1) * P.interval) / W);
                if (r1 > P.H) r1 = P.H;
            }
            const uint64_t o0 = P.offsets[slot], o1 = P.offsets[slot + 1];
            const uint64_t n_total = r1 > r0 ? (uint64_t) (r1 - r0) * W * nblk : 0;
            const bool     irregular = !(r1 > r0) || n_total > 0xffffffffull || (o1 - o0) > 0x1fffffffull;
            if (blockIdx.x == 0 && threadIdx.x == 0) {
                flagged[slot] = irregular ? 1u : 0u;
                if (!irregular) P.status[slot] = (o1 - o0) * 8 < n_total ? JPEG_SM100_ERR_TRUNCATED_ECS : 0;
            }
            if (irregular) continue;
            const uint8_t *base = P.ecs + o0;
            const uint32_t nbits = (uint32_t) ((o1 - o0) * 8), N = (uint32_t) n_total;
            for (uint32_t n = blockIdx.x * blockDim.x + threadIdx.x; n < N && n < nbits; n += gridDim.x * blockDim.x) {
                const uint32_t mcu = n / nblk, b = n - mcu * nblk;
                const uint32_t row = mcu / W, mx = mcu - row * W, my = (uint32_t) r0 + row;
                const int      c = P.blk_comp[b];
                if (!P.plane[c]) continue;
                const uint32_t x = mx * (uint32_t) P.fx[c] + P.blk_dx[b], y = my * (uint32_t) P.fy[c] + P.blk_dy[b];
                if (x >= (uint32_t) P.ux[c] || y >= (uint32_t) P.uy[c]) continue;
                const uint32_t bit = (__ldg(base + (n >> 3)) >> (7u - (n & 7u))) & 1u;
                int16_t       *dst = P.plane[c] + (size_t) img * P.image_stride[c] + 64 * ((size_t) P.ux[c] * y + x);
                if (bit) *dst = (int16_t) (*dst | (int16_t) (1u << P.al));
            }
        }
    }
}

// ---- progressive AC refinement scans (kind 4) with few, large intervals: masks -> serial parse -> parallel apply ----------------
// decode.swift:3072-3152.  A refinement scan does not self-synchronise: how many correction bits a symbol is followed by depends on
// which coefficients of ITS block were non-zero before the scan, so a parser that does not know its block cannot know its bit
// position either, and the scan is serial in the bit stream.  What CAN be taken off the serial path is everything but the bit
// count: coefficients placed by this scan are never revisited by it, so the "non-zero" test only needs the state before the scan.
//   k_acr_masks  (one thread per block)      64-bit map of the block's non-zero coefficients inside the band;
//   k_acr_parse  (one warp per interval)     walks the symbols with bit operations on the maps: n-th zero -> landing position,
//                                            popcount -> correction bits to skip; inside an EOB run a block is one popcount and the
//                                            warp takes 32 blocks per step with a prefix sum.  Leaves every block's first bit;
//   k_acr_apply  (one thread per block)      decodes its block from that bit and updates the coefficients.
// A file without DRI -- everything the reference's encoder writes -- is ONE interval per scan: the one-thread-per-interval kernel
// walks it coefficient by coefficient through global memory (400 ms for a 4K luma plane).  Anything irregular flags the interval;
// flagged intervals are left untouched here and redone by k_decode_progressive, which owns the reference's error codes.
constexpr uint32_t ACR_EOB = 0x80000000u;

struct AcrReader {  // 1-padded bit stream (jpeg.swift:1873-1916) addressed by absolute bit position; all lanes of a warp may share one
    const uint8_t *base;
    uint32_t       nbytes, count;  // count = 8 * nbytes
    uint32_t       pos;
    __device__ __forceinline__ uint32_t byte_at(uint32_t i) const { return i < nbytes ? (uint32_t) __ldg(base + i) : 0xffu; }
    // the 32 bits at `pos` (MSB first)
    __device__ __forceinline__ uint32_t peek32() const
    {
        const uint32_t b = pos >> 3, s = pos & 7u;
        uint32_t       hi, lo;
        if (b + 8u <= nbytes) {  // the two aligned words around byte b hold bytes b .. b + 4
            const uintptr_t a = reinterpret_cast<uintptr_t>(base + b);
            const uint32_t *w = reinterpret_cast<const uint32_t *>(a & ~uintptr_t(3));
            const uint32_t  sh = (uint32_t) (a & 3u) * 8u;
            const uint32_t  w0 = __byte_perm(__ldg(w), 0, 0x0123), w1 = __byte_perm(__ldg(w + 1), 0, 0x0123);
            hi = __funnelshift_l(w1, w0, sh);
            lo = w1 << sh;  // its top byte is byte b + 4
        } else {
            hi = (byte_at(b) << 24) | (byte_at(b + 1) << 16) | (byte_at(b + 2) << 8) | byte_at(b + 3);
            lo = byte_at(b + 4) << 24;
        }
        return __funnelshift_l(lo, hi, s);
    }
};

// one block of the scan.  m: bit z set <=> coefficient z (band_lo <= z < band_hi) was non-zero before this scan.  APPLY = false only
// moves the bit position; APPLY = true also updates the coefficients (blk = the block's 64 coefficients).  0 or the reference's error.
template <bool APPLY>
__device__ __forceinline__ int acr_block(AcrReader &br, const uint64_t m, int &skip, const int band_lo, const int band_hi, const int al,
                                         const uint16_t *entries, const int tn, const int tz, const uint32_t toff, int16_t *blk)
{
    int z = band_lo;
    while (z < band_hi) {
        int zeroes, delta = 0;
        if (skip > 0) {
            zeroes = 64;
            skip -= 1;
        } else {
            if (!(br.pos < br.count)) return JPEG_SM100_ERR_TRUNCATED_ECS;
            const uint32_t w = br.peek32();
            const uint32_t ent = lut_lookup(entries, tn, tz, toff, w >> 16);
            const int      sym = (int) (ent & 0xffu), len = (int) (ent >> 8);
            br.pos += (uint32_t) len;
            const int sz = sym >> 4, binade = sym & 15;
            // (len <= 16 and the extra bits <= 15: both inside the 32-bit window)
            if (binade == 0) {
                if (sz == 0) {
                    zeroes = 64;
                } else if (sz <= 14) {
                    if (!(br.pos + (uint32_t) sz <= br.count)) return JPEG_SM100_ERR_TRUNCATED_ECS;
                    const int run = (1 << sz) | (int) ((w << len) >> (32 - sz));
                    br.pos += (uint32_t) sz;
                    zeroes = 64;
                    skip = run - 1;
                } else {
                    zeroes = 15;
                }
            } else {
                if (!(br.pos + (uint32_t) binade <= br.count)) return JPEG_SM100_ERR_TRUNCATED_ECS;
                const int v = extend16(binade, (w << len) >> (32 - binade));
                br.pos += (uint32_t) binade;
                if (!(v >= -1 && v <= 1)) return JPEG_SM100_ERR_INVALID_COMPOSITE_VALUE;
                zeroes = sz;
                delta = v;
            }
        }
        // decode.swift:3118-3149: pass `zeroes` zero coefficients, refining every non-zero one on the way, and land on the next zero
        const uint64_t from = ~0ull << z, band = from & (band_hi < 64 ? ~(~0ull << band_hi) : ~0ull);
        uint64_t       Z = ~m & band;
        int            target = -1;
        if (zeroes < 64) {
            for (int i = 0; i < zeroes; ++i) Z &= Z - 1;
            if (Z) target = __ffsll((long long) Z) - 1;
        }
        const uint64_t passed = m & band & (target >= 0 ? ~(~0ull << target) : ~0ull);  // the non-zero coefficients in [z, target)
        const uint32_t k = (uint32_t) __popcll(passed);
        if (!(br.pos + k <= br.count)) return JPEG_SM100_ERR_TRUNCATED_ECS;
        if (APPLY) {
            uint64_t rest = passed;
            uint32_t w = 0, have = 0;
            while (rest) {
                if (have == 0) {
                    w = br.peek32();
                    have = 32;
                }
                const int zz = __ffsll((long long) rest) - 1;
                rest &= rest - 1;
                if (w & 0x80000000u) {
                    const int16_t c = blk[zz];
                    blk[zz] = (int16_t) (c + (int16_t) ((uint32_t) (c < 0 ? -1 : 1) << al));
                }
                w <<= 1;
                have -= 1;
                br.pos += 1;
            }
            if (target >= 0 && delta != 0) blk[target] = (int16_t) ((uint32_t) delta << al);
        } else {
            br.pos += k;
        }
        if (target < 0) break;  // the walk ran off the band: the block is complete
        z = target + 1;
    }
    return 0;
}

__global__ void __launch_bounds__(256) k_acr_masks(const __grid_constant__ ScanParams P, const uint32_t n_images, uint64_t *const masks)
{
    const uint32_t per = (uint32_t) P.ux[0] * (uint32_t) P.uy[0];
    const uint64_t total = (uint64_t) per * n_images;
    const uint64_t band = (~0ull << P.band_lo) & (P.band_hi < 64 ? ~(~0ull << P.band_hi) : ~0ull);
    for (uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t) gridDim.x * blockDim.x) {
        const uint32_t img = (uint32_t) (i / per), b = (uint32_t) (i - (uint64_t) img * per);
        const uint4   *src = reinterpret_cast<const uint4 *>(P.plane[0] + (size_t) img * P.image_stride[0] + (size_t) 64 * b);
        uint64_t       m = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const uint4    c = __ldg(src + j);
            const uint32_t w[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if (w[q] & 0xffffu) m |= 1ull << (8 * j + 2 * q);
                if (w[q] >> 16) m |= 1ull << (8 * j + 2 * q + 1);
            }
        }
        masks[i] = m & band;
    }
}

// The serial parse keeps its bit window in registers (the word for the next refill is loaded one refill ahead, so no load sits on the
// symbol-to-symbol chain) and reads symbols from the 8-bit fast tables (fast_entry_prog: one shared-memory load per symbol).
// (kept out of line: the padded bytes around the two ends of an interval are read a handful of times per scan)
__device__ __noinline__ uint32_t acr_edge_word(const uint32_t *w0, uint32_t i, uint32_t lead, uint32_t nbytes)
{
    uint32_t v = 0;
    for (int j = 0; j < 4; ++j) {
        const int64_t b = (int64_t) i * 4 + j - (int64_t) lead;  // byte offset inside the interval
        v = (v << 8) | ((b >= 0 && b < (int64_t) nbytes) ? (uint32_t) __ldg(reinterpret_cast<const uint8_t *>(w0) + (size_t) i * 4 + j) : 0xffu);
    }
    return v;
}
struct AcrWindow {  // absolute bit cursor + the three stream words around it; the 32 bits at the cursor are one funnel shift away
    const uint32_t *w0;  // the aligned word that holds the interval's first byte
    uint32_t        lead, nbytes;
    uint32_t        w_lo, w_n;  // words w_lo .. w_lo + w_n - 1 lie wholly inside the interval
    uint32_t        a, end;     // cursor and end of data, in bits from w0
    uint32_t        wi, hi, lo, nxt;  // words wi, wi + 1, wi + 2 (big-endian); wi == a >> 5
    __device__ __forceinline__ uint32_t word(uint32_t i) const  // big-endian word i of the 1-padded stream (jpeg.swift:1881-1887)
    {
        if (i - w_lo < w_n) return __byte_perm(__ldg(w0 + i), 0, 0x0123);
        return acr_edge_word(w0, i, lead, nbytes);
    }
    __device__ __forceinline__ void seek(uint32_t abs_bit)
    {
        a = abs_bit;
        wi = a >> 5;
        hi = word(wi), lo = word(wi + 1), nxt = word(wi + 2);
    }
    __device__ __forceinline__ void init(const uint8_t *base, uint32_t n)
    {
        lead = (uint32_t) (reinterpret_cast<uintptr_t>(base) & 3u);
        w0 = reinterpret_cast<const uint32_t *>(base - lead);
        nbytes = n;
        end = 8u * (lead + n);
        w_lo = lead ? 1u : 0u;
        const uint32_t w_hi = (lead + n) >> 2;  // first word that is not wholly inside
        w_n = w_hi > w_lo ? w_hi - w_lo : 0u;
        seek(8u * lead);
    }
    __device__ __forceinline__ uint32_t pos() const { return a - 8u * lead; }
    __device__ __forceinline__ uint32_t top() const { return __funnelshift_l(lo, hi, a); }
    __device__ __forceinline__ void advance(uint32_t n)
    {
        a += n;
        const uint32_t d = (a >> 5) - wi;
        if (d == 0u) return;
        if (d == 1u) {
            hi = lo, lo = nxt, wi += 1u;
            nxt = word(wi + 2u);  // (needed two words from now: the load is off the symbol-to-symbol chain)
            return;
        }
        seek(a);
    }
};

// One block of the parse on the fast tables; false: something the sequential kernel has to look at (truncation, an invalid code
// or value, a code that leans on the padding) -- the interval is flagged, never guessed.
// A symbol (run r, +-1) skips r ZERO coefficients and lands on the next one; every non-zero coefficient it passes owns one
// correction bit.  With sel[n] = position of the block's n-th zero (band order) and rho = zeros consumed so far, the landing
// position is sel[rho + r] and the bits to skip are (landing - z) - r: one shared-memory load and two subtractions per symbol.
// The warp builds sel[] for the block together (lane p ranks positions p and p + 32 with a popcount), then every lane walks
// the symbols redundantly (uniform addresses, broadcast loads).
__device__ __forceinline__ bool acr_parse_block(AcrWindow &br, const uint64_t m, int &skip, const int band_lo, const int band_hi, const uint32_t *tab,
                                                uint8_t *sel, const uint32_t lane)
{
    const uint64_t in_band = (~0ull << band_lo) & (band_hi < 64 ? ~(~0ull << band_hi) : ~0ull);
    const uint64_t Zm = ~m & in_band;
    __syncwarp();  // every lane is done with the table of two blocks ago (racecheck: blocks inside end-of-band runs do not come here)
    {
        const uint32_t lo = (uint32_t) Zm, hi = (uint32_t) (Zm >> 32);
        const uint32_t below = (1u << lane) - 1u;
        if ((lo >> lane) & 1u) sel[__popc(lo & below)] = (uint8_t) lane;
        if ((hi >> lane) & 1u) sel[__popc(lo) + __popc(hi & below)] = (uint8_t) (lane + 32u);
    }
    const int nz = __popcll(Zm);
    __syncwarp();
    int z = band_lo, rho = 0;
    while (z < band_hi) {
        const uint32_t top = br.top();
        uint32_t       ent = tab[top >> (32 - FAST_BITS)];
        if (ent & FAST_LINK) ent = tab[(ent >> 10) + (((top >> 16) & ((1u << (16 - FAST_BITS)) - 1u)) >> (ent & 7u))];
        const uint32_t len = ent & 0x7fu, size = (ent >> 8) & 0xffu, adv = (ent >> 16) & 0xffu, total = ent >> 24;
        if (ent == 0u) return false;
        int zeroes;
        if (adv & 0x80u) {  // EOBn: the rest of this block and of the next run - 1 blocks is correction bits
            zeroes = 64;
            skip = (int) ((1u << size) | (size ? (top << len) >> (32 - size) : 0u)) - 1;
        } else if (size == 0u) {
            zeroes = 15;  // ZRL
        } else if (size == 1u) {
            zeroes = (int) adv - 1;
        } else {
            return false;  // decode.swift:3099-3102: a refinement value must be -1, 0 or 1
        }
        const int n = rho + zeroes;
        if (n >= nz) {  // fewer zeros left than the symbol passes: the walk runs off the band, refining what is left
            const uint32_t k = total + (uint32_t) ((band_hi - z) - (nz - rho));
            if (!(br.a + k <= br.end)) return false;
            br.advance(k);
            break;
        }
        const int      target = (int) sel[n];
        const uint32_t k = total + (uint32_t) (target - z - zeroes);  // the symbol, then one correction bit per non-zero coefficient passed
        if (!(br.a + k <= br.end)) return false;
        br.advance(k);
        z = target + 1;
        rho = n + 1;
    }
    return true;
}

// one warp per (interval, image); every lane runs the same parse (uniform addresses), lane 0 writes.  rec[block] = first bit of the
// block inside its interval | ACR_EOB when the block lies inside an end-of-band run (then it holds correction bits only)
template <bool LUT_SMEM>
__global__ void __launch_bounds__(WARP) k_acr_parse(const __grid_constant__ ScanParams P, const uint64_t *const masks, uint32_t *const rec,
                                                    uint32_t *const flagged)
{
    extern __shared__ __align__(16) uint8_t smem[];
    __shared__ uint8_t s_sel[2][64];  // positions of the current block's zeros (two blocks: the next build never waits for this walk)
    const uint32_t   img = blockIdx.y, e = blockIdx.x, lane = threadIdx.x;
    const uint8_t   *lut_img = P.luts + (size_t) img * P.lut_stride;
    const LutHeader *hdr = reinterpret_cast<const LutHeader *>(smem);
    const uint16_t  *entries = reinterpret_cast<const uint16_t *>(lut_img + sizeof(LutHeader));
    {
        const uint32_t  total = reinterpret_cast<const LutHeader *>(lut_img)->total_all;  // reference entries + fast tables
        const uint32_t  words = (uint32_t) sizeof(LutHeader) / 4 + (LUT_SMEM ? (total + 1) / 2 : 0);
        uint32_t       *dst = reinterpret_cast<uint32_t *>(smem);
        const uint32_t *src = reinterpret_cast<const uint32_t *>(lut_img);
        for (uint32_t i = lane; i < words; i += WARP) dst[i] = src[i];
        __syncwarp();
        if (LUT_SMEM) entries = reinterpret_cast<const uint16_t *>(smem + sizeof(LutHeader));
    }
    const size_t slot = (size_t) img * P.n_ecs + e;
    int64_t      r0 = 0, r1 = P.H;
    if (P.interval != UINT64_MAX) {
        r0 = (int64_t) (((uint64_t) e * P.interval) / (uint32_t) P.W);
        r1 = (int64_t) (((uint64_t) (e + 1) * P.interval) / (uint32_t) P.W);
        if (r1 > P.H) r1 = P.H;
    }
    const uint64_t o0 = P.offsets[slot], o1 = P.offsets[slot + 1];
    // General.Range2's quirks (empty or inverted row ranges) and streams beyond 2^31 bits belong to the sequential kernel
    bool bad = !(r1 > r0) || (o1 - o0) > 0x0fffffffull;
    if (!bad) {
        const uint32_t per = (uint32_t) P.ux[0] * (uint32_t) P.uy[0];
        const uint32_t b0 = (uint32_t) r0 * (uint32_t) P.W, b1 = (uint32_t) r1 * (uint32_t) P.W;
        const uint64_t *mk = masks + (size_t) img * per;
        uint32_t       *rc = rec + (size_t) img * per;
        const int       ti = P.ac[0];
        int             skip = 0;
        uint32_t        b = b0;
        if (LUT_SMEM) {
            const uint32_t *tab = reinterpret_cast<const uint32_t *>(entries + hdr->fast[ti]);
            AcrWindow       br;
            br.init(P.ecs + o0, (uint32_t) (o1 - o0));
            uint64_t m_next = b < b1 ? __ldg(mk + b) : 0ull;
            while (b < b1 && !bad) {
                if (skip > 0) {
                    // inside an end-of-band run every block is its correction bits: 32 blocks per step
                    const uint32_t n = min(min((uint32_t) skip, b1 - b), (uint32_t) WARP);
                    const uint32_t k = lane < n ? (uint32_t) __popcll(__ldg(mk + b + lane)) : 0u;
                    uint32_t       inc = k;
#pragma unroll
                    for (int d = 1; d < WARP; d <<= 1) {
                        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
                        if ((int) lane >= d) inc += t;
                    }
                    if (lane < n) rc[b + lane] = (br.pos() + inc - k) | ACR_EOB;
                    const uint32_t sum = __shfl_sync(0xffffffffu, inc, WARP - 1);
                    if (!(br.a + sum <= br.end)) bad = true;
                    else br.advance(sum);
                    skip -= (int) n;
                    b += n;
                    m_next = b < b1 ? __ldg(mk + b) : 0ull;
                    continue;
                }
                if (lane == 0) rc[b] = br.pos();
                const uint64_t m = m_next;
                b += 1;
                m_next = b < b1 ? __ldg(mk + b) : 0ull;  // (the next block's map is on its way while this one is parsed)
                if (!acr_parse_block(br, m, skip, P.band_lo, P.band_hi, tab, s_sel[b & 1u], lane)) bad = true;
            }
        } else {
            const int      tn = hdr->n[ti], tz = hdr->zeta[ti];
            const uint32_t toff = hdr->offset[ti];
            AcrReader      br;
            br.base = P.ecs + o0;
            br.nbytes = (uint32_t) (o1 - o0);
            br.count = 8u * br.nbytes;
            br.pos = 0;
            while (b < b1 && !bad) {
                const uint32_t r = br.pos | (skip > 0 ? ACR_EOB : 0u);
                if (lane == 0) rc[b] = r;
                if (acr_block<false>(br, __ldg(mk + b), skip, P.band_lo, P.band_hi, P.al, entries, tn, tz, toff, nullptr) != 0) bad = true;
                b += 1;
            }
        }
    }
    if (lane == 0) {
        flagged[slot] = bad ? 1u : 0u;
        if (!bad) P.status[slot] = 0;
    }
}

template <bool LUT_SMEM>
__global__ void __launch_bounds__(128) k_acr_apply(const __grid_constant__ ScanParams P, const uint64_t *const masks, const uint32_t *const rec,
                                                   const uint32_t *const flagged)
{
    extern __shared__ __align__(16) uint8_t smem[];
    const uint32_t   img = blockIdx.z, e = blockIdx.y;
    const uint8_t   *lut_img = P.luts + (size_t) img * P.lut_stride;
    const LutHeader *hdr = reinterpret_cast<const LutHeader *>(smem);
    const uint16_t  *entries = reinterpret_cast<const uint16_t *>(lut_img + sizeof(LutHeader));
    const size_t     slot = (size_t) img * P.n_ecs + e;
    if (flagged[slot]) return;
    {
        const uint32_t  total = reinterpret_cast<const LutHeader *>(lut_img)->total_entries;
        const uint32_t  words = (uint32_t) sizeof(LutHeader) / 4 + (LUT_SMEM ? (total + 1) / 2 : 0);
        uint32_t       *dst = reinterpret_cast<uint32_t *>(smem);
        const uint32_t *src = reinterpret_cast<const uint32_t *>(lut_img);
        for (uint32_t i = threadIdx.x; i < words; i += blockDim.x) dst[i] = src[i];
        __syncthreads();
        if (LUT_SMEM) entries = reinterpret_cast<const uint16_t *>(smem + sizeof(LutHeader));
    }
    int64_t r0 = 0, r1 = P.H;
    if (P.interval != UINT64_MAX) {
        r0 = (int64_t) (((uint64_t) e * P.interval) / (uint32_t) P.W);
        r1 = (int64_t) (((uint64_t) (e + 1) * P.interval) / (uint32_t) P.W);
        if (r1 > P.H) r1 = P.H;
    }
    const uint32_t per = (uint32_t) P.ux[0] * (uint32_t) P.uy[0];
    const uint32_t b0 = (uint32_t) r0 * (uint32_t) P.W, b1 = (uint32_t) r1 * (uint32_t) P.W;
    const uint64_t o0 = P.offsets[slot], o1 = P.offsets[slot + 1];
    const int      ti = P.ac[0];
    const int      tn = hdr->n[ti], tz = hdr->zeta[ti];
    const uint32_t toff = hdr->offset[ti];
    for (uint32_t b = b0 + blockIdx.x * blockDim.x + threadIdx.x; b < b1; b += gridDim.x * blockDim.x) {
        const uint32_t r = __ldg(rec + (size_t) img * per + b);
        AcrReader      br;
        br.base = P.ecs + o0;
        br.nbytes = (uint32_t) (o1 - o0);
        br.count = 8u * br.nbytes;
        br.pos = r & ~ACR_EOB;
        int skip = (r & ACR_EOB) ? 1 : 0;  // (the run's length beyond this block is the next blocks' business)
        acr_block<true>(br, __ldg(masks + (size_t) img * per + b), skip, P.band_lo, P.band_hi, P.al, entries, tn, tz, toff,
                        P.plane[0] + (size_t) img * P.image_stride[0] + (size_t) 64 * b);
    }
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

// clear MCU rows [row0, H) of every plane of the scan, all images (cudaMemset2D: one row of the 2-D set per image)
int zero_plane_rows(jpeg_sm100_ctx *ctx, const ScanParams &P, int n_comp, uint32_t n_images, int row0)
{
    for (int c = 0; c < n_comp; ++c) {
        if (!P.plane[c]) continue;
        const int y0 = row0 * P.fy[c] < P.uy[c] ? row0 * P.fy[c] : P.uy[c];
        const size_t bytes = (size_t) 128 * P.ux[c] * (P.uy[c] - y0);
        if (!bytes) continue;
        CU_TRY(ctx, cudaMemset2DAsync(P.plane[c] + (size_t) 64 * P.ux[c] * y0, P.image_stride[c] * 2, 0, bytes, n_images, ctx->stream));
    }
    return JPEG_SM100_OK;
}

// The three-phase AC refinement path (k_acr_*) pays when intervals are few and long; with many short intervals the
// one-thread-per-interval kernel already fills the GPU.  JPEG_SM100_ACR=0 / 1 forces it off / on (A/B validation).
bool acr_wanted(jpeg_sm100_ctx *ctx, const ScanParams &P, uint32_t n_images, uint32_t n_ecs, uint64_t interval)
{
    const char *seq = getenv("JPEG_SM100_HUFF");
    if (seq && strcmp(seq, "seq") == 0) return false;
    if (n_ecs > 65535u || n_images > 65535u || !P.plane[0]) return false;
    if ((uint64_t) P.ux[0] * (uint64_t) P.uy[0] > 0x7fffffffull) return false;
    const char *e = getenv("JPEG_SM100_ACR");
    if (e) return atoi(e) != 0;
    const uint64_t rows = (interval == JPEG_SM100_INTERVAL_NONE) ? (uint64_t) P.H : (interval + P.W - 1) / P.W;
    return rows * (uint64_t) P.W >= 1024 && (uint64_t) n_images * n_ecs <= (uint64_t) ctx->sm_count * 8;
}

}  // namespace

// decode.swift:2884-2895, 3186-3203 + 310-351: the error a sequential scan raises for this table set before it reads a bit
// (0: none).  Used by the batch entry points to give every image its own status.
int jpeg_huffman_validate_tables(const jpeg_sm100_scan_desc *scan, const jpeg_sm100_huff_table *set8)
{
    for (int c = 0; c < scan->n_comp; ++c)
        for (int k = 0; k < 2; ++k) {
            const int sel = k == 0 ? scan->comp[c].dc : scan->comp[c].ac;
            if (sel < 0 || sel > 3) return JPEG_SM100_ERR_INVALID_ARGUMENT;
            const jpeg_sm100_huff_table &t = set8[4 * k + sel];
            if (!t.present) return k == 0 ? JPEG_SM100_ERR_UNDEFINED_DC : JPEG_SM100_ERR_UNDEFINED_AC;
            int n, z, leaves = 0;
            if (!huff_size(t.counts, n, z)) return JPEG_SM100_ERR_INVALID_HUFFMAN;
            for (int l = 0; l < 16; ++l) leaves += t.counts[l];
            if (leaves > 256) return JPEG_SM100_ERR_INVALID_HUFFMAN;
        }
    return JPEG_SM100_OK;
}

// Layer-B implementation.  scratch slots 8 (LUTs) and 9 (per-ECS status) belong to this file.
static int decode_scan_impl(jpeg_sm100_ctx *ctx, const jpeg_sm100_scan_desc *scan, const uint8_t *d_ecs,
                            const uint64_t *d_offsets, uint32_t n_ecs, uint64_t interval, int extend,
                            const jpeg_sm100_huff_table *tables, int tables_shared,
                            const jpeg_sm100_dev_spectral *sp, int32_t *d_status);

// JPEG_SM100_TRACE=1: every scan is bracketed by stream synchronisations and its wall time goes to stderr (a diagnostic: it
// serialises the stream, never set it for a measurement of anything else)
int jpeg_huffman_decode_scan(jpeg_sm100_ctx *ctx, const jpeg_sm100_scan_desc *scan, const uint8_t *d_ecs,
                             const uint64_t *d_offsets, uint32_t n_ecs, uint64_t interval, int extend,
                             const jpeg_sm100_huff_table *tables, int tables_shared,
                             const jpeg_sm100_dev_spectral *sp, int32_t *d_status)
{
    static const bool trace = getenv("JPEG_SM100_TRACE") != nullptr;
    if ((extend & JPEG_SM100_SCAN_T81) && jpeg_virtual_scan_needed(scan, sp, interval)) {
        // ITU-T T.81 placement of an interval that is not whole MCU rows: the same scan on a virtual MCU grid (remap.cu)
        JpegVirtualScan v;
        J_TRY(jpeg_virtual_scan_setup(ctx, scan, sp, interval, &v));
        if (!(extend & JPEG_SM100_SCAN_FRESH)) J_TRY(jpeg_virtual_scan_copy(ctx, &v, true));
        J_TRY(decode_scan_impl(ctx, &v.scan, d_ecs, d_offsets, n_ecs, interval, extend & ~JPEG_SM100_SCAN_T81, tables, tables_shared, &v.sp, d_status));
        return jpeg_virtual_scan_copy(ctx, &v, false);
    }
    extend &= ~JPEG_SM100_SCAN_T81;
    if (!trace) return decode_scan_impl(ctx, scan, d_ecs, d_offsets, n_ecs, interval, extend, tables, tables_shared, sp, d_status);
    cudaStreamSynchronize(ctx->stream);
    timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    const uint64_t l0 = ctx->launches;
    const int      r = decode_scan_impl(ctx, scan, d_ecs, d_offsets, n_ecs, interval, extend, tables, tables_shared, sp, d_status);
    cudaStreamSynchronize(ctx->stream);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    fprintf(stderr, "[jpeg_sm100] scan band %d..%d bits %d/%d comps %d images %u intervals %u extend %d: %.3f ms, %llu launches, rc %d\n",
            scan ? scan->band_lo : -1, scan ? scan->band_hi : -1, scan ? scan->bit_hi : 0, scan ? scan->bit_lo : 0, scan ? scan->n_comp : 0,
            sp ? sp->n_images : 0u, n_ecs, extend, (t1.tv_sec - t0.tv_sec) * 1e3 + (t1.tv_nsec - t0.tv_nsec) * 1e-6,
            (unsigned long long) (ctx->launches - l0), r);
    return r;
}

static int decode_scan_impl(jpeg_sm100_ctx *ctx, const jpeg_sm100_scan_desc *scan, const uint8_t *d_ecs,
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
    const bool fresh = (extend & JPEG_SM100_SCAN_FRESH) != 0;  // planes are to be treated as newly created (all zero)
    P.extend = ((extend & JPEG_SM100_SCAN_EXTEND) && P.kind <= 1) ? 1 : 0;
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
    size_t  max_entries = 0, max_fast = 0;  // uint16 units: whole LUT set / its fast tables only
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
                uint8_t depth[FAST_ENTRIES];
                h.fast[ti] = total;
                total += 2 * (FAST_ENTRIES + sub_depths(raw[s].counts[ti], depth));  // 32-bit entries: fast table + sub-tables
            }
        h.total_all = total;
        if (total > max_entries) max_entries = total;
        if (total - h.total_entries > max_fast) max_fast = total - h.total_entries;
    }
    const size_t entry_bytes = (max_entries * 2 + 15) & ~size_t(15);
    const size_t stride = sizeof(LutHeader) + entry_bytes + 16;
    void *d_raw = nullptr, *d_luts = nullptr;
    J_TRY(scratch_reserve(ctx, 7, sizeof(RawSet) * (size_t) n_sets, &d_raw));
    J_TRY(scratch_reserve(ctx, 8, stride * n_sets, &d_luts));
    CU_TRY(ctx, cudaMemcpyAsync(d_raw, raw, sizeof(RawSet) * (size_t) n_sets, cudaMemcpyHostToDevice, ctx->stream));
    J_TRY(pinned_release(ctx, slot));
    k_build_luts<<<dim3(8, n_sets), 128, 0, ctx->stream>>>(reinterpret_cast<const RawSet *>(d_raw),
                                                           reinterpret_cast<uint8_t *>(d_luts), stride, (P.kind == 3 || P.kind == 4) ? 1 : 0);
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
        if (((P.plane[c] - plane0) & 63) || (P.image_stride[c] & 63) || span / 64 > 0x7ffffff0ull) fast_ok = false;
    }
    if (!plane0) fast_ok = false;
    if (P.kind <= 1 && !use_flat && fast_ok) {
        const size_t smem2 = sizeof(LutHeader) + 12 * sizeof(BlkInfo) + 4 * WARP * sizeof(int) + entry_bytes;
        static const bool no_par = [] {
            const char *e = getenv("JPEG_SM100_HUFF");
            return e && strcmp(e, "seq") == 0;  // one thread per interval only (A/B validation of the parallel decoder)
        }();
        bool par_done = false;
        // DC-first scans cut into many SHORT intervals (config #4: 5.7 Kbit each): one thread per interval is faster than sixteen
        // (measured 1.04 vs 1.19 ms for 256 x 1080p) -- the block-in-MCU index never re-synchronises, so the sixteen subsequences of
        // an interval are repaired one round at a time anyway.  (JPEG_SM100_PAR_T set: tests force the parallel decoder.)
        const bool dc_short = P.kind == 1 && interval != JPEG_SM100_INTERVAL_NONE && !getenv("JPEG_SM100_PAR_T") &&
                              (ctx->hint_interval_bytes ? 8 * ctx->hint_interval_bytes
                                                        : 8 * ((interval + P.W - 1) / P.W) * (uint64_t) P.W * (uint64_t) volume) < 16384;
        if (!no_par && !dc_short) {
            // subsequence-parallel decode (+ fused row clearing and DC prefix sums), then the sequential kernel for whatever it flagged.
            // `extend` (the first scan of a file, decode.swift:3214-3236): the planes are already sized, so the flag only means "rows
            // stop silently where the data ends"; the parallel pass flags an interval that runs dry (or shows sixteen 1-bits at the
            // start of a row) and k_decode_fast(only_flagged) reproduces the silent stop.
            const bool     dc_first = P.kind == 1;
            const uint64_t rows_max = (interval == JPEG_SM100_INTERVAL_NONE) ? (uint64_t) P.H : (interval + P.W - 1) / P.W + 1;
            const uint64_t dc_per_interval = rows_max * (uint64_t) P.W * (uint64_t) volume;
            const uint64_t slots = (uint64_t) n_images * n_ecs;
            void          *d_dc = nullptr, *d_flag = nullptr;
            if (dc_per_interval <= 0x7fffffffull && slots * dc_per_interval * 2 <= (4ull << 30)) {
                J_TRY(scratch_reserve(ctx, 12, (size_t) (slots * dc_per_interval * 2 + 256), &d_dc));
                J_TRY(scratch_reserve(ctx, 13, (size_t) (slots * 4 + 256), &d_flag));
                const size_t smem_par = PAR_PRE + ((max_fast * 2 + 15) & ~size_t(15));
                // Threads per interval (16 .. 128).  Every subsequence is parsed ~(2 + warm-up / length) times, so long
                // subsequences (>= 3 Kbit) waste the least work; small batches take more threads per interval to fill the GPU.
                // tuning / test overrides, read per call: log2(threads per interval), warm-up bits
                const char *env_ts = getenv("JPEG_SM100_PAR_T"), *env_ws = getenv("JPEG_SM100_PAR_WARM");
                const int   env_t = env_ts ? atoi(env_ts) : 0, env_warm = env_ws ? atoi(env_ws) : 0;
                const uint64_t rows_typ = (interval == JPEG_SM100_INTERVAL_NONE) ? (uint64_t) P.H : (interval + P.W - 1) / P.W;
                const uint64_t est_bits = ctx->hint_interval_bytes ? 8 * ctx->hint_interval_bytes
                                                                   : (dc_first ? 8 : 96) * rows_typ * (uint64_t) P.W * (uint64_t) volume;
                // few, large intervals (a file without DRI is ONE interval): a 512-thread CTA per interval
                const bool big = (est_bits >> 7) >= (dc_first ? 2048 : 16384) && slots * 128 < (uint64_t) ctx->sm_count * 1024;
                const int  nt = big ? PAR_BIG_THREADS : PAR_THREADS, tmax = big ? 9 : 7;
                // (measured on 64 x 4K, 134 Kbit per interval: 16 threads 1.89 ms, 32 threads 2.00 ms; warm-up 2048 bits best)
                int        tshift = tmax;
                while (tshift > 4 && (est_bits >> tshift) < 8192) --tshift;
                while (tshift < tmax && (slots << tshift) < (uint64_t) ctx->sm_count * 512 && (est_bits >> (tshift + 1)) >= (uint64_t) PAR_MIN_BITS) ++tshift;
                if (env_t >= 4 && env_t <= tmax) tshift = env_t;
                const uint32_t warm_bits = env_ws ? (uint32_t) (env_warm > 0 ? env_warm : 0) : (dc_first ? (tshift >= 6 ? 768u : 256u) : 2048u);
                const uint32_t G = (uint32_t) nt >> tshift;
                const dim3     grid_par((n_ecs + G - 1) / G, n_images);
                const size_t   smem_total = smem_par + (dc_first ? 16 + (size_t) nt * 12 * 8 : (size_t) nt * (PAR_BUF_STRIDE + 4 * PAR_RING) + 32 + (PAR_SWIZZLE ? 112 : 0));
                if (!ctx->par_smem_set) {  // same bound from every ctx of the process: LUTs (< 48 KB) + stage (<= 96 KB) + block buffers
                    CU_TRY(ctx, cudaFuncSetAttribute(k_decode_par<PAR_THREADS, PAR_MIN_CTAS, MODE_SEQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
                    CU_TRY(ctx, cudaFuncSetAttribute(k_decode_par<PAR_BIG_THREADS, 1, MODE_SEQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
                    CU_TRY(ctx, cudaFuncSetAttribute(k_decode_par<PAR_THREADS, PAR_MIN_CTAS, MODE_DC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
                    CU_TRY(ctx, cudaFuncSetAttribute(k_decode_par<PAR_BIG_THREADS, 1, MODE_DC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
                    CU_TRY(ctx, cudaFuncSetAttribute(k_decode_par_cluster, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
                    ctx->par_smem_set = 200 * 1024;
                }
                static const bool want_stats = getenv("JPEG_SM100_PAR_STATS") != nullptr;
                uint32_t         *d_stats = nullptr;
                if (want_stats) {
                    void *p = nullptr;
                    J_TRY(scratch_reserve(ctx, 14, 128, &p));
                    d_stats = reinterpret_cast<uint32_t *>(p);
                    CU_TRY(ctx, cudaMemsetAsync(d_stats, 0, 128, ctx->stream));
                }
                if (dc_first) {
                    // a DC-first scan defines coefficient 0 only: fresh planes are cleared as a whole, the rest of a block is not touched
                    if (fresh) J_TRY(zero_plane_rows(ctx, P, scan->n_comp, n_images, 0));
                } else if (fresh && interval != JPEG_SM100_INTERVAL_NONE && ((uint64_t) n_ecs * interval) / (uint64_t) P.W < (uint64_t) P.H)
                    // rows no interval reaches (a file with too few intervals) stay as a fresh plane has them: zero
                    J_TRY(zero_plane_rows(ctx, P, scan->n_comp, n_images, (int) (((uint64_t) n_ecs * interval) / (uint64_t) P.W)));
                // few, large intervals: a cluster of CTAs per interval (JPEG_SM100_PAR_CLUSTER=0: one 512-thread CTA, kept for A/B)
                const char *env_cl = getenv("JPEG_SM100_PAR_CLUSTER");
                uint32_t    csize = (big && !dc_first) ? (env_cl ? (uint32_t) atoi(env_cl) : 8u) : 0u;
                while (csize > 1 && (est_bits / (CL_THREADS * csize)) < 2048) csize >>= 1;
                if (csize > 8 || (csize & (csize - 1))) csize = 8;
                if (dc_first) {
                    if (big)
                        k_decode_par<PAR_BIG_THREADS, 1, MODE_DC><<<grid_par, PAR_BIG_THREADS, smem_total, ctx->stream>>>(
                            P, plane0, reinterpret_cast<int16_t *>(d_dc), (uint32_t) dc_per_interval, reinterpret_cast<uint32_t *>(d_flag),
                            d_stats, tshift, warm_bits, (uint32_t) smem_par);
                    else
                        k_decode_par<PAR_THREADS, PAR_MIN_CTAS, MODE_DC><<<grid_par, PAR_THREADS, smem_total, ctx->stream>>>(
                            P, plane0, reinterpret_cast<int16_t *>(d_dc), (uint32_t) dc_per_interval, reinterpret_cast<uint32_t *>(d_flag),
                            d_stats, tshift, warm_bits, (uint32_t) smem_par);
                } else if (big && csize >= 1 && !(env_cl && atoi(env_cl) == 0)) {
                    cudaLaunchConfig_t cfg;
                    memset(&cfg, 0, sizeof cfg);
                    cfg.gridDim = dim3(csize * n_ecs, n_images);
                    cfg.blockDim = dim3(CL_THREADS);
                    cfg.dynamicSmemBytes = smem_par + (size_t) CL_THREADS * (PAR_BUF_STRIDE + 4 * PAR_RING) + 32 + (PAR_SWIZZLE ? 112 : 0);
                    cfg.stream = ctx->stream;
                    cudaLaunchAttribute attr[1];
                    attr[0].id = cudaLaunchAttributeClusterDimension;
                    attr[0].val.clusterDim.x = csize, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = 1;
                    cfg.attrs = attr, cfg.numAttrs = 1;
                    CU_TRY(ctx, cudaLaunchKernelEx(&cfg, k_decode_par_cluster, P, plane0, reinterpret_cast<int16_t *>(d_dc),
                                                   (uint32_t) dc_per_interval, reinterpret_cast<uint32_t *>(d_flag), warm_bits, (uint32_t) smem_par));
                } else if (big)
                    k_decode_par<PAR_BIG_THREADS, 1, MODE_SEQ><<<grid_par, PAR_BIG_THREADS, smem_total, ctx->stream>>>(
                        P, plane0, reinterpret_cast<int16_t *>(d_dc), (uint32_t) dc_per_interval, reinterpret_cast<uint32_t *>(d_flag),
                        d_stats, tshift, warm_bits, (uint32_t) smem_par);
                else
                    k_decode_par<PAR_THREADS, PAR_MIN_CTAS, MODE_SEQ><<<grid_par, PAR_THREADS, smem_total, ctx->stream>>>(
                        P, plane0, reinterpret_cast<int16_t *>(d_dc), (uint32_t) dc_per_interval, reinterpret_cast<uint32_t *>(d_flag),
                        d_stats, tshift, warm_bits, (uint32_t) smem_par);
                LAUNCH_CHECK(ctx);
                if (want_stats) {
                    uint32_t h[32];
                    CU_TRY(ctx, cudaMemcpyAsync(h, d_stats, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
                    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
                    {
                        const unsigned long long *ph = reinterpret_cast<const unsigned long long *>(h) + 4;
                        const double ctas = (double) grid_par.x * grid_par.y;
                        fprintf(stderr, "[k_decode_par] cycles per CTA: stage %.0f, round0 %.0f, rounds %.0f, zero+scan %.0f, final %.0f, dc %.0f\n",
                                ph[0] / ctas, ph[1] / ctas, ph[2] / ctas, ph[3] / ctas, ph[4] / ctas, ph[5] / ctas);
                    }
                    fprintf(stderr, "[k_decode_par] flagged intervals %u (block total off: %u); subsequences: broken %u, count off %u\n", h[22], h[23], h[20], h[21]);
                    fprintf(stderr, "[k_decode_par] T %d, warm %u, intervals %u, subsequences/interval %.1f, rounds avg %.2f max %u, re-parses per subsequence %.2f\n",
                            1 << tshift, warm_bits, h[3], (double) h[2] / h[3], (double) h[1] / h[3], h[4], (double) h[0] / h[2]);
                }
                if (!dc_first) {
                    k_zero_flagged<<<dim3(n_ecs, n_images), 128, 0, ctx->stream>>>(P, reinterpret_cast<const uint32_t *>(d_flag));
                    LAUNCH_CHECK(ctx);
                }
                k_decode_fast<<<grid, WARP, smem2, ctx->stream>>>(P, plane0, reinterpret_cast<const uint32_t *>(d_flag));
                par_done = true;
            }
        }
        if (!par_done) {
            if (fresh) J_TRY(zero_plane_rows(ctx, P, scan->n_comp, n_images, 0));
            k_decode_fast<<<grid, WARP, smem2, ctx->stream>>>(P, plane0, nullptr);
        }
    } else if (P.kind <= 1) {
        if (fresh) J_TRY(zero_plane_rows(ctx, P, scan->n_comp, n_images, 0));
        if (P.lut_smem) k_decode_flat<true><<<grid, WARP, smem, ctx->stream>>>(P);
        else k_decode_flat<false><<<grid, WARP, smem, ctx->stream>>>(P);
    } else if (P.kind == 3 && fast_ok && !(getenv("JPEG_SM100_HUFF") && strcmp(getenv("JPEG_SM100_HUFF"), "seq") == 0) &&
               (uint64_t) n_images * n_ecs * 4 <= (1ull << 30) && interval != 0 &&
               (interval == JPEG_SM100_INTERVAL_NONE || interval % (uint64_t) P.W == 0)) {
        // progressive AC-first scan: the subsequence-parallel decoder (par_run_ac), then the sequential one for flagged intervals
        if (fresh) J_TRY(zero_plane_rows(ctx, P, scan->n_comp, n_images, 0));
        const uint64_t slots = (uint64_t) n_images * n_ecs;
        void          *d_flag = nullptr;
        J_TRY(scratch_reserve(ctx, 13, (size_t) (slots * 4 + 256), &d_flag));
        const uint64_t rows_typ = (interval == JPEG_SM100_INTERVAL_NONE) ? (uint64_t) P.H : (interval + P.W - 1) / P.W;
        const uint64_t est_bits = ctx->hint_interval_bytes ? 8 * ctx->hint_interval_bytes : 96 * rows_typ * (uint64_t) P.W;
        const char    *env_ts = getenv("JPEG_SM100_PAR_T"), *env_ws = getenv("JPEG_SM100_PAR_WARM");
        int            tshift = 7;
        while (tshift > 4 && (est_bits >> tshift) < 4096) --tshift;
        while (tshift < 7 && (slots << tshift) < (uint64_t) ctx->sm_count * 1024 && (est_bits >> (tshift + 1)) >= (uint64_t) PAR_MIN_BITS) ++tshift;
        if (env_ts && atoi(env_ts) >= 4 && atoi(env_ts) <= 7) tshift = atoi(env_ts);
        const uint32_t warm_bits = env_ws ? (uint32_t) (atoi(env_ws) > 0 ? atoi(env_ws) : 0) : 512u;
        const uint32_t G = (uint32_t) PAR_THREADS >> tshift;
        const size_t   smem_par = PAR_PRE + ((max_fast * 2 + 15) & ~size_t(15));
        if (!ctx->par_smem_ac) {
            CU_TRY(ctx, cudaFuncSetAttribute(k_decode_par<PAR_THREADS, PAR_MIN_CTAS, MODE_AC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            ctx->par_smem_ac = 1;
        }
        // (the side array / block buffers of the sequential-scan variant are unused: dc_per_interval only bounds N_total)
        k_decode_par<PAR_THREADS, PAR_MIN_CTAS, MODE_AC><<<dim3((n_ecs + G - 1) / G, n_images), PAR_THREADS, smem_par, ctx->stream>>>(
            P, plane0, nullptr, 0xffffffffu, reinterpret_cast<uint32_t *>(d_flag), nullptr, tshift, warm_bits, (uint32_t) smem_par);
        LAUNCH_CHECK(ctx);
        k_zero_band_flagged<<<dim3(n_ecs, n_images), 128, 0, ctx->stream>>>(P, reinterpret_cast<const uint32_t *>(d_flag));
        LAUNCH_CHECK(ctx);
        if (P.lut_smem) k_decode_progressive<true><<<grid, WARP, smem, ctx->stream>>>(P, reinterpret_cast<const uint32_t *>(d_flag));
        else k_decode_progressive<false><<<grid, WARP, smem, ctx->stream>>>(P, reinterpret_cast<const uint32_t *>(d_flag));
    } else if (P.kind == 2 && !(getenv("JPEG_SM100_HUFF") && strcmp(getenv("JPEG_SM100_HUFF"), "seq") == 0) &&
               (uint64_t) n_images * n_ecs * 4 <= (1ull << 30)) {
        // progressive DC refinement: a bit gather, one thread per block; intervals with Range2 quirks go to the sequential kernel
        if (fresh) J_TRY(zero_plane_rows(ctx, P, scan->n_comp, n_images, 0));
        void *d_flag = nullptr;
        J_TRY(scratch_reserve(ctx, 13, (size_t) ((uint64_t) n_images * n_ecs * 4 + 256), &d_flag));
        const uint64_t rows_typ = (interval == JPEG_SM100_INTERVAL_NONE) ? (uint64_t) P.H : (interval + P.W - 1) / P.W;
        const uint64_t per = rows_typ * (uint64_t) P.W * (uint64_t) volume;
        const dim3     g2((uint32_t) std::min<uint64_t>((per + 255) / 256 ? (per + 255) / 256 : 1, 1024), std::min<uint32_t>(n_ecs, 65535u),
                          std::min<uint32_t>(n_images, 65535u));
        k_decode_dc_refine<<<g2, 256, 0, ctx->stream>>>(P, n_images, reinterpret_cast<uint32_t *>(d_flag));
        LAUNCH_CHECK(ctx);
        if (P.lut_smem) k_decode_progressive<true><<<grid, WARP, smem, ctx->stream>>>(P, reinterpret_cast<const uint32_t *>(d_flag));
        else k_decode_progressive<false><<<grid, WARP, smem, ctx->stream>>>(P, reinterpret_cast<const uint32_t *>(d_flag));
    } else if (P.kind == 4 && acr_wanted(ctx, P, n_images, n_ecs, interval)) {
        // progressive AC refinement over few, large intervals: non-zero maps, a serial parse per interval (one warp), a parallel apply
        if (fresh) J_TRY(zero_plane_rows(ctx, P, scan->n_comp, n_images, 0));
        const uint64_t per = (uint64_t) P.ux[0] * (uint64_t) P.uy[0], blocks = per * n_images;
        void          *d_work = nullptr, *d_flag = nullptr;
        J_TRY(scratch_reserve(ctx, 12, (size_t) (blocks * 12 + 256), &d_work));
        J_TRY(scratch_reserve(ctx, 13, (size_t) ((uint64_t) n_images * n_ecs * 4 + 256), &d_flag));
        uint64_t *d_masks = reinterpret_cast<uint64_t *>(d_work);
        uint32_t *d_rec = reinterpret_cast<uint32_t *>(d_masks + blocks);
        uint32_t *d_fl = reinterpret_cast<uint32_t *>(d_flag);
        k_acr_masks<<<(uint32_t) std::min<uint64_t>((blocks + 255) / 256, (uint64_t) ctx->sm_count * 16), 256, 0, ctx->stream>>>(P, n_images, d_masks);
        LAUNCH_CHECK(ctx);
        if (P.lut_smem) k_acr_parse<true><<<dim3(n_ecs, n_images), WARP, smem, ctx->stream>>>(P, d_masks, d_rec, d_fl);
        else k_acr_parse<false><<<dim3(n_ecs, n_images), WARP, smem, ctx->stream>>>(P, d_masks, d_rec, d_fl);
        LAUNCH_CHECK(ctx);
        const uint64_t rows_typ = (interval == JPEG_SM100_INTERVAL_NONE) ? (uint64_t) P.H : (interval + P.W - 1) / P.W;
        const dim3     ga((uint32_t) std::min<uint64_t>((rows_typ * (uint64_t) P.W + 127) / 128, 4096), n_ecs, n_images);
        if (P.lut_smem) k_acr_apply<true><<<ga, 128, smem, ctx->stream>>>(P, d_masks, d_rec, d_fl);
        else k_acr_apply<false><<<ga, 128, smem, ctx->stream>>>(P, d_masks, d_rec, d_fl);
        LAUNCH_CHECK(ctx);
        if (P.lut_smem) k_decode_progressive<true><<<grid, WARP, smem, ctx->stream>>>(P, d_fl);
        else k_decode_progressive<false><<<grid, WARP, smem, ctx->stream>>>(P, d_fl);
    } else {
        if (fresh) J_TRY(zero_plane_rows(ctx, P, scan->n_comp, n_images, 0));
        if (P.lut_smem) k_decode_progressive<true><<<grid, WARP, smem, ctx->stream>>>(P, nullptr);
        else k_decode_progressive<false><<<grid, WARP, smem, ctx->stream>>>(P, nullptr);
    }
    LAUNCH_CHECK(ctx);
    if (d_status) {
        k_reduce_status<<<n_images, 128, 0, ctx->stream>>>(P.status, n_ecs, d_status);
        LAUNCH_CHECK(ctx);
    }
    return JPEG_SM100_OK;
}

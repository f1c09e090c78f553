// lexer.cu -- N1 (SURVEY §8 "next"): the per-scan half of the container lexer on the GPU.
//
// Replaces, for the bytes between an SOS header and the next non-RSTn marker, what Bytestream.segment(prefix: true)
// (decode.swift:130-190) does one byte at a time inside the scan loop of Context.decompress (decode.swift:3895-3933):
//   * byte unstuffing   FF 00 -> FF
//   * fill bytes        FF FF .. FF x  ==  FF x
//   * RSTn splitting    FF Dn ends an entropy-coded segment; n must equal (index mod 8), else
//                       DecodingError.invalidRestartPhase (decode.swift:3929-3932)
// and produces exactly the inputs jpeg_sm100_dev_decode_scan consumes: the unstuffed bytes of all images back to back
// and n_images * n_ecs + 1 byte offsets.
//
// The lexer's automaton has one bit of state ("the previous byte was FF"), and that bit is a function of the previous
// byte alone:   state(i) = (raw[i-1] == FF).   So every byte can be classified independently,
//   emits(i)  = (prev != FF && cur != FF) || (prev == FF && cur == 00)        value = prev == FF ? FF : cur
//   split(i)  =  prev == FF && (cur & F8) == D0
//   foreign(i)=  prev == FF && cur not in {00, FF, D0..D7}                     (a marker the host should have cut at)
// and the whole thing is a stream compaction: count (k_lex_count) -> scan (k_lex_scan) -> scatter (k_lex_scatter).
// HBM-bound: 2 reads + 1 write of the scan bytes; 16 bytes per thread with SIMD-in-a-register byte compares (VSETxx),
// output tiles staged in shared memory and stored as aligned 16-byte vectors.
#include "common.cuh"

namespace {

constexpr int LEX_THREADS = 256;
constexpr int LEX_CHUNK = 16;                        // raw bytes per thread
constexpr int LEX_TILE = LEX_THREADS * LEX_CHUNK;    // 4 KB of raw bytes per tile
constexpr int LEX_COUNT_TILES = 4;                   // tiles per CTA of the count pass

struct LexImage {
    uint64_t raw_off, raw_len;
};

struct Chunk {
    uint32_t emit[4];   // bit 7 of every byte that is emitted
    uint32_t split[4];  // bit 7 of every byte that ends an entropy-coded segment (the n of an RSTn)
    uint32_t out[4];    // output byte values
    uint32_t cur[4];    // raw bytes (masked to the image)
    uint32_t foreign;   // any foreign marker
    bool     plain;     // no FF in or right before the chunk, all 16 bytes inside the image: everything is emitted as it is
};

__device__ __forceinline__ uint32_t bytes_below(int n)  // 0xFF in bytes j < n
{
    return n <= 0 ? 0u : (n >= 4 ? 0xFFFFFFFFu : ((1u << (8 * n)) - 1u));
}
// SIMD-in-a-register byte compare, exact: bit 7 of every byte of v that is zero (three integer instructions; the SIMD video
// instructions the first generation used -- __vcmpeq4 -- are emulated on sm_100: the count pass spent 236 and the scatter pass
// 518 instructions per 16 bytes on them, which made both passes issue-bound at a quarter of the HBM rate)
__device__ __forceinline__ uint32_t zero_bytes(uint32_t v) { return ~(((v & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | v | 0x7F7F7F7Fu); }

// classify the 16 bytes at raw[base .. base+16); only bytes in [lo, hi) belong to the image
__device__ __forceinline__ Chunk classify(const uint8_t *__restrict__ raw, uint64_t base, uint64_t lo, uint64_t hi)
{
    Chunk c;
    uint4 v = __ldg(reinterpret_cast<const uint4 *>(raw + base));
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    const uint32_t before = (base > lo) ? (uint32_t) __ldg(raw + base - 1) : 0u;
    const int s = lo > base ? (int) (lo - base) : 0;
    const int e = hi - base >= 16 ? 16 : (int) (hi - base);
    const bool full = s == 0 && e == 16;
    c.foreign = 0;
    // entropy-coded data holds an FF every few hundred bytes: most chunks have none, and then every byte is emitted as it is
    // ((~w - 0x01..01) & w & 0x80..80 is non-zero iff some byte of w is FF)
    const uint32_t any_ff = (((~w[0] - 0x01010101u) & w[0]) | ((~w[1] - 0x01010101u) & w[1]) | ((~w[2] - 0x01010101u) & w[2]) |
                             ((~w[3] - 0x01010101u) & w[3])) & 0x80808080u;
    c.plain = full && any_ff == 0u && before != 0xFFu;
    if (c.plain) {
#pragma unroll
        for (int i = 0; i < 4; ++i) c.emit[i] = 0x80808080u, c.split[i] = 0u, c.out[i] = c.cur[i] = w[i];
        return c;
    }
    uint32_t carry = before == 0xFFu ? 0x80u : 0u;  // "the previous byte was FF", as bit 7 of the byte before byte 0
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint32_t valid = full ? 0xFFFFFFFFu : (bytes_below(e - 4 * i) & ~bytes_below(s - 4 * i));
        const uint32_t cur = w[i] & valid, v80 = valid & 0x80808080u;
        const uint32_t cur_ff = zero_bytes(~cur);  // (bytes outside the image are 0 here: never FF)
        const uint32_t prev_ff = (cur_ff << 8) | carry;  // little endian: the byte before byte j sits 8 bits lower
        carry = cur_ff >> 24;
        const uint32_t cur_00 = zero_bytes(cur);
        const uint32_t cur_rst = zero_bytes((cur & 0xF8F8F8F8u) ^ 0xD0D0D0D0u);
        c.emit[i] = ((~prev_ff & ~cur_ff) | (prev_ff & cur_00)) & v80;
        c.split[i] = prev_ff & cur_rst & v80;
        c.out[i] = cur | ((prev_ff >> 7) * 0xFFu);
        c.cur[i] = cur;
        c.foreign |= prev_ff & ~cur_00 & ~cur_ff & ~cur_rst & v80;
    }
    return c;
}

__device__ __forceinline__ uint32_t count_bytes(const uint32_t m[4]) { return __popc(m[0]) + __popc(m[1]) + __popc(m[2]) + __popc(m[3]); }

// block-wide exclusive scan of a packed (emit | split << 16) value; returns the exclusive prefix, *total = block sum
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t *total)
{
    __shared__ uint32_t s_warp[LEX_THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t  inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, inc, d);
        if (lane >= d) inc += o;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    uint32_t base = 0, sum = 0;
#pragma unroll
    for (int i = 0; i < LEX_THREADS / 32; ++i) {
        const uint32_t t = s_warp[i];
        if (i < warp) base += t;
        sum += t;
    }
    *total = sum;
    __syncthreads();
    return base + inc - v;
}

// pass 1: per-tile totals.  grid = (tiles_max, n_images)
__global__ void __launch_bounds__(LEX_THREADS) k_lex_count(const uint8_t *__restrict__ raw, const LexImage *__restrict__ images,
                                                            uint32_t tiles_max, uint32_t *__restrict__ tile_counts,
                                                            uint32_t *__restrict__ foreign, unsigned long long *__restrict__ img_emit,
                                                            uint32_t *__restrict__ img_split)
{
    // a CTA takes LEX_COUNT_TILES consecutive tiles: their 16-byte loads are independent and all in flight together (with one
    // tile per CTA the pass was bound by the latency of its short dependent chain image table -> bytes, not by bandwidth)
    const uint32_t img = blockIdx.y, tile0 = blockIdx.x * LEX_COUNT_TILES;
    const LexImage im = images[img];
    const uint64_t lo = im.raw_off, hi = im.raw_off + im.raw_len;
    __shared__ uint32_t s_total[LEX_COUNT_TILES];
    if (threadIdx.x < LEX_COUNT_TILES) s_total[threadIdx.x] = 0;
    __syncthreads();
    uint32_t bad = 0, v[LEX_COUNT_TILES];
#pragma unroll
    for (int t = 0; t < LEX_COUNT_TILES; ++t) {
        const uint64_t base = (lo & ~15ull) + (uint64_t) (tile0 + t) * LEX_TILE + threadIdx.x * LEX_CHUNK;
        v[t] = 0;
        if (base < hi) {
            const Chunk c = classify(raw, base, lo, hi);
            v[t] = c.plain ? 16u : (count_bytes(c.emit) | (count_bytes(c.split) << 16));
            bad |= c.foreign;
        }
    }
    // only the tile totals are needed here: a warp reduction and one shared-memory atomic per warp and tile (not a block scan)
#pragma unroll
    for (int t = 0; t < LEX_COUNT_TILES; ++t) {
        const uint32_t wsum = __reduce_add_sync(0xFFFFFFFFu, v[t]);
        if ((threadIdx.x & 31) == 0 && wsum) atomicAdd(&s_total[t], wsum);
    }
    __syncthreads();
    if (threadIdx.x < LEX_COUNT_TILES && tile0 + threadIdx.x < tiles_max) {
        const uint32_t total = s_total[threadIdx.x];
        tile_counts[(size_t) img * tiles_max + tile0 + threadIdx.x] = total;
        if (total) {  // per-image totals: the second level of the scan
            atomicAdd(&img_emit[img], (unsigned long long) (total & 0xFFFFu));
            if (total >> 16) atomicAdd(&img_split[img], total >> 16);
        }
    }
    if (bad) atomicOr(&foreign[img], 1u);
}

// pass 2: one CTA per image.  Exclusive scan over all (image-major) tile totals -- 64-bit emitted-byte bases, 32-bit split bases --
// in two levels: the image's base is the sum of the totals of the images before it (k_lex_count accumulated them), its tiles
// are scanned by the CTA.  (A single CTA walking all n_images * tiles_max totals took 70 us of the lexer's 360.)
__global__ void __launch_bounds__(1024) k_lex_scan(const uint32_t *__restrict__ tile_counts, uint32_t tiles_max, uint32_t n_images,
                                                   const unsigned long long *__restrict__ img_emit, const uint32_t *__restrict__ img_split,
                                                   uint64_t *__restrict__ emit_base, uint32_t *__restrict__ split_base)
{
    __shared__ unsigned long long s_base_e;
    __shared__ uint32_t           s_base_s, s_we[32], s_ws[32];
    const uint32_t img = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_base_e = 0, s_base_s = 0;
    __syncthreads();
    {
        unsigned long long e = 0;
        uint32_t           sp = 0;
        for (uint32_t j = tid; j < img; j += 1024) e += img_emit[j], sp += img_split[j];
        if (e) atomicAdd(&s_base_e, e);
        if (sp) atomicAdd(&s_base_s, sp);
    }
    __syncthreads();
    uint64_t carry_e = s_base_e;
    uint32_t carry_s = s_base_s;
    for (uint32_t t0 = 0; t0 < tiles_max; t0 += 1024) {
        const uint32_t t = t0 + tid;
        const uint32_t c = t < tiles_max ? tile_counts[(size_t) img * tiles_max + t] : 0u;
        uint32_t       e = c & 0xFFFFu, sp = c >> 16;  // a chunk of 1024 tiles emits < 2^22 bytes
        const uint32_t e0 = e, s0 = sp;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t oe = __shfl_up_sync(0xFFFFFFFFu, e, d), os = __shfl_up_sync(0xFFFFFFFFu, sp, d);
            if (lane >= (uint32_t) d) e += oe, sp += os;
        }
        if (lane == 31) s_we[warp] = e, s_ws[warp] = sp;
        __syncthreads();
        uint32_t be = 0, bs = 0, te = 0, ts = 0;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            const uint32_t we = s_we[i], ws = s_ws[i];
            if ((uint32_t) i < warp) be += we, bs += ws;
            te += we, ts += ws;
        }
        if (t < tiles_max) {
            emit_base[(size_t) img * tiles_max + t] = carry_e + be + e - e0;
            split_base[(size_t) img * tiles_max + t] = carry_s + bs + sp - s0;
        }
        carry_e += te, carry_s += ts;
        __syncthreads();
    }
    if (img + 1 == n_images && tid == 0) {
        emit_base[(size_t) n_images * tiles_max] = carry_e;
        split_base[(size_t) n_images * tiles_max] = carry_s;
    }
}

// pass 3: scatter.  grid = (tiles_max, n_images)
__global__ void __launch_bounds__(LEX_THREADS) k_lex_scatter(const uint8_t *__restrict__ raw, const LexImage *__restrict__ images,
                                                              uint32_t tiles_max, const uint64_t *__restrict__ emit_base,
                                                              const uint32_t *__restrict__ split_base, uint32_t n_ecs,
                                                              uint8_t *__restrict__ out, uint64_t *__restrict__ offsets,
                                                              uint32_t *__restrict__ bad_phase)
{
    __shared__ __align__(16) uint8_t s_out[LEX_TILE + 32];
    const uint32_t img = blockIdx.y, tile = blockIdx.x;
    const LexImage im = images[img];
    const uint64_t lo = im.raw_off, hi = im.raw_off + im.raw_len;
    const uint64_t tile_base = (lo & ~15ull) + (uint64_t) tile * LEX_TILE;
    if (tile_base >= hi) return;
    const size_t   ti = (size_t) img * tiles_max + tile;
    const uint64_t g0 = emit_base[ti];
    const uint32_t k0 = split_base[ti] - split_base[(size_t) img * tiles_max];
    const uint64_t base = tile_base + threadIdx.x * LEX_CHUNK;
    Chunk          c;
    uint32_t       v = 0;
    if (base < hi) {
        c = classify(raw, base, lo, hi);
        v = c.plain ? 16u : (count_bytes(c.emit) | (count_bytes(c.split) << 16));
    }
    uint32_t       total;
    const uint32_t ex = block_exclusive_scan(v, &total);
    const uint32_t a = (uint32_t) (g0 & 15);  // keep the destination's alignment inside the staging tile
    if (v) {
        uint32_t pos = a + (ex & 0xFFFFu);
        uint32_t k = k0 + (ex >> 16);
        if ((v >> 16) == 0 && (v & 0xFFFFu) == 16) {
            // all 16 bytes go out (94 % of the chunks): word stores through a byte funnel for the middle, byte stores only for the
            // up to three bytes on either side that share a word with a neighbour's output
            const uint32_t sh = pos & 3u;
            if (sh == 0u) {
#pragma unroll
                for (int i = 0; i < 4; ++i) *reinterpret_cast<uint32_t *>(s_out + pos + 4 * i) = c.out[i];
            } else {
                const uint32_t head = 4u - sh;  // bytes that complete the first (shared) word
                for (uint32_t q = 0; q < head; ++q) s_out[pos + q] = (uint8_t) (c.out[0] >> (8 * q));
#pragma unroll
                for (int i = 0; i < 3; ++i)
                    *reinterpret_cast<uint32_t *>(s_out + pos + head + 4 * i) = __funnelshift_r(c.out[i], c.out[i + 1], 8u * head);
                for (uint32_t q = 0; q < sh; ++q) s_out[pos + head + 12 + q] = (uint8_t) (c.out[3] >> (8 * (head + q)));
            }
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if ((c.emit[i] >> (8 * j + 7)) & 1u) s_out[pos++] = (uint8_t) (c.out[i] >> (8 * j));
                    if ((c.split[i] >> (8 * j + 7)) & 1u) {
                        // the ECS that ends here has index k; the marker's phase must be k mod 8 (decode.swift:3929)
                        const uint32_t phase = ((c.cur[i] >> (8 * j)) & 7u);
                        if (phase != (k & 7u)) atomicMin(&bad_phase[img], k);
                        if (k + 1 < n_ecs) offsets[(size_t) img * n_ecs + k + 1] = g0 + (pos - a);
                        ++k;
                    }
                }
        }
    }
    __syncthreads();
    const uint32_t n = total & 0xFFFFu;
    if (n == 0) return;
    uint8_t *gdst = out + (g0 - a);  // 16-byte aligned (out is)
    for (uint32_t u = threadIdx.x; 16 * u < a + n; u += LEX_THREADS) {
        const uint32_t p = 16 * u;
        if (p >= a && p + 16 <= a + n)
            *reinterpret_cast<uint4 *>(gdst + p) = *reinterpret_cast<const uint4 *>(s_out + p);
        else
            for (uint32_t q = (p > a ? p : a); q < p + 16 && q < a + n; ++q) gdst[q] = s_out[q];
    }
}

// pass 4: one thread per image: first / missing / final offsets and the lexer status
__global__ void k_lex_finish(uint32_t n_images, uint32_t tiles_max, uint32_t n_ecs, const uint64_t *__restrict__ emit_base,
                             const uint32_t *__restrict__ split_base, const uint32_t *__restrict__ foreign,
                             const uint32_t *__restrict__ bad_phase, uint64_t *__restrict__ offsets, int32_t *__restrict__ status,
                             uint32_t *__restrict__ n_splits)
{
    const uint32_t img = blockIdx.x * blockDim.x + threadIdx.x;
    if (img >= n_images) return;
    const size_t   t0 = (size_t) img * tiles_max, t1 = t0 + tiles_max;
    const uint64_t begin = emit_base[t0], end = emit_base[t1];
    const uint32_t splits = split_base[t1] - split_base[t0];
    if (n_splits) n_splits[img] = splits;
    if (n_ecs) {
        offsets[(size_t) img * n_ecs] = begin;
        for (uint32_t k = splits + 1; k < n_ecs; ++k) offsets[(size_t) img * n_ecs + k] = end;
        if (img == n_images - 1) offsets[(size_t) n_images * n_ecs] = end;
    }
    int32_t st = JPEG_SM100_OK;
    if (foreign[img])
        st = JPEG_SM100_ERR_INVALID_ARGUMENT;
    else if (bad_phase[img] != 0xFFFFFFFFu)
        st = JPEG_SM100_ERR_RESTART_PHASE;
    else if (n_ecs && splits + 1 != n_ecs)
        st = JPEG_SM100_ERR_ECS_COUNT;
    if (status) status[img] = st;
}

}  // namespace

// Device-side state of one lexing job: scratch slot 15 holds the image table, the tile tables and the flags.
// (same definition in api.cu)
struct LexPlan {
    uint32_t  n_images, tiles_max;
    uint64_t  n_tiles;
    void     *d_images;
    uint32_t *d_counts, *d_split_base, *d_foreign, *d_bad_phase, *d_n_splits;
    uint64_t *d_emit_base;
};

// passes 1 + 2 (independent of n_ecs).  raw_offsets / raw_lengths are HOST arrays.
int jpeg_lex_count(jpeg_sm100_ctx *ctx, const uint8_t *d_raw, const uint64_t *raw_offsets, const uint64_t *raw_lengths,
                   uint32_t n_images, LexPlan *plan)
{
    uint64_t longest = 0;
    for (uint32_t i = 0; i < n_images; ++i) {
        const uint64_t span = (raw_offsets[i] & 15) + raw_lengths[i];
        longest = span > longest ? span : longest;
    }
    const uint64_t tiles_max64 = longest ? (longest + LEX_TILE - 1) / LEX_TILE : 1;
    if (tiles_max64 > 0x7FFFFFFFull || n_images > 65535) return JPEG_SM100_ERR_UNSUPPORTED;
    plan->n_images = n_images;
    plan->tiles_max = (uint32_t) tiles_max64;
    plan->n_tiles = (uint64_t) n_images * plan->tiles_max;
    // layout of slot 15: images | counts | emit_base | split_base | foreign | bad_phase | n_splits
    size_t       off = 0;
    const size_t o_images = off;
    off += (sizeof(LexImage) * n_images + 255) & ~(size_t) 255;
    const size_t o_counts = off;
    off += (4 * plan->n_tiles + 255) & ~(size_t) 255;
    const size_t o_emit = off;
    off += (8 * (plan->n_tiles + 1) + 255) & ~(size_t) 255;
    const size_t o_split = off;
    off += (4 * (plan->n_tiles + 1) + 255) & ~(size_t) 255;
    const size_t o_flags = off;
    off += ((size_t) 12 * n_images + 255) & ~(size_t) 255;
    const size_t o_totals = off;  // per image: emitted bytes (u64) | splits (u32)
    off += ((size_t) 12 * n_images + 255) & ~(size_t) 255;
    void *base = nullptr;
    J_TRY(scratch_reserve(ctx, 15, off, &base));
    uint8_t *b = reinterpret_cast<uint8_t *>(base);
    plan->d_images = b + o_images;
    plan->d_counts = reinterpret_cast<uint32_t *>(b + o_counts);
    plan->d_emit_base = reinterpret_cast<uint64_t *>(b + o_emit);
    plan->d_split_base = reinterpret_cast<uint32_t *>(b + o_split);
    plan->d_foreign = reinterpret_cast<uint32_t *>(b + o_flags);
    plan->d_bad_phase = plan->d_foreign + n_images;
    plan->d_n_splits = plan->d_bad_phase + n_images;

    void *h = nullptr;
    int   slot = 0;
    J_TRY(pinned_acquire(ctx, sizeof(LexImage) * n_images, &h, &slot));
    LexImage *hi = reinterpret_cast<LexImage *>(h);
    for (uint32_t i = 0; i < n_images; ++i) hi[i] = LexImage{raw_offsets[i], raw_lengths[i]};
    CU_TRY(ctx, cudaMemcpyAsync(plan->d_images, hi, sizeof(LexImage) * n_images, cudaMemcpyHostToDevice, ctx->stream));
    J_TRY(pinned_release(ctx, slot));
    CU_TRY(ctx, cudaMemsetAsync(plan->d_foreign, 0, 4 * (size_t) n_images, ctx->stream));
    CU_TRY(ctx, cudaMemsetAsync(plan->d_bad_phase, 0xFF, 4 * (size_t) n_images, ctx->stream));
    unsigned long long *d_img_emit = reinterpret_cast<unsigned long long *>(b + o_totals);
    uint32_t           *d_img_split = reinterpret_cast<uint32_t *>(b + o_totals + 8 * (size_t) n_images);
    CU_TRY(ctx, cudaMemsetAsync(d_img_emit, 0, 12 * (size_t) n_images, ctx->stream));
    const dim3 grid((plan->tiles_max + LEX_COUNT_TILES - 1) / LEX_COUNT_TILES, n_images);
    k_lex_count<<<grid, LEX_THREADS, 0, ctx->stream>>>(d_raw, reinterpret_cast<const LexImage *>(plan->d_images), plan->tiles_max, plan->d_counts,
                                                       plan->d_foreign, d_img_emit, d_img_split);
    LAUNCH_CHECK(ctx);
    k_lex_scan<<<n_images, 1024, 0, ctx->stream>>>(plan->d_counts, plan->tiles_max, n_images, d_img_emit, d_img_split, plan->d_emit_base,
                                                   plan->d_split_base);
    LAUNCH_CHECK(ctx);
    return JPEG_SM100_OK;
}

// n_ecs == 0: only the per-image split counts are produced (d_n_splits), nothing is written to d_ecs / d_offsets
int jpeg_lex_scatter(jpeg_sm100_ctx *ctx, const uint8_t *d_raw, const LexPlan *plan, uint32_t n_ecs, uint8_t *d_ecs,
                     uint64_t *d_offsets, int32_t *d_status)
{
    if (n_ecs) {
        const dim3 grid(plan->tiles_max, plan->n_images);
        k_lex_scatter<<<grid, LEX_THREADS, 0, ctx->stream>>>(d_raw, reinterpret_cast<const LexImage *>(plan->d_images), plan->tiles_max, plan->d_emit_base,
                                                             plan->d_split_base, n_ecs, d_ecs, d_offsets, plan->d_bad_phase);
        LAUNCH_CHECK(ctx);
    }
    k_lex_finish<<<(plan->n_images + 127) / 128, 128, 0, ctx->stream>>>(plan->n_images, plan->tiles_max, n_ecs, plan->d_emit_base,
                                                                        plan->d_split_base, plan->d_foreign, plan->d_bad_phase,
                                                                        d_offsets, d_status, plan->d_n_splits);
    LAUNCH_CHECK(ctx);
    return JPEG_SM100_OK;
}

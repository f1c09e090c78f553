// encode.cu -- K6/K7: entropy encode = symbolise + histogram -> optimal Huffman tables (host) -> bit lengths ->
// prefix sums -> bit packing -> byte stuffing (+ optional RSTn markers).
//
// Replaces Spectral.encode(scan:) (reference encode.swift:1559-1620): block symbolisation (919-959), the sequential
// / DC / AC scan encoders (962-1557), Composite.decomposed (850-879), optimal table construction
// (602-772 with General.Heap, common.swift:127-296), canonical code assignment (664-680, 799-820) and
// Bitstream.append / bytes(escaping:with:) (jpeg.swift:1920-2007).
//
// B200 design
//   * sequential, DC-first and DC-refinement scans are embarrassingly parallel once the DC predecessor is read from the
//     neighbouring block: one THREAD per 8x8 block in scan order for the histogram, bit-length and emit passes;
//     variable-length output is placed with two prefix sums (bits per block, stuffed bytes per chunk) -- hazard H4.
//   * bits are OR-ed into a zeroed big-endian word stream with fire-and-forget atomics (RED.OR), so neighbouring
//     blocks that share a word never race.
//   * Huffman table construction stays on the host (<= 257 symbols; must reproduce the reference's heap tie-breaking
//     bit for bit) between the statistics pass and the emit pass.
//   * progressive AC scans: end-of-band runs tie blocks together, but who heads a run follows from one flag per block and two
//     scans (k_enc_acp_flags / k_enc_acp_chain), after which they too run one thread per block through the same three passes.
#include <time.h>

#include <algorithm>
#include <atomic>
#include <thread>

#include "common.cuh"

namespace {

// ================================================================================================================
// host: optimal Huffman tables exactly as the reference builds them
// ================================================================================================================
struct MinHeap {  // common.swift:127-296: 1-based binary heap, strict '<' comparisons
    std::vector<std::pair<int64_t, int>> a{{0, 0}};
    int  count() const { return (int) a.size() - 1; }
    void sift_down(int i)
    {
        for (;;) {
            const int l = i << 1, r = l + 1, end = count() + 1;
            if (!(l < end)) return;
            int c;
            if (!(r < end)) {
                if (!(a[l].first < a[i].first)) return;
                c = l;
            } else {
                c = a[r].first < a[l].first ? r : l;
                if (!(a[c].first < a[i].first)) return;
            }
            std::swap(a[i], a[c]);
            i = c;
        }
    }
    void sift_up(int i)
    {
        for (int p = i >> 1; p >= 1 && a[i].first < a[p].first; i = p, p = i >> 1) std::swap(a[i], a[p]);
    }
    void enqueue(int64_t key, int value)
    {
        a.emplace_back(key, value);
        sift_up(count());
    }
    bool dequeue(std::pair<int64_t, int> &out)
    {
        if (count() == 0) return false;
        if (count() > 1) std::swap(a[1], a[count()]);
        out = a.back();
        a.pop_back();
        if (count() > 1) sift_down(1);
        return true;
    }
};

// encode.swift:701-772 init(frequencies:target:), 602-661 limit(height:of:)
void huffman_from_frequencies(const uint32_t freq[256], jpeg_sm100_huff_table &out)
{
    memset(&out, 0, sizeof out);
    std::vector<std::pair<int64_t, int>> sorted;  // (frequency, symbol)
    for (int v = 0; v < 256; ++v)
        if (freq[v] > 0) sorted.emplace_back((int64_t) freq[v], v);
    if (sorted.empty()) return;
    std::stable_sort(sorted.begin(), sorted.end(), [](const auto &x, const auto &y) { return x.first > y.first; });
    struct Node { int left, right; };
    std::vector<Node> tree;
    MinHeap           heap;
    for (auto it = sorted.rbegin(); it != sorted.rend(); ++it) {
        heap.a.emplace_back(it->first, (int) tree.size());
        tree.push_back({-1, -1});
    }
    for (int i = (heap.count() >> 1); i >= 1; --i) heap.sift_down(i);  // heapify: (startIndex ..< halfway).reversed()
    heap.enqueue(0, (int) tree.size());  // the dummy leaf that reserves the all-ones codeword
    tree.push_back({-1, -1});
    std::pair<int64_t, int> first, second;
    int                     root = -1;
    while (heap.dequeue(first)) {
        if (!heap.dequeue(second)) {
            root = first.second;
            break;
        }
        heap.enqueue(first.first + second.first, (int) tree.size());
        tree.push_back({first.second, second.second});
    }
    // leaves per depth, breadth first (encode.swift:576-595), root level dropped
    std::vector<int> levels, queue{root};
    while (!queue.empty()) {
        std::vector<int> next;
        int              leaves = 0;
        for (int n : queue) {
            if (tree[n].left < 0) ++leaves;
            else {
                next.push_back(tree[n].left);
                next.push_back(tree[n].right);
            }
        }
        levels.push_back(leaves);
        queue.swap(next);
    }
    levels.erase(levels.begin());
    const int height = 16;
    if ((int) levels.size() <= height) levels.back() -= 1;
    else {
        int unhoused = 0;
        for (int l = (int) levels.size() - 1; l >= height; --l) {
            const int pairs = levels[l] >> 1;
            unhoused += pairs;
            levels[l - 1] += pairs;
        }
        levels.resize(height);
        int split = height - 2;
        while (unhoused > 0) {
            if (!(levels[split] > 0)) {
                --split;
                continue;
            }
            const int resettled = std::min(levels[split], unhoused);
            unhoused -= resettled;
            levels[split] -= resettled;
            levels[split + 1] += 2 * resettled;
            if (split < height - 2) ++split;
        }
        levels[height - 1] -= 1;
    }
    int base = 0;
    for (size_t l = 0; l < levels.size() && l < 16; ++l) {
        out.counts[l] = (uint8_t) levels[l];
        for (int i = 0; i < levels[l]; ++i) out.values[base + i] = (uint8_t) sorted[base + i].second;
        base += levels[l];
    }
    out.present = 1;
}

// encode.swift:664-680 assign, 799-820 encoder(): code | length << 16, indexed by symbol
void canonical_codes(const jpeg_sm100_huff_table &t, uint32_t out[256])
{
    memset(out, 0, 256 * sizeof(uint32_t));
    uint32_t counter = 0;
    int      base = 0;
    for (int l = 0; l < 16; ++l) {
        for (int i = 0; i < t.counts[l]; ++i) out[t.values[base + i]] = (counter++ & 0xffffu) | ((uint32_t) (l + 1) << 16);
        base += t.counts[l];
        counter <<= 1;
    }
}

// ================================================================================================================
// device
// ================================================================================================================
struct EncParams {
    int32_t  kind, band_lo, band_hi, al;
    int32_t  n_comp, W, H, mcu_blocks;
    uint8_t  blk_comp[12], blk_dx[12], blk_dy[12], blk_first[12], blk_count[12];
    const int16_t *plane[4];
    uint64_t image_stride[4];
    int32_t  ux[4], uy[4], fx[4], fy[4];
    int32_t  dc[4], ac[4];          // table indices (dc slot, 4 + ac slot)
    uint32_t blocks_per_interval;   // scan blocks per restart interval
    uint32_t S;                     // scan blocks per image
    uint32_t n_intervals;
    uint32_t *hist;                 // n_images x 8 x 256
    const uint32_t *codes;          // n_images x 8 x 256  (code | len << 16)
    uint32_t *blk_bits;             // n_images x S: bits per block, then exclusive offset within its interval
    uint32_t *ac_info;              // n_images x S (progressive AC scans): flags, then run | absorbed (see k_enc_acp_chain)
    uint64_t *ivl;                  // n_images x (n_intervals + 1): bits per interval, then byte offset; [n] = total bytes
    uint32_t *raw;                  // unstuffed stream, big-endian words, zeroed
    uint64_t  raw_stride;           // bytes per image
    uint8_t  *out;
    uint64_t  out_stride;
    uint64_t *out_len;
};

// decode.swift:2757-2771 compact(): (binade, tail)
__device__ __forceinline__ void compact16(int x, int &binade, uint32_t &tail)
{
    const int mag = x < 0 ? -x : x;
    binade = 32 - __clz(mag);
    const uint32_t sign = ((uint32_t) x >> 15) & 1u;
    tail = (((uint32_t) x & 0xffffu) - sign) & ((1u << binade) - 1u);
}

struct BlockRef {
    const int16_t *ptr;  // nullptr: out of plane -> all zeros (decode.swift:1459-1464)
};

__device__ __forceinline__ BlockRef locate(const EncParams &P, uint32_t img, uint32_t s, int &comp, int &b)
{
    const uint32_t mcu = s / (uint32_t) P.mcu_blocks;
    b = (int) (s - mcu * P.mcu_blocks);
    const uint32_t my = mcu / (uint32_t) P.W, mx = mcu - my * P.W;
    comp = P.blk_comp[b];
    const int x = (int) mx * P.fx[comp] + P.blk_dx[b], y = (int) my * P.fy[comp] + P.blk_dy[b];
    BlockRef r;
    r.ptr = (x < P.ux[comp] && y < P.uy[comp])
                ? P.plane[comp] + (size_t) img * P.image_stride[comp] + 64 * ((size_t) P.ux[comp] * y + x)
                : nullptr;
    return r;
}

// DC value of the previous block of the same component in scan order (0 at the start of an interval)
__device__ __forceinline__ int predecessor_dc(const EncParams &P, uint32_t img, uint32_t s, int b)
{
    const uint32_t in_ivl = s % P.blocks_per_interval;
    uint32_t       prev;
    if (!P.blk_first[b]) prev = s - 1;
    else {
        if (in_ivl < (uint32_t) P.mcu_blocks) return 0;  // first MCU of the interval: predictor reset
        prev = s - P.mcu_blocks + P.blk_count[b] - 1;
    }
    int      c, pb;
    BlockRef r = locate(P, img, prev, c, pb);
    return r.ptr ? (int) r.ptr[0] : 0;
}

// sinks ---------------------------------------------------------------------------------------------------------
struct HistSink {
    uint32_t *h;  // shared: 8 x 256
    __device__ __forceinline__ void symbol(int table, int sym, uint32_t, int) { atomicAdd(&h[table * 256 + sym], 1u); }
    __device__ __forceinline__ void raw(uint32_t, int) {}
};
struct LenSink {
    const uint32_t *codes;
    uint32_t        bits;
    __device__ __forceinline__ void symbol(int table, int sym, uint32_t, int nextra) { bits += (codes[table * 256 + sym] >> 16) + nextra; }
    __device__ __forceinline__ void raw(uint32_t, int n) { bits += n; }
};
struct EmitSink {
    const uint32_t *codes;
    uint32_t       *words;  // big-endian word stream of this image
    uint64_t        acc;    // MSB-aligned pending bits
    int             filled;
    uint64_t        word;   // index of the word acc starts at
    __device__ __forceinline__ void begin(uint64_t bitpos)
    {
        word = bitpos >> 5;
        filled = (int) (bitpos & 31);
        acc = 0;
    }
    __device__ __forceinline__ void put(uint32_t v, int n)
    {
        if (n == 0) return;
        acc |= (uint64_t) (v & ((n >= 32) ? 0xffffffffu : ((1u << n) - 1u))) << (64 - filled - n);
        filled += n;
        if (filled >= 32) {
            atomicOr(&words[word], __byte_perm((uint32_t) (acc >> 32), 0, 0x0123));
            acc <<= 32;
            filled -= 32;
            word += 1;
        }
    }
    __device__ __forceinline__ void symbol(int table, int sym, uint32_t extra, int nextra)
    {
        // code and extra bits leave as ONE bit string (ncu: the two shift-mask-merge sequences of put() were a third of the emit
        // pass's instructions at 9 of 32 lanes); only a 16-bit code followed by 16 extra bits (DC category 16) needs two
        const uint32_t c = codes[table * 256 + sym];
        const int      len = (int) (c >> 16);
        if (len + nextra <= 32 && nextra < 32) {
            put(((c & 0xffffu) << nextra) | (extra & ((1u << nextra) - 1u)), len + nextra);
        } else {
            put(c & 0xffffu, len);
            put(extra, nextra);
        }
    }
    __device__ __forceinline__ void raw(uint32_t v, int n) { put(v, n); }
    __device__ __forceinline__ void pad_to_byte()
    {
        const int r = (8 - (filled & 7)) & 7;
        put(0xffu, r);  // jpeg.swift:1884-1887: pad with 1-bits
    }
    __device__ __forceinline__ void finish()
    {
        if (filled > 0) atomicOr(&words[word], __byte_perm((uint32_t) (acc >> 32), 0, 0x0123));
    }
};

// progressive AC scans read their block from a transposed copy in shared memory (AcView, below)
struct AcView {
    const uint32_t *col;  // this thread's column: word w of the block at col[32 * w]
    __device__ __forceinline__ int operator[](int z) const { return (int) (short) (col[32 * (z >> 1)] >> (16 * (z & 1))); }
};
template <class Sink>
__device__ __forceinline__ uint32_t encode_block_ac(const EncParams &P, const AcView pl, uint32_t info, Sink &sink);  // (below)
// A thread's block is 128 contiguous bytes, so the lanes of a warp read 32 different lines whenever they read "coefficient z":
// 63 two-byte loads per block and pass cost 63 x 32 L1 wavefronts per warp (ncu-less arithmetic: 1.9 ms per pass over 8.3 M
// blocks).  Instead every lane fetches its block with eight 16-byte loads and parks it as a column of a word-transposed tile in
// shared memory (conflict-free to write and to read): 63 single-wavefront shared loads per pass.
__device__ __forceinline__ AcView stage_block_ac(const EncParams &P, uint32_t img, uint32_t blk, bool valid, uint32_t *tile /* [32][blockDim.x] */)
{
    uint32_t *col = tile + (threadIdx.x & ~31u) * 32u + (threadIdx.x & 31u);
    if (valid) {
        const uint4 *src = reinterpret_cast<const uint4 *>(P.plane[0] + (size_t) img * P.image_stride[0] + (size_t) blk * 64);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const uint4 v = __ldg(src + j);
            col[32 * (4 * j + 0)] = v.x, col[32 * (4 * j + 1)] = v.y, col[32 * (4 * j + 2)] = v.z, col[32 * (4 * j + 3)] = v.w;
        }
    }
    __syncwarp();
    return AcView{col};
}

// encode.swift:919-959 (sequential) / 1013-1044, 1386-1510 (DC first) / 1046-1058, 1513-1557 (DC refine)
template <class Sink>
__device__ __forceinline__ void encode_block(const EncParams &P, uint32_t img, uint32_t s, Sink &sink)
{
    int            comp, b;
    const BlockRef r = locate(P, img, s, comp, b);
    if (P.kind == 2) {
        const int c0 = r.ptr ? (int) r.ptr[0] : 0;
        sink.raw((uint32_t) ((c0 >> P.al) & 1), 1);
        return;
    }
    const int c0 = r.ptr ? (int) r.ptr[0] : 0;
    const int pred = predecessor_dc(P, img, s, b);
    int       diff;
    if (P.kind == 1) diff = (int) (short) ((c0 >> P.al) - (pred >> P.al));
    else diff = (int) (short) (c0 - pred);
    int      binade;
    uint32_t tail;
    compact16(diff, binade, tail);
    sink.symbol(P.dc[comp], binade, tail, binade);
    if (P.kind == 1) return;
    const int ac = P.ac[comp];
    int       zeroes = 0;
    if (r.ptr) {
        const uint4 *v = reinterpret_cast<const uint4 *>(r.ptr);
#pragma unroll 1
        for (int j = 0; j < 8; ++j) {
            const uint4    q = __ldg(v + j);
            const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                if (j == 0 && k == 0) continue;
                const int c = (int) (short) ((w[k >> 1] >> (16 * (k & 1))) & 0xffffu);
                if (c == 0) {
                    if (zeroes == 15) {
                        sink.symbol(ac, 0xf0, 0, 0);
                        zeroes = 0;
                    } else
                        ++zeroes;
                } else {
                    compact16(c, binade, tail);
                    sink.symbol(ac, (zeroes << 4) | binade, tail, binade);
                    zeroes = 0;
                }
            }
        }
    } else {
        // 63 zeros: ZRL, ZRL, ZRL, then 15 pending zeros
        sink.symbol(ac, 0xf0, 0, 0);
        sink.symbol(ac, 0xf0, 0, 0);
        sink.symbol(ac, 0xf0, 0, 0);
        zeroes = 15;
    }
    if (zeroes > 0) sink.symbol(ac, 0x00, 0, 0);
}

__global__ void __launch_bounds__(256) k_enc_hist(const __grid_constant__ EncParams P)
{
    __shared__ uint32_t h[8 * 256];
    extern __shared__ uint32_t ac_tile[];  // [32][256] for progressive AC scans (dynamic: the other kinds do not pay for it)
    for (int i = threadIdx.x; i < 8 * 256; i += blockDim.x) h[i] = 0;
    __syncthreads();
    const uint32_t img = blockIdx.y;
    HistSink       sink{h};
    if (P.kind >= 3) {
        uint32_t *tile = ac_tile;
        for (uint32_t s0 = blockIdx.x * blockDim.x; s0 < P.S; s0 += gridDim.x * blockDim.x) {
            const uint32_t s = s0 + threadIdx.x;
            const AcView   v = stage_block_ac(P, img, s, s < P.S, tile);
            if (s < P.S) encode_block_ac(P, v, P.ac_info[(size_t) img * P.S + s], sink);
            __syncwarp();
        }
    } else {
        for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < P.S; s += gridDim.x * blockDim.x) encode_block(P, img, s, sink);
    }
    __syncthreads();
    uint32_t *g = P.hist + (size_t) img * 8 * 256;
    for (int i = threadIdx.x; i < 8 * 256; i += blockDim.x)
        if (h[i]) atomicAdd(&g[i], h[i]);
}

__global__ void __launch_bounds__(256) k_enc_len(const __grid_constant__ EncParams P)
{
    extern __shared__ uint32_t ac_tile[];
    const uint32_t img = blockIdx.y;
    if (P.kind >= 3) {
        for (uint32_t s0 = blockIdx.x * blockDim.x; s0 < P.S; s0 += gridDim.x * blockDim.x) {
            const uint32_t s = s0 + threadIdx.x;
            const AcView   v = stage_block_ac(P, img, s, s < P.S, ac_tile);
            if (s < P.S) {
                LenSink sink{P.codes + (size_t) img * 8 * 256, 0u};
                encode_block_ac(P, v, P.ac_info[(size_t) img * P.S + s], sink);
                P.blk_bits[(size_t) img * P.S + s] = sink.bits;
            }
            __syncwarp();
        }
        return;
    }
    for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < P.S; s += gridDim.x * blockDim.x) {
        LenSink sink{P.codes + (size_t) img * 8 * 256, 0u};
        encode_block(P, img, s, sink);
        P.blk_bits[(size_t) img * P.S + s] = sink.bits;
    }
}

// one CTA per (interval, image): exclusive scan of the block bit lengths inside the interval, its total into ivl[e]
// (k_enc_ivl_offsets turns the totals into byte offsets).  One CTA per image walking its intervals one after the other
// (k_enc_scan_bits below, the first generation) took 0.19 ms of config #3's 6 ms.
__global__ void __launch_bounds__(1024) k_enc_scan_bits_ivl(const __grid_constant__ EncParams P)
{
    __shared__ uint32_t warp_sums[32];
    const uint32_t      e = blockIdx.x, img = blockIdx.y;
    uint32_t           *bits = P.blk_bits + (size_t) img * P.S;
    const int           lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint32_t      lo = e * P.blocks_per_interval, hi = min(P.S, lo + P.blocks_per_interval);
    uint64_t            carry = 0;  // (every thread keeps the same running total)
    for (uint32_t base = lo; base < hi; base += 1024) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t v = i < hi ? bits[i] : 0u;
        uint32_t       x = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
            if (lane >= d) x += y;
        }
        if (lane == 31) warp_sums[wid] = x;
        __syncthreads();
        uint32_t before = 0, total = 0;
#pragma unroll
        for (int w = 0; w < 32; ++w) {
            const uint32_t t = warp_sums[w];
            if (w < wid) before += t;
            total += t;
        }
        if (i < hi) bits[i] = (uint32_t) (carry + before + (x - v));  // < 2^32 bits per interval (checked on the host)
        carry += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) P.ivl[(size_t) img * (P.n_intervals + 1) + e] = carry;
}

// one CTA per image: exclusive scan of block bit lengths inside every interval; interval totals -> byte offsets
__global__ void __launch_bounds__(1024) k_enc_scan_bits(const __grid_constant__ EncParams P)
{
    __shared__ uint32_t warp_sums[32];
    __shared__ uint64_t carry;
    const uint32_t      img = blockIdx.x;
    uint32_t           *bits = P.blk_bits + (size_t) img * P.S;
    uint64_t           *ivl = P.ivl + (size_t) img * (P.n_intervals + 1);
    const int           lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (uint32_t e = 0; e < P.n_intervals; ++e) {
        const uint32_t lo = e * P.blocks_per_interval;
        const uint32_t hi = min(P.S, lo + P.blocks_per_interval);
        if (threadIdx.x == 0) carry = 0;
        __syncthreads();
        for (uint32_t base = lo; base < hi; base += 1024) {
            const uint32_t i = base + threadIdx.x;
            const uint32_t v = i < hi ? bits[i] : 0u;
            uint32_t       x = v;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
                if (lane >= d) x += y;
            }
            if (lane == 31) warp_sums[wid] = x;
            __syncthreads();
            if (wid == 0) {
                uint32_t w = warp_sums[lane];
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t y = __shfl_up_sync(0xffffffffu, w, d);
                    if (lane >= d) w += y;
                }
                warp_sums[lane] = w;
            }
            __syncthreads();
            const uint64_t c = carry;
            const uint32_t before = (wid ? warp_sums[wid - 1] : 0u) + (x - v);
            if (i < hi) bits[i] = (uint32_t) (c + before);  // < 2^32 bits per interval (checked on the host)
            __syncthreads();
            if (threadIdx.x == 1023) carry = c + warp_sums[31];
            __syncthreads();
        }
        if (threadIdx.x == 0) ivl[e] = carry;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        uint64_t off = 0;
        for (uint32_t e = 0; e < P.n_intervals; ++e) {
            const uint64_t nbits = ivl[e];
            ivl[e] = off;
            off += (nbits + 7) >> 3;
        }
        ivl[P.n_intervals] = off;
    }
}

__global__ void __launch_bounds__(256) k_enc_emit(const __grid_constant__ EncParams P)
{
    extern __shared__ uint32_t ac_tile[];
    const uint32_t  img = blockIdx.y;
    const uint64_t *ivl = P.ivl + (size_t) img * (P.n_intervals + 1);
    for (uint32_t s0 = blockIdx.x * blockDim.x; s0 < P.S; s0 += gridDim.x * blockDim.x) {
        const uint32_t s = s0 + threadIdx.x;
        AcView         v{nullptr};
        if (P.kind >= 3) v = stage_block_ac(P, img, s, s < P.S, ac_tile);
        if (s < P.S) {
            const uint32_t e = s / P.blocks_per_interval;
            EmitSink       sink;
            sink.codes = P.codes + (size_t) img * 8 * 256;
            sink.words = reinterpret_cast<uint32_t *>(reinterpret_cast<uint8_t *>(P.raw) + (size_t) img * P.raw_stride);
            sink.begin(ivl[e] * 8 + P.blk_bits[(size_t) img * P.S + s]);
            if (P.kind >= 3) encode_block_ac(P, v, P.ac_info[(size_t) img * P.S + s], sink);
            else encode_block(P, img, s, sink);
            const bool last = (s + 1 == P.S) || ((s + 1) % P.blocks_per_interval == 0);
            if (last) sink.pad_to_byte();
            sink.finish();
        }
        if (P.kind >= 3) __syncwarp();
    }
}

// one CTA per image: FF -> FF 00 stuffing (jpeg.swift:1977-2007) and RSTn insertion between intervals
__global__ void __launch_bounds__(1024) k_enc_stuff(const __grid_constant__ EncParams P)
{
    __shared__ uint32_t warp_sums[32];
    __shared__ uint64_t carry;
    const uint32_t      img = blockIdx.x;
    const uint64_t     *ivl = P.ivl + (size_t) img * (P.n_intervals + 1);
    const uint8_t      *raw = reinterpret_cast<const uint8_t *>(P.raw) + (size_t) img * P.raw_stride;
    uint8_t            *out = P.out + (size_t) img * P.out_stride;
    const int           lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    constexpr int       PER = 16;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint32_t e = 0; e < P.n_intervals; ++e) {
        const uint64_t lo = ivl[e], hi = ivl[e + 1];
        if (e > 0) {
            if (threadIdx.x == 0) {
                const uint64_t o = carry;
                if (o + 2 <= P.out_stride) {
                    out[o] = 0xff;
                    out[o + 1] = (uint8_t) (0xd0 + ((e - 1) & 7));
                }
                carry = o + 2;
            }
            __syncthreads();
        }
        for (uint64_t base = lo; base < hi; base += 1024 * PER) {
            const uint64_t i0 = base + (uint64_t) threadIdx.x * PER;
            uint8_t        b[PER];
            uint32_t       n = 0, ff = 0;
#pragma unroll
            for (int k = 0; k < PER; ++k) {
                if (i0 + k < hi) {
                    b[k] = raw[i0 + k];
                    ++n;
                    ff += b[k] == 0xff;
                }
            }
            const uint32_t v = n + ff;
            uint32_t       x = v;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
                if (lane >= d) x += y;
            }
            if (lane == 31) warp_sums[wid] = x;
            __syncthreads();
            if (wid == 0) {
                uint32_t w = warp_sums[lane];
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t y = __shfl_up_sync(0xffffffffu, w, d);
                    if (lane >= d) w += y;
                }
                warp_sums[lane] = w;
            }
            __syncthreads();
            const uint64_t c = carry;
            uint64_t       o = c + (wid ? warp_sums[wid - 1] : 0u) + (x - v);
            for (uint32_t k = 0; k < n; ++k) {
                if (o < P.out_stride) out[o] = b[k];
                ++o;
                if (b[k] == 0xff) {
                    if (o < P.out_stride) out[o] = 0x00;
                    ++o;
                }
            }
            __syncthreads();
            if (threadIdx.x == 1023) carry = c + warp_sums[31];
            __syncthreads();
        }
    }
    if (threadIdx.x == 0) P.out_len[img] = carry;
}

// ---- the same byte stuffing as a two-kernel stream expansion over 4 KB tiles of the raw stream (k_enc_stuff walks an image
// with ONE CTA: 1.26 ms per 64 4K frames; this is HBM-bound).  The output position of raw byte i is
//     i  +  (FF bytes before i)  +  2 * (index of the interval i lies in)            [every interval but the first is led by RSTn]
constexpr int STUFF_THREADS = 256, STUFF_CHUNK = 16, STUFF_TILE = STUFF_THREADS * STUFF_CHUNK;

__global__ void __launch_bounds__(STUFF_THREADS) k_enc_stuff_count(const __grid_constant__ EncParams P, uint32_t tiles_max, uint32_t *tile_ff)
{
    const uint32_t  img = blockIdx.y, tile = blockIdx.x;
    const uint64_t  total = P.ivl[(size_t) img * (P.n_intervals + 1) + P.n_intervals];
    const uint8_t  *raw = reinterpret_cast<const uint8_t *>(P.raw) + (size_t) img * P.raw_stride;
    const uint64_t  i0 = (uint64_t) tile * STUFF_TILE + threadIdx.x * STUFF_CHUNK;
    uint32_t        n = 0;
    if (i0 < total) {
        const uint4    v = *reinterpret_cast<const uint4 *>(raw + i0);  // (raw_stride keeps the tail readable)
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            uint32_t m = __vcmpeq4(w[k], 0xffffffffu) & 0x01010101u;
            if (i0 + 4 * k + 4 > total) m &= (i0 + 4 * k >= total) ? 0u : (0xffffffffu >> (8 * (4 - (int) (total - i0 - 4 * k))));
            n += __popc(m);
        }
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) n += __shfl_xor_sync(0xffffffffu, n, d);
    __shared__ uint32_t s_w[STUFF_THREADS / 32];
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = n;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int i = 0; i < STUFF_THREADS / 32; ++i) t += s_w[i];
        tile_ff[(size_t) img * tiles_max + tile] = t;
    }
}

__global__ void __launch_bounds__(STUFF_THREADS) k_enc_stuff_scatter(const __grid_constant__ EncParams P, uint32_t tiles_max, const uint32_t *tile_ff)
{
    __shared__ __align__(16) uint8_t s_out[4 * STUFF_TILE + 64];  // worst case per raw byte: FF 00 preceded by a marker
    __shared__ uint32_t s_w[STUFF_THREADS / 32], s_base, s_tile_total;
    const uint32_t  img = blockIdx.y, tile = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint64_t *ivl = P.ivl + (size_t) img * (P.n_intervals + 1);
    const uint64_t  total = ivl[P.n_intervals];
    const uint64_t  t0 = (uint64_t) tile * STUFF_TILE;
    if (t0 >= total && !(tile == 0 && total == 0)) return;
    const uint8_t *raw = reinterpret_cast<const uint8_t *>(P.raw) + (size_t) img * P.raw_stride;
    uint8_t       *out = P.out + (size_t) img * P.out_stride;
    // FF bytes in the tiles before this one
    {
        uint32_t part = 0;
        for (uint32_t t = tid; t < tile; t += STUFF_THREADS) part += tile_ff[(size_t) img * tiles_max + t];
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) part += __shfl_xor_sync(0xffffffffu, part, d);
        if (lane == 0) s_w[wid] = part;
        __syncthreads();
        if (tid == 0) {
            uint32_t t = 0;
            for (int i = 0; i < STUFF_THREADS / 32; ++i) t += s_w[i];
            s_base = t;
        }
        __syncthreads();
    }
    const uint64_t i0 = t0 + (uint64_t) tid * STUFF_CHUNK;
    uint8_t        b[STUFF_CHUNK];
    uint32_t       n = 0, ff = 0, marks = 0;
    uint32_t       e = 0;  // interval of the thread's first byte
    if (i0 < total) {
        const uint4    v = *reinterpret_cast<const uint4 *>(raw + i0);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
        n = total - i0 >= STUFF_CHUNK ? STUFF_CHUNK : (uint32_t) (total - i0);
#pragma unroll
        for (int k = 0; k < STUFF_CHUNK; ++k) {
            b[k] = (uint8_t) (w[k >> 2] >> (8 * (k & 3)));
            ff += ((uint32_t) k < n && b[k] == 0xff) ? 1u : 0u;
        }
        // upper_bound(ivl, i0) - 1 over the interval starts ivl[0 .. n_intervals)
        uint32_t lo = 0, hi = P.n_intervals;
        while (hi - lo > 1) {
            const uint32_t mid = (lo + hi) >> 1;
            if (ivl[mid] <= i0) lo = mid;
            else hi = mid;
        }
        e = lo;
        // interval boundaries strictly inside (i0, i0 + n): the markers this thread emits (a boundary AT i0 belongs to it too)
        uint32_t e2 = e;
        while (e2 + 1 < P.n_intervals && ivl[e2 + 1] < i0 + n) ++e2;
        marks = (e2 - e) + ((e > 0 && ivl[e] == i0) ? 1u : 0u);
    }
    // exclusive scan of the bytes each thread produces
    const uint32_t v = n + ff + 2u * marks;
    uint32_t       x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
        if (lane >= (uint32_t) d) x += y;
    }
    if (lane == 31) s_w[wid] = x;
    __syncthreads();
    uint32_t before = x - v;
    for (uint32_t i = 0; i < wid; ++i) before += s_w[i];
    if (tid == STUFF_THREADS - 1) s_tile_total = before + v;
    // where the tile's output starts: raw position + FFs before + 2 per interval boundary before the tile's first byte
    uint32_t e_tile = 0;
    {
        uint32_t lo = 0, hi = P.n_intervals;
        while (hi - lo > 1) {
            const uint32_t mid = (lo + hi) >> 1;
            if (ivl[mid] <= t0) lo = mid;
            else hi = mid;
        }
        e_tile = lo;
    }
    // markers before the tile: one per interval start in (0, t0); the one AT t0 (if any) is emitted by thread 0 of this tile
    const uint32_t marks_before = e_tile - ((e_tile > 0 && ivl[e_tile] == t0) ? 1u : 0u);
    const uint64_t g0 = t0 + s_base + 2ull * marks_before;
    const uint32_t a = (uint32_t) ((reinterpret_cast<uintptr_t>(out) + g0) & 15);  // keep the destination's alignment in the tile
    if (n) {
        uint32_t pos = a + before, ee = e;
        if (ee > 0 && ivl[ee] == i0) {  // the thread's first byte opens an interval
            s_out[pos++] = 0xff;
            s_out[pos++] = (uint8_t) (0xd0 + ((ee - 1) & 7));
        }
#pragma unroll
        for (int k = 0; k < STUFF_CHUNK; ++k) {
            if ((uint32_t) k < n) {
                if (k > 0 && ee + 1 < P.n_intervals && ivl[ee + 1] == i0 + k) {
                    ++ee;
                    s_out[pos++] = 0xff;
                    s_out[pos++] = (uint8_t) (0xd0 + ((ee - 1) & 7));
                }
                s_out[pos++] = b[k];
                if (b[k] == 0xff) s_out[pos++] = 0x00;
            }
        }
    }
    __syncthreads();
    const uint32_t produced = s_tile_total;
    // copy out: bytes up to the first 16-byte boundary, aligned vectors, tail bytes; nothing beyond the caller's capacity
    const uint64_t cap = P.out_stride;
    const uint32_t end = a + produced;
    for (uint32_t p = tid * 16; p < end; p += STUFF_THREADS * 16) {
        const uint64_t g = g0 - a + p;  // 16-byte aligned global offset of this piece
        if (p >= a && p + 16 <= end && g + 16 <= cap)
            *reinterpret_cast<uint4 *>(out + g) = *reinterpret_cast<const uint4 *>(s_out + p);
        else
            for (uint32_t q = p < a ? a : p; q < p + 16 && q < end; ++q)
                if (g0 - a + q < cap) out[g0 - a + q] = s_out[q];
    }
    if (t0 + STUFF_TILE >= total && tid == 0) P.out_len[img] = g0 + produced;  // the last tile knows the stuffed length
}

// ---- progressive AC scans: one thread per restart interval (encode.swift:1060-1205) -------------------------------
// `pure` = the block emits nothing but (a share of) an EOB run: AC first: every |c| >> al == 0 in the band;
// AC refine: no coefficient becomes newly significant.
__device__ __forceinline__ int16_t coef_at(const int16_t *pl, int ux, uint32_t blk, int z) { return pl[(size_t) blk * 64 + z]; }

template <class Sink>
__device__ void encode_interval_ac(const EncParams &P, uint32_t img, uint32_t e, Sink &sink)
{
    const int16_t *pl = P.plane[0] + (size_t) img * P.image_stride[0];
    const int      ac = P.ac[0], al = P.al, lo = P.band_lo, hi = P.band_hi;
    const uint32_t b0 = e * P.blocks_per_interval, b1 = min(P.S, b0 + P.blocks_per_interval);
    const bool     refine = P.kind == 4;
    const int      mask = (int) (short) (uint16_t) (0xffffu << (al + 1));

    auto is_pure = [&](uint32_t blk) {
        for (int z = lo; z < hi; ++z) {
            const int c = coef_at(pl, 0, blk, z), mag = c < 0 ? -c : c;
            if (!refine) {
                if ((mag >> al) != 0) return false;
            } else if ((mag & mask) == 0 && ((mag & ~mask) >> al) != 0)
                return false;
        }
        return true;
    };
    auto emit_eob = [&](int run) {
        const int binade = 31 - __clz(run);
        sink.symbol(ac, binade << 4, (uint32_t) (run & ~(1 << binade)), binade);
    };
    // correction bits of coefficients z0..hi-1 of a block (AC refine only): one bit per already-significant coefficient
    auto emit_corrections = [&](uint32_t blk, int z0) {
        for (int z = z0; z < hi; ++z) {
            const int c = coef_at(pl, 0, blk, z), mag = c < 0 ? -c : c;
            if ((mag & mask) != 0) sink.raw((uint32_t) (((mag & ~mask) >> al) & 1), 1);
        }
    };

    uint32_t blk = b0;
    while (blk < b1) {
        // ---- the block's run symbols -------------------------------------------------------------------------
        int  zeroes = 0;
        int  tail_from = lo;      // first coefficient not yet covered by an emitted symbol's correction bits
        bool has_tail_bits = false;
        for (int z = lo; z < hi; ++z) {
            const int c = coef_at(pl, 0, blk, z), mag = c < 0 ? -c : c, sign = c < 0 ? -1 : 1;
            if (!refine) {
                const int high = sign * (mag >> al);
                if (high == 0) {
                    ++zeroes;
                    continue;
                }
                for (int i = 0; i < zeroes / 16; ++i) sink.symbol(ac, 0xf0, 0, 0);
                int      binade;
                uint32_t tail;
                compact16(high, binade, tail);
                sink.symbol(ac, ((zeroes % 16) << 4) | binade, tail, binade);
                zeroes = 0;
            } else {
                const int product = mag & mask, low = sign * ((mag & ~mask) >> al);
                if (product != 0) {
                    has_tail_bits = true;
                    continue;
                }
                if (low == 0) {
                    ++zeroes;
                    continue;
                }
                // newly significant: ZRLs (each followed by the correction bits staged during its 16 zeros), then
                // (zeroes % 16, low) followed by the remaining staged bits -- encode.swift:1143-1164
                int zseen = 0, zz = tail_from;
                for (int i = 0; i < zeroes / 16; ++i) {
                    sink.symbol(ac, 0xf0, 0, 0);
                    for (; zz < z; ++zz) {
                        const int c2 = coef_at(pl, 0, blk, zz), m2 = c2 < 0 ? -c2 : c2;
                        if ((m2 & mask) != 0) sink.raw((uint32_t) (((m2 & ~mask) >> al) & 1), 1);
                        else if (++zseen % 16 == 0) {
                            ++zz;
                            break;
                        }
                    }
                }
                int      binade;
                uint32_t tail;
                compact16(low, binade, tail);
                sink.symbol(ac, ((zeroes % 16) << 4) | binade, tail, binade);
                for (; zz < z; ++zz) {
                    const int c2 = coef_at(pl, 0, blk, zz), m2 = c2 < 0 ? -c2 : c2;
                    if ((m2 & mask) != 0) sink.raw((uint32_t) (((m2 & ~mask) >> al) & 1), 1);
                }
                zeroes = 0;
                tail_from = z + 1;
                has_tail_bits = false;
            }
        }
        if (refine) {
            // does the tail (tail_from ..< hi) hold correction bits?
            has_tail_bits = false;
            for (int z = tail_from; z < hi; ++z) {
                const int c = coef_at(pl, 0, blk, z), mag = c < 0 ? -c : c;
                if ((mag & mask) != 0) has_tail_bits = true;
            }
        }
        if (!(zeroes > 0 || (refine && has_tail_bits))) {
            ++blk;
            continue;
        }
        // ---- an EOB run starts here: absorb the following pure blocks (<= 4096 in total) -------------------------
        uint32_t end = blk + 1;
        while (end < b1 && end - blk < 4096 && is_pure(end)) {
            // a pure block joins only if it would emit an EOB itself: zeroes > 0 or correction bits present.
            // (a pure block always has hi - lo >= 1 coefficients that are either zero or significant.)
            ++end;
        }
        emit_eob((int) (end - blk));
        if (refine) {
            emit_corrections(blk, tail_from);
            for (uint32_t k = blk + 1; k < end; ++k) emit_corrections(k, lo);
        }
        blk = end;
    }
}

// ---- progressive AC scans, one thread per block -----------------------------------------------------------------------------------
// What ties the blocks of an interval together is only the end-of-band run: a block whose symbols leave trailing zeros (or, in a
// refinement scan, trailing correction bits) opens a run that absorbs the following `pure` blocks, 4096 blocks at most
// (encode.swift:1101-1117, 1166-1198).  Whether a block is pure, and whether it opens a run, depends on the block alone; who heads
// a run and how long it is follows from the positions of the non-pure blocks around it -- two scans over one flag per block:
//   k_enc_acp_flags   thread per block: PURE (emits nothing but its share of a run), OPENS (ends with something for an EOB);
//   k_enc_acp_chain   CTA per image: last non-pure block at or before i, next non-pure block after i  ->  info[i] = run length
//                     if block i heads a run (its offset in the chain of pure blocks is a multiple of 4096) | ABSORBED;
//   then the histogram / length / emit passes of the sequential scans run one thread per block (encode_block_ac): its own symbols
//   unless absorbed, the EOBn symbol if it heads a run, its (tail) correction bits -- in exactly the order the serial walk emits.
constexpr uint32_t ACF_PURE = 1u, ACF_OPENS = 2u, ACI_ABSORBED = 0x80000000u;

struct NullSink {
    __device__ __forceinline__ void symbol(int, int, uint32_t, int) {}
    __device__ __forceinline__ void raw(uint32_t, int) {}
};

// info: 0 while the flags are being computed.  Returns the block's flags.
template <class Sink>
__device__ __forceinline__ uint32_t encode_block_ac(const EncParams &P, const AcView pl, uint32_t info, Sink &sink)
{
    const int      ac = P.ac[0], al = P.al, lo = P.band_lo, hi = P.band_hi;
    const bool     refine = P.kind == 4;
    const int      mask = (int) (short) (uint16_t) (0xffffu << (al + 1));
    const bool     absorbed = (info & ACI_ABSORBED) != 0u;
    const int      run = (int) (info & 0xffffu);
    int            zeroes = 0, tail_from = lo;
    bool           pure = true;
    if (!absorbed) {
        for (int z = lo; z < hi; ++z) {
            const int c = pl[z], mag = c < 0 ? -c : c, sign = c < 0 ? -1 : 1;
            if (!refine) {  // encode.swift:1076-1099
                const int high = sign * (mag >> al);
                if (high == 0) {
                    ++zeroes;
                    continue;
                }
                pure = false;
                for (int i = 0; i < zeroes / 16; ++i) sink.symbol(ac, 0xf0, 0, 0);
                int      binade;
                uint32_t tail;
                compact16(high, binade, tail);
                sink.symbol(ac, ((zeroes % 16) << 4) | binade, tail, binade);
                zeroes = 0;
            } else {  // encode.swift:1130-1164
                const int product = mag & mask, low = sign * ((mag & ~mask) >> al);
                if (product != 0) continue;
                if (low == 0) {
                    ++zeroes;
                    continue;
                }
                pure = false;
                // newly significant: ZRLs (each followed by the correction bits staged during its 16 zeros), then (zeroes % 16, low)
                // followed by the remaining staged bits
                int zseen = 0, zz = tail_from;
                for (int i = 0; i < zeroes / 16; ++i) {
                    sink.symbol(ac, 0xf0, 0, 0);
                    for (; zz < z; ++zz) {
                        const int c2 = pl[zz], m2 = c2 < 0 ? -c2 : c2;
                        if ((m2 & mask) != 0) sink.raw((uint32_t) (((m2 & ~mask) >> al) & 1), 1);
                        else if (++zseen % 16 == 0) {
                            ++zz;
                            break;
                        }
                    }
                }
                int      binade;
                uint32_t tail;
                compact16(low, binade, tail);
                sink.symbol(ac, ((zeroes % 16) << 4) | binade, tail, binade);
                for (; zz < z; ++zz) {
                    const int c2 = pl[zz], m2 = c2 < 0 ? -c2 : c2;
                    if ((m2 & mask) != 0) sink.raw((uint32_t) (((m2 & ~mask) >> al) & 1), 1);
                }
                zeroes = 0;
                tail_from = z + 1;
            }
        }
    }
    if (run > 0) {
        const int binade = 31 - __clz(run);
        sink.symbol(ac, binade << 4, (uint32_t) (run & ~(1 << binade)), binade);
    }
    bool tail_bits = false;
    if (refine)  // the correction bits after the block's last symbol (all of them for an absorbed block)
        for (int z = tail_from; z < hi; ++z) {
            const int c = pl[z], mag = c < 0 ? -c : c;
            if ((mag & mask) != 0) {
                tail_bits = true;
                sink.raw((uint32_t) (((mag & ~mask) >> al) & 1), 1);
            }
        }
    return (pure ? ACF_PURE : 0u) | ((zeroes > 0 || tail_bits) ? ACF_OPENS : 0u);
}

__global__ void __launch_bounds__(256) k_enc_acp_flags(const __grid_constant__ EncParams P)
{
    __shared__ uint32_t tile[32 * 256];
    const uint32_t img = blockIdx.y;
    NullSink       sink;
    for (uint32_t s0 = blockIdx.x * blockDim.x; s0 < P.S; s0 += gridDim.x * blockDim.x) {  // (whole warps stay in the loop: stage_block_ac syncs them)
        const uint32_t s = s0 + threadIdx.x;
        const AcView   v = stage_block_ac(P, img, s, s < P.S, tile);
        if (s < P.S) P.ac_info[(size_t) img * P.S + s] = encode_block_ac(P, v, 0u, sink);
        __syncwarp();
    }
}

// one CTA per image; interval by interval, tiles of 1024 blocks forwards (last non-pure block at or before i) and then backwards
// (next non-pure block after i), carries in shared memory.  blk_bits serves as scratch for the forward result.
__global__ void __launch_bounds__(1024) k_enc_acp_chain(const __grid_constant__ EncParams P)
{
    __shared__ int32_t warp_v[32];
    __shared__ int32_t carry;
    const uint32_t img = blockIdx.x;
    uint32_t      *info = P.ac_info + (size_t) img * P.S;
    int32_t       *last_np = reinterpret_cast<int32_t *>(P.blk_bits + (size_t) img * P.S);
    const int      lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (uint32_t e = 0; e < P.n_intervals; ++e) {
        const int32_t lo = (int32_t) (e * P.blocks_per_interval), hi = (int32_t) min(P.S, (uint32_t) lo + P.blocks_per_interval);
        // ---- forwards: inclusive max-scan of (i if block i is not pure else lo - 1)
        if (threadIdx.x == 0) carry = lo - 1;
        __syncthreads();
        for (int32_t base = lo; base < hi; base += 1024) {
            const int32_t i = base + (int32_t) threadIdx.x;
            int32_t       x = (i < hi && !(info[i] & ACF_PURE)) ? i : lo - 1;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int32_t y = __shfl_up_sync(0xffffffffu, x, d);
                if (lane >= d) x = max(x, y);
            }
            if (lane == 31) warp_v[wid] = x;
            __syncthreads();
            if (wid == 0) {
                int32_t w = warp_v[lane];
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const int32_t y = __shfl_up_sync(0xffffffffu, w, d);
                    if (lane >= d) w = max(w, y);
                }
                warp_v[lane] = w;
            }
            __syncthreads();
            const int32_t c = carry;
            const int32_t v = max(max(x, wid ? warp_v[wid - 1] : lo - 1), c);
            if (i < hi) last_np[i] = v;
            __syncthreads();
            if (threadIdx.x == 1023) carry = max(c, warp_v[31]);
            __syncthreads();
        }
        // ---- backwards: exclusive min-scan of (i if block i is not pure else hi) from the right, then the verdict per block
        if (threadIdx.x == 0) carry = hi;
        __syncthreads();
        const int32_t n_tiles = (hi - lo + 1023) / 1024;
        for (int32_t t = n_tiles - 1; t >= 0; --t) {
            const int32_t i = lo + t * 1024 + 1023 - (int32_t) threadIdx.x;  // thread 0 takes the tile's last block
            const uint32_t f = i < hi ? info[i] : ACF_PURE;
            const int32_t  own = (i < hi && !(f & ACF_PURE)) ? i : hi;
            int32_t        x = own;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int32_t y = __shfl_up_sync(0xffffffffu, x, d);
                if (lane >= d) x = min(x, y);
            }
            if (lane == 31) warp_v[wid] = x;
            __syncthreads();
            if (wid == 0) {
                int32_t w = warp_v[lane];
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const int32_t y = __shfl_up_sync(0xffffffffu, w, d);
                    if (lane >= d) w = min(w, y);
                }
                warp_v[lane] = w;
            }
            __syncthreads();
            const int32_t c = carry;
            // exclusive: the blocks strictly after i = the threads before this one in the reversed tile, and the tiles to the right
            int32_t after = __shfl_up_sync(0xffffffffu, x, 1);
            if (lane == 0) after = hi;
            const int32_t next_np = min(min(after, wid ? warp_v[wid - 1] : hi), c);
            if (i < hi) {
                uint32_t out;
                if (!(f & ACF_PURE)) {
                    out = (f & ACF_OPENS) ? (uint32_t) min(4096, next_np - i) : 0u;  // heads a run iff it ends with something for an EOB
                } else {
                    const int32_t j = last_np[i];  // last non-pure block before i (lo - 1: none)
                    const int32_t head = (j >= lo && (info[j] & ACF_OPENS)) ? j : j + 1;
                    const int32_t off = i - head;
                    out = (off % 4096 == 0) ? (uint32_t) min(4096, next_np - i) : ACI_ABSORBED;
                }
                // (the flags of block i are read by the pure blocks after it: info[] is rewritten only after the whole interval is judged)
                last_np[i] = (int32_t) out;
            }
            __syncthreads();
            if (threadIdx.x == 1023) carry = min(c, warp_v[31]);
            __syncthreads();
        }
        for (int32_t i = lo + (int32_t) threadIdx.x; i < hi; i += 1024) info[i] = (uint32_t) last_np[i];
        __syncthreads();
    }
}

struct GlobalHistSink {
    uint32_t *h;
    __device__ __forceinline__ void symbol(int table, int sym, uint32_t, int) { atomicAdd(&h[table * 256 + sym], 1u); }
    __device__ __forceinline__ void raw(uint32_t, int) {}
};

template <int PASS>  // 0 histogram, 1 bit count, 2 emit
__global__ void __launch_bounds__(32) k_enc_ac(const __grid_constant__ EncParams P)
{
    const uint32_t img = blockIdx.y, e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= P.n_intervals) return;
    uint64_t *ivl = P.ivl + (size_t) img * (P.n_intervals + 1);
    if (PASS == 0) {
        GlobalHistSink sink{P.hist + (size_t) img * 8 * 256};
        encode_interval_ac(P, img, e, sink);
    } else if (PASS == 1) {
        LenSink sink{P.codes + (size_t) img * 8 * 256, 0u};
        encode_interval_ac(P, img, e, sink);
        ivl[e] = sink.bits;
    } else {
        EmitSink sink;
        sink.codes = P.codes + (size_t) img * 8 * 256;
        sink.words = reinterpret_cast<uint32_t *>(reinterpret_cast<uint8_t *>(P.raw) + (size_t) img * P.raw_stride);
        sink.begin(ivl[e] * 8);
        encode_interval_ac(P, img, e, sink);
        sink.pad_to_byte();
        sink.finish();
    }
}

__global__ void k_enc_ivl_offsets(const __grid_constant__ EncParams P)
{
    const uint32_t img = blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t      *ivl = P.ivl + (size_t) img * (P.n_intervals + 1);
    uint64_t       off = 0;
    for (uint32_t e = 0; e < P.n_intervals; ++e) {
        const uint64_t nbits = ivl[e];
        ivl[e] = off;
        off += (nbits + 7) >> 3;
    }
    ivl[P.n_intervals] = off;
}

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace

// scratch slots 10 (work buffers) and 11 (raw stream) belong to this file
int jpeg_huffman_encode_scan(jpeg_sm100_ctx *ctx, const jpeg_sm100_scan_desc *scan, const jpeg_sm100_dev_spectral *sp,
                             uint64_t interval_mcus, jpeg_sm100_huff_table *tables_out, uint8_t *d_ecs,
                             uint64_t ecs_image_stride, uint64_t *d_ecs_len, uint64_t *h_needed)
{
    if (!scan || !sp || !tables_out || !d_ecs || !d_ecs_len) return JPEG_SM100_ERR_INVALID_ARGUMENT;
    if (scan->n_comp < 1 || scan->n_comp > 4) return JPEG_SM100_ERR_INVALID_ARGUMENT;
    if (interval_mcus && jpeg_virtual_scan_needed(scan, sp, interval_mcus)) {
        // an interval that is not whole MCU rows (ITU-T T.81 E.1.4): the same MCUs in the same order on a virtual grid (remap.cu)
        JpegVirtualScan v;
        J_TRY(jpeg_virtual_scan_setup(ctx, scan, sp, interval_mcus, &v));
        J_TRY(jpeg_virtual_scan_copy(ctx, &v, true));
        return jpeg_huffman_encode_scan(ctx, &v.scan, &v.sp, interval_mcus, tables_out, d_ecs, ecs_image_stride, d_ecs_len, h_needed);
    }
    if (!(scan->band_lo >= 0 && scan->band_lo < scan->band_hi && scan->band_hi <= 64)) return JPEG_SM100_ERR_INVALID_ARGUMENT;
    const uint32_t n_images = sp->n_images;
    if (n_images == 0) return JPEG_SM100_OK;
    const bool initial = scan->bit_hi < 0;
    EncParams  P;
    memset(&P, 0, sizeof P);
    if (scan->band_lo == 0 && scan->band_hi == 64) {
        if (!initial) return JPEG_SM100_ERR_PRECONDITION;
        P.kind = 0;
    } else if (scan->band_lo == 0 && scan->band_hi == 1)
        P.kind = initial ? 1 : 2;
    else
        P.kind = initial ? 3 : 4;
    if (P.kind >= 3 && scan->n_comp != 1) return JPEG_SM100_ERR_PRECONDITION;  // encode.swift:1595
    if (scan->bit_lo < 0 || scan->bit_lo > 14) return JPEG_SM100_ERR_INVALID_ARGUMENT;
    P.band_lo = scan->band_lo;
    P.band_hi = scan->band_hi;
    P.al = scan->bit_lo;
    P.n_comp = scan->n_comp;
    const bool interleaved = scan->n_comp > 1;
    int        volume = 0;
    for (int c = 0; c < scan->n_comp; ++c) {
        const int p = scan->comp[c].plane;
        if (p < 0 || p >= (int) sp->n_planes) return JPEG_SM100_ERR_PRECONDITION;  // encode.swift:1223, 1239
        if (scan->comp[c].dc < 0 || scan->comp[c].dc > 3 || scan->comp[c].ac < 0 || scan->comp[c].ac > 3)
            return JPEG_SM100_ERR_INVALID_ARGUMENT;
        P.plane[c] = sp->plane[p].coef;
        P.image_stride[c] = sp->plane[p].image_stride;
        P.ux[c] = sp->plane[p].units_x;
        P.uy[c] = sp->plane[p].units_y;
        P.fx[c] = interleaved ? scan->comp[c].factor_x : 1;
        P.fy[c] = interleaved ? scan->comp[c].factor_y : 1;
        P.dc[c] = scan->comp[c].dc;
        P.ac[c] = 4 + scan->comp[c].ac;
        if (P.fx[c] < 1 || P.fy[c] < 1 || P.fx[c] > 4 || P.fy[c] > 4) return JPEG_SM100_ERR_INVALID_ARGUMENT;
        const int first = volume;
        for (int dy = 0; dy < P.fy[c]; ++dy)
            for (int dx = 0; dx < P.fx[c]; ++dx) {
                if (volume >= 12) return JPEG_SM100_ERR_INVALID_ARGUMENT;
                P.blk_comp[volume] = (uint8_t) c;
                P.blk_dx[volume] = (uint8_t) dx;
                P.blk_dy[volume] = (uint8_t) dy;
                P.blk_first[volume] = volume == first;
                P.blk_count[volume] = (uint8_t) (P.fx[c] * P.fy[c]);
                ++volume;
            }
    }
    P.mcu_blocks = volume;
    P.W = interleaved ? scan->blocks_x : P.ux[0];
    P.H = interleaved ? scan->blocks_y : P.uy[0];
    if (P.W <= 0 || P.H < 0) return JPEG_SM100_ERR_INVALID_ARGUMENT;
    const uint64_t S64 = (uint64_t) P.W * P.H * volume;
    if (S64 > 0x7fffffffull) return JPEG_SM100_ERR_UNSUPPORTED;
    P.S = (uint32_t) S64;
    uint64_t rows_per_interval = P.H > 0 ? (uint64_t) P.H : 1;
    if (interval_mcus) {
        if (interval_mcus % (uint64_t) P.W) return JPEG_SM100_ERR_UNSUPPORTED;  // whole MCU rows only (header comment)
        rows_per_interval = interval_mcus / P.W;
    }
    P.blocks_per_interval = (uint32_t) std::min<uint64_t>(rows_per_interval * P.W * volume, 0x7fffffffull);
    if (P.blocks_per_interval == 0) P.blocks_per_interval = 1;
    P.n_intervals = P.S ? (P.S + P.blocks_per_interval - 1) / P.blocks_per_interval : 1;
    // worst case per block: 64 x (16 + 16) bits; intervals must stay below 2^32 bits for the 32-bit in-interval offsets
    if ((uint64_t) P.blocks_per_interval * 64 * 32 > 0xffffffffull) {
        // tighter, still safe bound is not available before the length pass; cap the interval size instead
        if ((uint64_t) P.blocks_per_interval > (1ull << 21)) return JPEG_SM100_ERR_UNSUPPORTED;
    }

    // ---- work buffers -------------------------------------------------------------------------------------------
    const size_t hist_bytes = (size_t) n_images * 8 * 256 * 4;
    const size_t code_bytes = hist_bytes;
    const size_t bits_bytes = align_up((size_t) n_images * std::max<uint32_t>(P.S, 1) * 4, 256);
    const size_t ivl_bytes = align_up((size_t) n_images * (P.n_intervals + 1) * 8, 256);
    // progressive AC scans: one thread per block (k_enc_acp_*); JPEG_SM100_ENC_AC=seq keeps the one-thread-per-interval kernels (A/B)
    const char *ac_env = getenv("JPEG_SM100_ENC_AC");
    const bool  ac_blocks = P.kind >= 3 && !(ac_env && strcmp(ac_env, "seq") == 0);
    void        *work = nullptr;
    J_TRY(scratch_reserve(ctx, 10, hist_bytes + code_bytes + bits_bytes * (ac_blocks ? 2 : 1) + ivl_bytes + 1024, &work));
    uint8_t *wp = reinterpret_cast<uint8_t *>(work);
    P.hist = reinterpret_cast<uint32_t *>(wp);
    P.codes = reinterpret_cast<uint32_t *>(wp + hist_bytes);
    P.blk_bits = reinterpret_cast<uint32_t *>(wp + hist_bytes + code_bytes);
    P.ivl = reinterpret_cast<uint64_t *>(wp + hist_bytes + code_bytes + bits_bytes);
    P.ac_info = ac_blocks ? reinterpret_cast<uint32_t *>(wp + hist_bytes + code_bytes + bits_bytes + ivl_bytes) : nullptr;
    P.out = d_ecs;
    P.out_stride = ecs_image_stride;
    P.out_len = d_ecs_len;
    CU_TRY(ctx, cudaMemsetAsync(P.hist, 0, hist_bytes, ctx->stream));
    CU_TRY(ctx, cudaMemsetAsync(P.ivl, 0, ivl_bytes, ctx->stream));

    const dim3 grid_blocks(std::max<uint32_t>(1, std::min<uint32_t>((P.S + 255) / 256, (uint32_t) ctx->sm_count * 8)), n_images);
    const dim3 grid_ivl((P.n_intervals + 31) / 32, n_images);
    const size_t ac_smem = ac_blocks ? (size_t) 32 * 256 * 4 : 0;  // the transposed block tile of stage_block_ac

    // JPEG_SM100_TRACE: wall-clock split of the call (a diagnostic: adds stream synchronisations)
    static const bool trace = getenv("JPEG_SM100_TRACE") != nullptr;
    timespec          tr_t[6];
    int               tr_n = 0;
    auto              tick = [&](bool sync) {
        if (!trace) return;
        if (sync) cudaStreamSynchronize(ctx->stream);
        clock_gettime(CLOCK_MONOTONIC, &tr_t[tr_n++]);
    };
    tick(true);
    // ---- pass 1: statistics -> optimal tables (host) ----------------------------------------------------------
    std::vector<uint32_t> h_hist((size_t) n_images * 8 * 256, 0u), h_codes((size_t) n_images * 8 * 256, 0u);
    const bool need_dc = P.kind == 0 || P.kind == 1, need_ac = P.kind == 0 || P.kind == 3 || P.kind == 4;
    if (P.kind != 2) {
        if (P.S) {
            if (ac_blocks) {
                k_enc_acp_flags<<<grid_blocks, 256, 0, ctx->stream>>>(P);
                LAUNCH_CHECK(ctx);
                k_enc_acp_chain<<<n_images, 1024, 0, ctx->stream>>>(P);
                LAUNCH_CHECK(ctx);
            }
            if (P.kind <= 1 || ac_blocks) k_enc_hist<<<grid_blocks, 256, ac_smem, ctx->stream>>>(P);
            else k_enc_ac<0><<<grid_ivl, 32, 0, ctx->stream>>>(P);
            LAUNCH_CHECK(ctx);
        }
        CU_TRY(ctx, cudaMemcpyAsync(h_hist.data(), P.hist, hist_bytes, cudaMemcpyDeviceToHost, ctx->stream));
        CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    }
    tick(false);  // [1] statistics on the device, read back
    // one optimal table per (image, slot): independent heaps of <= 257 leaves (~40 us each) -- a batch spreads them over host threads
    {
        bool used[8] = {false, false, false, false, false, false, false, false};
        for (int c = 0; c < scan->n_comp; ++c) {
            if (need_dc) used[P.dc[c]] = true;
            if (need_ac) used[P.ac[c]] = true;
        }
        std::atomic<int> failed{0};
        auto build = [&](uint32_t i0, uint32_t i1) {
            for (uint32_t i = i0; i < i1; ++i) {
                jpeg_sm100_huff_table *t = tables_out + (size_t) i * 8;
                memset(t, 0, sizeof(jpeg_sm100_huff_table) * 8);
                for (int k = 0; k < 8; ++k) {
                    if (!used[k]) continue;
                    const uint32_t *f = h_hist.data() + ((size_t) i * 8 + k) * 256;
                    bool any = false;
                    for (int v = 0; v < 256; ++v) any |= f[v] != 0;
                    if (!any) {  // encode.swift:705: all-zero frequencies trap
                        failed.store(1);
                        continue;
                    }
                    huffman_from_frequencies(f, t[k]);
                    canonical_codes(t[k], h_codes.data() + ((size_t) i * 8 + k) * 256);
                }
            }
        };
        const uint32_t hw = std::max(1u, std::thread::hardware_concurrency());
        const uint32_t n_thr = std::min<uint32_t>(std::min<uint32_t>(hw, 16u), n_images / 8u);
        if (n_thr <= 1) build(0, n_images);
        else {
            std::vector<std::thread> pool;
            const uint32_t           per = (n_images + n_thr - 1) / n_thr;
            uint32_t                 next = 0;
            try {
                for (; next < n_images; next += per) pool.emplace_back(build, next, std::min(n_images, next + per));
            } catch (...) {  // no more threads to be had (std::system_error must not cross the C ABI): the rest on this one
            }
            if (next < n_images) build(next, n_images);
            for (auto &th : pool) th.join();
        }
        if (failed.load()) return JPEG_SM100_ERR_PRECONDITION;
    }
    tick(false);  // [2] tables on the host
    CU_TRY(ctx, cudaMemcpyAsync(const_cast<uint32_t *>(P.codes), h_codes.data(), code_bytes, cudaMemcpyHostToDevice, ctx->stream));

    // ---- pass 2: lengths and offsets ----------------------------------------------------------------------------
    if (P.S) {
        if (P.kind <= 2 || ac_blocks) {
            k_enc_len<<<grid_blocks, 256, ac_smem, ctx->stream>>>(P);
            LAUNCH_CHECK(ctx);
            if (P.n_intervals <= 65535u) {
                k_enc_scan_bits_ivl<<<dim3(P.n_intervals, n_images), 1024, 0, ctx->stream>>>(P);
                LAUNCH_CHECK(ctx);
                k_enc_ivl_offsets<<<n_images, 1, 0, ctx->stream>>>(P);
            } else
                k_enc_scan_bits<<<n_images, 1024, 0, ctx->stream>>>(P);
            LAUNCH_CHECK(ctx);
        } else {
            k_enc_ac<1><<<grid_ivl, 32, 0, ctx->stream>>>(P);
            LAUNCH_CHECK(ctx);
            k_enc_ivl_offsets<<<n_images, 1, 0, ctx->stream>>>(P);
            LAUNCH_CHECK(ctx);
        }
    }
    // raw stream size: read back the per-image totals (the codes copy above is done by then as well)
    std::vector<uint64_t> h_ivl((size_t) n_images * (P.n_intervals + 1));
    CU_TRY(ctx, cudaMemcpyAsync(h_ivl.data(), P.ivl, h_ivl.size() * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    uint64_t max_raw = 0;
    tick(false);  // [3] lengths + offsets on the device, read back
    for (uint32_t i = 0; i < n_images; ++i) max_raw = std::max(max_raw, h_ivl[(size_t) i * (P.n_intervals + 1) + P.n_intervals]);
    P.raw_stride = align_up(max_raw + 8, 256);
    void *raw = nullptr;
    J_TRY(scratch_reserve(ctx, 11, P.raw_stride * n_images + 256, &raw));
    P.raw = reinterpret_cast<uint32_t *>(raw);
    CU_TRY(ctx, cudaMemsetAsync(raw, 0, P.raw_stride * n_images, ctx->stream));

    // ---- pass 3: emit, stuff --------------------------------------------------------------------------------------
    if (P.S) {
        if (P.kind <= 2 || ac_blocks) k_enc_emit<<<grid_blocks, 256, ac_smem, ctx->stream>>>(P);
        else k_enc_ac<2><<<grid_ivl, 32, 0, ctx->stream>>>(P);
        LAUNCH_CHECK(ctx);
    }
    static const bool serial_stuff = [] { const char *e = getenv("JPEG_SM100_STUFF"); return e && strcmp(e, "serial") == 0; }();
    if (serial_stuff) {  // first generation (one CTA per image), kept for A/B validation
        k_enc_stuff<<<n_images, 1024, 0, ctx->stream>>>(P);
        LAUNCH_CHECK(ctx);
    } else {
        const uint32_t tiles_max = (uint32_t) std::max<uint64_t>(1, (max_raw + STUFF_TILE - 1) / STUFF_TILE);
        void          *tf = nullptr;
        J_TRY(scratch_reserve(ctx, 14, (size_t) n_images * tiles_max * 4 + 256, &tf));
        const dim3 grid_t(tiles_max, n_images);
        k_enc_stuff_count<<<grid_t, STUFF_THREADS, 0, ctx->stream>>>(P, tiles_max, reinterpret_cast<uint32_t *>(tf));
        LAUNCH_CHECK(ctx);
        k_enc_stuff_scatter<<<grid_t, STUFF_THREADS, 0, ctx->stream>>>(P, tiles_max, reinterpret_cast<const uint32_t *>(tf));
        LAUNCH_CHECK(ctx);
    }
    if (trace) {
        tick(true);  // [4] emit + stuffing
        auto ms = [&](int a, int b) { return (tr_t[b].tv_sec - tr_t[a].tv_sec) * 1e3 + (tr_t[b].tv_nsec - tr_t[a].tv_nsec) * 1e-6; };
        fprintf(stderr, "[jpeg_sm100] encode scan band %d..%d bits %d/%d comps %d images %u: statistics %.3f ms, host tables %.3f, lengths %.3f, emit %.3f\n",
                scan->band_lo, scan->band_hi, scan->bit_hi, scan->bit_lo, scan->n_comp, n_images, ms(0, 1), ms(1, 2), ms(2, 3), ms(3, 4));
    }
    if (h_needed) {
        std::vector<uint64_t> lens(n_images);
        CU_TRY(ctx, cudaMemcpyAsync(lens.data(), d_ecs_len, n_images * 8, cudaMemcpyDeviceToHost, ctx->stream));
        CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        *h_needed = 0;
        for (uint64_t l : lens) *h_needed = std::max(*h_needed, l);
    }
    return JPEG_SM100_OK;
}

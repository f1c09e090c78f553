// common.cuh -- shared host/device plumbing for libjpeg_sm100.so (sm_100a only).
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/jpeg_sm100.h"

#define JPEG_API extern "C" __attribute__((visibility("default")))

struct DeviceBuffer {
    void  *ptr = nullptr;
    size_t cap = 0;
};

// pinned host staging slots: async H2D of small per-call metadata without a host-side stream sync
struct PinnedSlot {
    void       *ptr = nullptr;
    size_t      cap = 0;
    cudaEvent_t done = nullptr;
    bool        in_flight = false;
};

struct jpeg_sm100_ctx {
    int          device = 0;
    cudaStream_t stream = nullptr;
    bool         owns_stream = false;
    int          sm_count = 148;
    uint64_t     launches = 0;
    uint64_t     h2d_bytes = 0, d2h_bytes = 0;  // layer A bookkeeping (jpeg_sm100_transfer_counts)
    std::string  last_error;
    // grow-only scratch used by layer A (host-buffer entry points)
    DeviceBuffer scratch[20];
    PinnedSlot   pinned[4];
    int          pinned_next = 0;
    // copy streams + events of the chunked host-buffer pipeline (jpeg_sm100_decode_batch_rgb8)
    cudaStream_t copy_in = nullptr, copy_out = nullptr;
    std::vector<cudaEvent_t> events;
    cudaEvent_t  wait_event = nullptr;  // blocking-sync event (JPEG_SM100_WAIT=block)
    // average bytes per restart interval of the next entropy-decode call, when the caller knows it (0: unknown); consumed by K3
    uint64_t hint_interval_bytes = 0;
    int      par_smem_ac = 0;   // same for its progressive AC-first instantiation
    size_t   par_smem_set = 0;  // largest dynamic shared-memory size k_decode_par has been opted into on this device
    bool     fused_smem_set = false;  // k_idct_rgb420 opted into its dynamic shared memory on this device
    bool     idct_smem_set[2] = {false, false};  // k_idct_tma<u8 / u16> opted into > 48 KB of dynamic shared memory (per device, so per ctx)
    // cuTensorMapEncodeTiled, resolved through the runtime (no link-time libcuda dependency)
    void *encode_tiled = nullptr;
};

// ---- error handling ------------------------------------------------------------------------------------------
static inline int jpeg_cuda_fail(jpeg_sm100_ctx *ctx, cudaError_t e, const char *what, const char *file, int line)
{
    if (ctx) {
        char buf[512];
        snprintf(buf, sizeof buf, "%s: %s (%s:%d)", what, cudaGetErrorString(e), file, line);
        ctx->last_error = buf;
    }
    return JPEG_SM100_ERR_CUDA;
}
#define CU_TRY(ctx, expr)                                                                                            \
    do {                                                                                                             \
        cudaError_t _e = (expr);                                                                                     \
        if (_e != cudaSuccess) return jpeg_cuda_fail((ctx), _e, #expr, __FILE__, __LINE__);                          \
    } while (0)
#define J_TRY(expr)                                                                                                  \
    do {                                                                                                             \
        int _r = (expr);                                                                                             \
        if (_r != JPEG_SM100_OK) return _r;                                                                          \
    } while (0)
#define LAUNCH_CHECK(ctx)                                                                                            \
    do {                                                                                                             \
        (ctx)->launches += 1;                                                                                        \
        CU_TRY((ctx), cudaGetLastError());                                                                           \
    } while (0)

// host <-> device copies of layer A, counted
static inline cudaError_t copy_h2d(jpeg_sm100_ctx *ctx, void *dst, const void *src, size_t n, cudaStream_t st)
{
    ctx->h2d_bytes += n;
    return cudaMemcpyAsync(dst, src, n, cudaMemcpyHostToDevice, st);
}
static inline cudaError_t copy_d2h(jpeg_sm100_ctx *ctx, void *dst, const void *src, size_t n, cudaStream_t st)
{
    ctx->d2h_bytes += n;
    return cudaMemcpyAsync(dst, src, n, cudaMemcpyDeviceToHost, st);
}

static inline int scratch_reserve(jpeg_sm100_ctx *ctx, int slot, size_t bytes, void **out)
{
    DeviceBuffer &b = ctx->scratch[slot];
    if (bytes > b.cap) {
        if (b.ptr) CU_TRY(ctx, cudaFree(b.ptr));
        b.ptr = nullptr;
        b.cap = 0;
        size_t want = bytes + (bytes >> 3) + 4096;
        CU_TRY(ctx, cudaMalloc(&b.ptr, want));
        b.cap = want;
    }
    *out = b.ptr;
    return JPEG_SM100_OK;
}

// Acquire a pinned staging slot of at least `bytes`; blocks only if that slot's previous copy is still in flight.
static inline int pinned_acquire(jpeg_sm100_ctx *ctx, size_t bytes, void **out, int *slot_out)
{
    const int   si = ctx->pinned_next;
    PinnedSlot &s = ctx->pinned[si];
    ctx->pinned_next = (si + 1) % 4;
    if (s.in_flight) {
        CU_TRY(ctx, cudaEventSynchronize(s.done));
        s.in_flight = false;
    }
    if (!s.done) CU_TRY(ctx, cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
    if (bytes > s.cap) {
        if (s.ptr) CU_TRY(ctx, cudaFreeHost(s.ptr));
        s.ptr = nullptr;
        s.cap = 0;
        CU_TRY(ctx, cudaMallocHost(&s.ptr, bytes + 4096));
        s.cap = bytes + 4096;
    }
    *out = s.ptr;
    *slot_out = si;
    return JPEG_SM100_OK;
}
// call after the async copies that read the slot have been enqueued
static inline int pinned_release(jpeg_sm100_ctx *ctx, int slot)
{
    CU_TRY(ctx, cudaEventRecord(ctx->pinned[slot].done, ctx->stream));
    ctx->pinned[slot].in_flight = true;
    return JPEG_SM100_OK;
}

// ---- restart intervals that are not whole MCU rows: the scan on a virtual MCU grid (remap.cu) ---------------------------------
struct JpegVirtualScan {
    jpeg_sm100_scan_desc    scan;  // the scan on the virtual grid (blocks_x = gcd(interval, MCUs), planes in scan order)
    jpeg_sm100_dev_spectral sp;    // virtual planes (context scratch)
    alignas(8) unsigned char params[320];
};
bool jpeg_virtual_scan_needed(const jpeg_sm100_scan_desc *scan, const jpeg_sm100_dev_spectral *sp, uint64_t interval);
int  jpeg_virtual_scan_setup(jpeg_sm100_ctx *ctx, const jpeg_sm100_scan_desc *scan, const jpeg_sm100_dev_spectral *sp, uint64_t interval,
                             JpegVirtualScan *v);
int  jpeg_virtual_scan_copy(jpeg_sm100_ctx *ctx, const JpegVirtualScan *v, bool to_virtual);

// ---- geometry helpers (decode.swift:1364-1369) -----------------------------------------------------------------
static inline int units_of(int size, int stride) { return size / stride + (size % stride != 0 ? 1 : 0); }

// zig-zag index of (k = horizontal frequency, h = vertical frequency): decode.swift:1289-1298
__host__ __device__ constexpr int zigzag_index(int x, int y)
{
    const int p = (x + y < 8) ? 1 : 0, q = (x + y) & 1;
    const int a = 72 * (p ^ 1), b = 2 * p - 1;
    const int n = b * (x + y) - 14 * p + 15;
    const int t = (n * (n + 1)) >> 1;
    return a + b * t - q * x - (q ^ 1) * y - 1;
}

// (r[k] * r[h]) * (scale * Q[zz(k,h)])  -- decode.swift:3984-4017, evaluated in binary32 exactly as the reference
static inline void modulate_quanta(const uint16_t quanta_zz[64], float scale, float out[64] /* [h*8+k] */)
{
    static const float R[8] = {1.0f,         1.387039845f, 1.306562965f, 1.175875602f,
                               1.0f,         0.785694958f, 0.541196100f, 0.275899379f};
    for (int h = 0; h < 8; ++h)
        for (int k = 0; k < 8; ++k) {
            volatile float hv = R[k] * R[h];
            volatile float row = scale * (float) quanta_zz[zigzag_index(k, h)];
            out[h * 8 + k] = hv * row;
        }
}

// ---- device helpers ---------------------------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t done;
    const uint32_t addr = smem_u32(bar);
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                     "selp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done)
                     : "r"(addr), "r"(parity)
                     : "memory");
    } while (!done);
}
// TMA: 2-D tiled tensor load global -> shared, completion signalled on an mbarrier (SASS: UTMALDG)
__device__ __forceinline__ void tma_load_2d(void *smem_dst, const CUtensorMap *tmap, int c0, int c1, uint64_t *bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(c0), "r"(c1), "r"(smem_u32(bar))
                 : "memory");
}
// TMA: 1-D bulk copy global -> shared (SASS: UBLKCP)
__device__ __forceinline__ void tma_load_1d(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap *tmap)
{
    asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}

// never-contracted binary32 arithmetic (the reference is plain Swift Float: no FMA)
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }

// two s32 -> saturated u8 bytes packed under the low half of c:  d = { c[15:0], sat8(a), sat8(b) }  (SASS: I2IP)
__device__ __forceinline__ uint32_t pack_sat_u8(int a, int b, uint32_t c)
{
    uint32_t d;
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
// four floats -> four bytes [p0 p1 p2 p3] little endian, each trunc-toward-zero then clamped to 0..255
__device__ __forceinline__ uint32_t pack4_u8_trunc(float p0, float p1, float p2, float p3)
{
    const uint32_t hi = pack_sat_u8(__float2int_rz(p3), __float2int_rz(p2), 0u);
    return pack_sat_u8(__float2int_rz(p1), __float2int_rz(p0), hi);
}
#endif

/*
 * jpeg_sm100.h -- C-ABI of libjpeg_sm100.so: the JPEG per-MCU block-transform hot path as hand-written CUDA
 * kernels for sm_100a (NVIDIA B200).
 *
 * This is the drop-in boundary for tayloraswift/jpeg (pure Swift, reference @ 8fe8fda1).  The reference has no
 * compute plugin interface; the seam is the set of stage functions its staged API already exposes.  The Swift
 * host keeps JPEG.Data.{Spectral,Planar,Rectangular}, JPEG.Layout, JPEG.Context and every public signature;
 * the BODIES of the functions cited below are replaced by calls to these symbols (swift/JPEGSM100Shim.swift,
 * INTEGRATION.md).  Citations are to sources/jpeg/<file>.swift:<line> of the reference.
 *
 * Conventions
 *   - plain C: pointers + sizes, no CUDA / torch types.  A "stream" is passed as void* (cudaStream_t).
 *   - every function returns 0 or a negative jpeg_sm100 error code; codes -1..-8 map 1:1 onto the
 *     JPEG.DecodingError cases the hot path can raise (error.swift:495-667).
 *   - Layer A (jpeg_sm100_<stage>): HOST buffers in, HOST buffers out, synchronous.  The caller (Swift) owns all
 *     memory; pointers are only used for the duration of the call.  This is what the shim binds.
 *   - Layer B (jpeg_sm100_dev_<stage>): DEVICE pointers, asynchronous on the ctx stream, batched over n_images
 *     images of identical geometry.  Layer A is implemented on top of layer B with n_images = 1.
 *   - one ctx = one device + one stream; a ctx is not thread-safe, different ctxs are independent.
 *   - there is NO CPU fallback: every entry point fails with JPEG_SM100_ERR_CUDA if no sm_100 device is usable.
 *
 * Memory layouts are the reference's own (decode.swift:1412-1480, 1548-1598, 1650-1718):
 *   Spectral.Plane   int16  coef[64 * (units_x * y + x) + z]          z = zig-zag index
 *   Planar.Plane     uint16 sample[x + 8 * units_x * y]               (uint8 variant: same indexing)
 *   Rectangular      uint16 value[(y * size_x + x) * n_planes + p]
 *   RGB / YCbCr      3 x uint8 per pixel, row-major
 */
#ifndef JPEG_SM100_H
#define JPEG_SM100_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define JPEG_SM100_ABI_VERSION 1

/* ---- status codes ---------------------------------------------------------------------------------------- */
enum {
    JPEG_SM100_OK                          = 0,
    JPEG_SM100_ERR_TRUNCATED_ECS           = -1,  /* DecodingError.truncatedEntropyCodedSegment   decode.swift:2778,2794,2811,2845,2862 */
    JPEG_SM100_ERR_INVALID_COMPOSITE_VALUE = -2,  /* DecodingError.invalidCompositeValue          decode.swift:3109 */
    JPEG_SM100_ERR_INVALID_BLOCK_RUN       = -3,  /* DecodingError.invalidCompositeBlockRun       decode.swift:2951,3284 */
    JPEG_SM100_ERR_UNDEFINED_DC            = -4,  /* DecodingError.undefinedScanHuffmanDCReference decode.swift:2888,3192 */
    JPEG_SM100_ERR_UNDEFINED_AC            = -5,  /* DecodingError.undefinedScanHuffmanACReference decode.swift:2894,3198 */
    JPEG_SM100_ERR_PRECONDITION            = -7,  /* the reference would trap (preconditionFailure / Range2 bounds) */
    JPEG_SM100_ERR_INVALID_HUFFMAN         = -8,  /* ParsingError.invalidHuffmanTable              decode.swift:524 */
    JPEG_SM100_ERR_RESTART_PHASE           = -9,  /* DecodingError.invalidRestartPhase             decode.swift:3931 */
    JPEG_SM100_ERR_ECS_COUNT               = -10, /* batch lexer only: an image's RSTn count differs from the batch's n_ecs - 1 */
    JPEG_SM100_ERR_MISSING_INTERVAL        = -11, /* DecodingError.missingRestartIntervalSegment   decode.swift:3719 */
    JPEG_SM100_ERR_INVALID_ARGUMENT        = -20,
    JPEG_SM100_ERR_UNSUPPORTED             = -21,
    JPEG_SM100_ERR_NO_MEMORY               = -22,
    JPEG_SM100_ERR_CUDA                    = -100 /* device missing / launch failure; see jpeg_sm100_last_cuda_error */
};

#define JPEG_SM100_INTERVAL_NONE UINT64_MAX /* no DRI: interval = Int.max, decode.swift:3708-3720 */
#define JPEG_SM100_BITS_MAX      (-1)       /* scan.bits.upperBound == .max (initial scan) */
/* bits of the `extend` argument of the decode_scan entry points */
#define JPEG_SM100_SCAN_EXTEND   1          /* the reference's `extend` flag (first scan, decode.swift:3214-3236) */
#define JPEG_SM100_SCAN_FRESH    2          /* layer B only: the planes of the scan's components are newly created
                                               Spectral planes (all zero, decode.swift:2241-2256) -- the library clears
                                               them as part of the call, whatever they hold */
#define JPEG_SM100_SCAN_T81      4          /* restart interval e starts at MCU e * interval wherever that falls in its row
                                               (ITU-T T.81 E.1.4), instead of the reference's integer-division ROW placement
                                               (decode.swift:3205-3207), which is only right for whole-row intervals.  Sequential
                                               and DC scans; an extension: the Swift decoder mis-places such files */

/* ---- lifecycle ---------------------------------------------------------------------------------------------- */
typedef struct jpeg_sm100_ctx jpeg_sm100_ctx;

int         jpeg_sm100_abi_version(void);
int         jpeg_sm100_create(int device, jpeg_sm100_ctx **out);                  /* owns a new non-blocking stream */
int         jpeg_sm100_create_on_stream(int device, void *cuda_stream, jpeg_sm100_ctx **out); /* borrows a stream */
void        jpeg_sm100_destroy(jpeg_sm100_ctx *ctx);
int         jpeg_sm100_sync(jpeg_sm100_ctx *ctx);
const char *jpeg_sm100_error_string(int code);
const char *jpeg_sm100_last_cuda_error(jpeg_sm100_ctx *ctx);
/* number of kernels this ctx has launched so far (our own kernels only; memsets/copies are not counted) */
uint64_t    jpeg_sm100_launch_count(jpeg_sm100_ctx *ctx);
/* device-memory helpers so a host without a CUDA binding can drive layer B */
int         jpeg_sm100_malloc(jpeg_sm100_ctx *ctx, size_t bytes, void **dev_ptr);
int         jpeg_sm100_free(jpeg_sm100_ctx *ctx, void *dev_ptr);
int         jpeg_sm100_malloc_host(jpeg_sm100_ctx *ctx, size_t bytes, void **pinned_ptr);
int         jpeg_sm100_free_host(jpeg_sm100_ctx *ctx, void *pinned_ptr);
int         jpeg_sm100_upload(jpeg_sm100_ctx *ctx, void *dev_dst, const void *host_src, size_t bytes);   /* async */
int         jpeg_sm100_download(jpeg_sm100_ctx *ctx, void *host_dst, const void *dev_src, size_t bytes); /* async */
int         jpeg_sm100_memset(jpeg_sm100_ctx *ctx, void *dev_dst, int value, size_t bytes);              /* async */

/* ---- shared descriptors ---------------------------------------------------------------------------------- */

/* the validated DHT payload held by JPEG.Table.Huffman (jpeg.swift:998-1012): BITS + HUFFVAL */
typedef struct {
    int32_t present;      /* 0 = slot is nil */
    uint8_t counts[16];   /* leaves per level 1..16 */
    uint8_t values[256];  /* symbols, level by level */
} jpeg_sm100_huff_table;

/* JPEG.Scan + the geometry Spectral.decode / Spectral.encode read from `self` (decode.swift:3476, encode.swift:1559) */
typedef struct {  /* one JPEG.Scan.Component (jpeg.swift:1135-1160), named so that bindings can address the array elements */
    int32_t plane;               /* index into the planes array, or -1: component without a plane (decode.swift:3251) */
    int32_t factor_x, factor_y;  /* layout.planes[c].component.factor */
    int32_t dc, ac;              /* table slots 0..3 */
} jpeg_sm100_scan_comp;
typedef struct {
    int32_t band_lo, band_hi;  /* scan.band, 0 <= lo < hi <= 64 */
    int32_t bit_lo, bit_hi;    /* scan.bits; bit_hi = JPEG_SM100_BITS_MAX for an initial scan */
    int32_t n_comp;            /* 1..4, in scan-header order */
    jpeg_sm100_scan_comp comp[4];
    int32_t blocks_x, blocks_y;      /* Spectral.blocks: the MCU grid */
} jpeg_sm100_scan_desc;

typedef struct { int16_t  *coef;    int32_t units_x, units_y; } jpeg_sm100_plane_i16;
typedef struct { uint16_t *samples; int32_t units_x, units_y; int32_t factor_x, factor_y; } jpeg_sm100_plane_u16;
typedef struct { uint8_t  *samples; int32_t units_x, units_y; int32_t factor_x, factor_y; } jpeg_sm100_plane_u8;

/* =============================================================================================================
 * LAYER A -- host buffers, synchronous: the functions the Swift shim binds
 * ============================================================================================================= */

/* ---- decode ---- */

/* replaces the body of  Spectral.decode(ecss:interval:scan:tables:extend:)   decode.swift:3476-3551
 * (the scan-kind dispatch and all ten per-kind decoders decode.swift:2880-3445; Huffman LUT decode.swift:1037-1265;
 * composites decode.swift:2773-2872; bitstream jpeg.swift:1873-1916).
 *   ecs_concat / ecs_offsets : the reference's ecss:[[UInt8]] (already unstuffed and split at RSTn by the lexer),
 *                              flattened; n_ecs + 1 offsets.
 *   interval                 : MCUs per restart interval, JPEG_SM100_INTERVAL_NONE if no DRI.  Interval e is placed as the
 *                              reference places it: MCU rows floor(e * interval / W) ..< floor((e + 1) * interval / W), W = blocks_x
 *                              (or units_x for single-component scans), decode.swift:3205-3207 -- exact for intervals that are
 *                              whole rows; for other values see JPEG_SM100_SCAN_T81 above.
 *   extend                   : the reference's `extend` flag (first scan): rows stop silently at the end of data
 *                              (decode.swift:3214-3220).  Planes must already be sized for the final height
 *                              (the host handles DNL before the call); growth beyond it is not performed.
 *   planes                   : in/out, n_planes Spectral.Plane buffers (progressive scans read-modify-write).
 * Quantisation-table binding (decode.swift:3451-3498) stays in Swift: it is pure bookkeeping. */
int jpeg_sm100_decode_scan(jpeg_sm100_ctx *ctx, const jpeg_sm100_scan_desc *scan,
                           const uint8_t *ecs_concat, const uint64_t *ecs_offsets, uint32_t n_ecs,
                           uint64_t interval, int extend,
                           const jpeg_sm100_huff_table dc[4], const jpeg_sm100_huff_table ac[4],
                           jpeg_sm100_plane_i16 *planes, uint32_t n_planes);

/* replaces  Spectral.Plane.idct(quanta:precision:) -> Planar.Plane   decode.swift:4101-4133
 * (modulate 3984, load 4020, idct8 4042, idct8x8 4095).  samples: 64 * units_x * units_y uint16. */
int jpeg_sm100_idct(jpeg_sm100_ctx *ctx, const int16_t *coef, uint32_t units_x, uint32_t units_y,
                    const uint16_t quanta_zigzag[64], int precision, uint16_t *samples);
/* same, 8-bit samples (precision must be 8): the 192 B/block variant the roofline is quoted on */
int jpeg_sm100_idct_u8(jpeg_sm100_ctx *ctx, const int16_t *coef, uint32_t units_x, uint32_t units_y,
                       const uint16_t quanta_zigzag[64], uint8_t *samples);

/* replaces  Planar.interleaved(cosite:) -> Rectangular   decode.swift:4182-4276 */
int jpeg_sm100_interleave(jpeg_sm100_ctx *ctx, const jpeg_sm100_plane_u16 *planes, uint32_t n_planes,
                          uint32_t size_x, uint32_t size_y, int cosited, uint16_t *interleaved);

/* replaces  JPEG.RGB.unpack(_:of:) / JPEG.YCbCr.unpack(_:of:)   jpeg.swift:551-572, 493-513 (YCbCr.rgb 441-453)
 * arity = 1 (y8 / nonconforming1x8) or 3 (ycc8 / nonconforming3x8) */
int jpeg_sm100_unpack_rgb8(jpeg_sm100_ctx *ctx, const uint16_t *interleaved, uint64_t n_pixels, int arity,
                           uint8_t *rgb);
int jpeg_sm100_unpack_ycc8(jpeg_sm100_ctx *ctx, const uint16_t *interleaved, uint64_t n_pixels, int arity,
                           uint8_t *ycc);

/* fused fast path for  Spectral.idct().interleaved(cosite:).unpack(as: RGB.self)  (decode.swift:4154, 4182, 4294):
 * coefficients -> RGB8 without materialising Planar / Rectangular on the host.
 * quanta: n_planes x 64 (zig-zag), factors: n_planes x (x, y). */
int jpeg_sm100_spectral_to_rgb8(jpeg_sm100_ctx *ctx, const jpeg_sm100_plane_i16 *planes, uint32_t n_planes,
                                const uint16_t *quanta_zigzag, const int32_t *factors_xy,
                                uint32_t size_x, uint32_t size_y, int cosited, uint8_t *rgb);

/* batch form of the above for N baseline (single sequential scan) images of identical geometry:
 *   for each image:  Spectral.decode(ecss:...) -> idct() -> interleaved(cosite:) -> unpack(as: RGB.self)
 * i.e. Rectangular.decompress(stream:) + unpack (decode.swift:4367, 4294) minus the container lexer.
 * Host buffers in, host RGB8 out, one call, copies included; pass pinned memory (jpeg_sm100_malloc_host) for full
 * PCIe rate.  Every plane of the image must be a component of `scan`.
 *   ecs_offsets : n_images * n_ecs + 1 offsets into ecs_concat (image-major, unstuffed bytes)
 *   tables      : n_images x 8 (dc[4], ac[4]) or 1 x 8 if tables_shared
 *   quanta      : n_planes x 64 (zig-zag), shared by all images
 *   rgb         : n_images * size_y * size_x * 3 bytes;  status: n_images codes (0 or the image's first error) */
int jpeg_sm100_decode_batch_rgb8(jpeg_sm100_ctx *ctx, const jpeg_sm100_scan_desc *scan, uint32_t n_images,
                                 const uint8_t *ecs_concat, const uint64_t *ecs_offsets, uint32_t n_ecs,
                                 uint64_t interval, const jpeg_sm100_huff_table *tables, int tables_shared,
                                 const uint16_t *quanta_zigzag, uint32_t size_x, uint32_t size_y, int cosited,
                                 uint8_t *rgb, int32_t *status);

/* ---- N1: the per-scan half of the lexer on the GPU ---- */

/* replaces the  `for index in 0... { (ecs, marker) = try stream.segment(prefix: true) ... }`  loop of
 * Context.decompress (decode.swift:3895-3933; Bytestream.segment decode.swift:130-190) for one scan:
 * byte unstuffing (FF 00 -> FF), fill bytes, RSTn splitting, restart-phase validation.
 *   raw / raw_len : the bytes that follow the SOS header up to (not including) the FF of the next non-RSTn marker
 *   ecs           : out, unstuffed bytes of all entropy-coded segments back to back (capacity >= raw_len)
 *   ecs_offsets   : out, *n_ecs + 1 offsets (capacity offsets_capacity entries; ERR_NO_MEMORY if too small)
 * Errors: ERR_RESTART_PHASE; ERR_INVALID_ARGUMENT if raw contains a marker other than RSTn. */
int jpeg_sm100_lex_scan(jpeg_sm100_ctx *ctx, const uint8_t *raw, uint64_t raw_len, uint8_t *ecs, uint64_t ecs_capacity,
                        uint64_t *ecs_offsets, uint32_t offsets_capacity, uint32_t *n_ecs);

/* lexer + Spectral.decode(ecss:...) in one call, the scan bytes never return to the host:
 * same arguments as jpeg_sm100_decode_scan with the raw (stuffed, RSTn-delimited) scan bytes instead of ecss.
 * interval = JPEG_SM100_INTERVAL_NONE with more than one segment -> ERR_MISSING_INTERVAL (decode.swift:3708-3720). */
int jpeg_sm100_decode_scan_raw(jpeg_sm100_ctx *ctx, const jpeg_sm100_scan_desc *scan, const uint8_t *raw, uint64_t raw_len,
                               uint64_t interval, int extend,
                               const jpeg_sm100_huff_table dc[4], const jpeg_sm100_huff_table ac[4],
                               jpeg_sm100_plane_i16 *planes, uint32_t n_planes);

/* jpeg_sm100_decode_batch_rgb8 fed with raw scan bytes: image i's scan data is raw_concat[raw_offsets[i] ..+ raw_lengths[i]).
 * Every image must contain exactly n_ecs segments (else its status is ERR_ECS_COUNT). */
int jpeg_sm100_decode_batch_raw_rgb8(jpeg_sm100_ctx *ctx, const jpeg_sm100_scan_desc *scan, uint32_t n_images,
                                     const uint8_t *raw_concat, const uint64_t *raw_offsets, const uint64_t *raw_lengths,
                                     uint32_t n_ecs, uint64_t interval, const jpeg_sm100_huff_table *tables, int tables_shared,
                                     const uint16_t *quanta_zigzag, uint32_t size_x, uint32_t size_y, int cosited,
                                     uint8_t *rgb, int32_t *status);

/* ---- encode ---- */

/* replaces  JPEG.RGB.pack(_:as:)   jpeg.swift:584-599 (RGB.ycc 463-478) */
int jpeg_sm100_pack_rgb8(jpeg_sm100_ctx *ctx, const uint8_t *rgb, uint64_t n_pixels, int arity,
                         uint16_t *interleaved);
/* replaces  Rectangular.decomposed() -> Planar   encode.swift:389-425 ; planes[].samples are outputs */
int jpeg_sm100_decompose(jpeg_sm100_ctx *ctx, const uint16_t *interleaved, uint32_t size_x, uint32_t size_y,
                         jpeg_sm100_plane_u16 *planes, uint32_t n_planes);
/* replaces  Spectral.Plane.fdct(_:quanta:precision:)   encode.swift:199-248 (load 80, fdct8 123, fdct8x8 191) */
int jpeg_sm100_fdct(jpeg_sm100_ctx *ctx, const uint16_t *samples, uint32_t units_x, uint32_t units_y,
                    const uint16_t quanta_zigzag[64], int precision, int16_t *coef);
/* replaces  Spectral.encode(scan:) -> (dc, ac, header, ecs)   encode.swift:1559-1620 (block symbols 919-959,
 * scan encoders 962-1557, optimal tables 602-772 + common.swift:127-296, bit packing jpeg.swift:1920-2007).
 *   interval_mcus = 0 reproduces the reference (single ECS).  >0 (a multiple of the row width) additionally
 *   emits RSTn markers -- an extension the reference encoder does not have.
 *   ecs: caller buffer of ecs_capacity bytes; *ecs_len receives the stuffed length (exactly what
 *   stream.format(prefix:) writes, encode.swift:1967).  If it does not fit: ERR_NO_MEMORY with *ecs_len = needed. */
int jpeg_sm100_encode_scan(jpeg_sm100_ctx *ctx, const jpeg_sm100_scan_desc *scan,
                           const jpeg_sm100_plane_i16 *planes, uint32_t n_planes, uint64_t interval_mcus,
                           jpeg_sm100_huff_table dc_out[4], jpeg_sm100_huff_table ac_out[4],
                           uint8_t *ecs, uint64_t ecs_capacity, uint64_t *ecs_len);
/* fused fast path for  pack -> decomposed() -> fdct(quanta:)  : RGB8 -> coefficient planes */
int jpeg_sm100_rgb8_to_spectral(jpeg_sm100_ctx *ctx, const uint8_t *rgb, uint32_t size_x, uint32_t size_y,
                                jpeg_sm100_plane_i16 *planes, uint32_t n_planes,
                                const uint16_t *quanta_zigzag, const int32_t *factors_xy);

/* ---- N3: spectral-domain operations (no reference library function; the loops of its examples) ---- */

/* the requantisation loop of examples/recompress/main.swift:35-58, one plane:
 *   out[z] = Int16(Double(Int16(q_old[z]) * coef[z]) / Double(q_new[z]) + 0.3 * sign)      (truncating conversion) */
int jpeg_sm100_requantize(jpeg_sm100_ctx *ctx, const int16_t *coef, uint32_t units_x, uint32_t units_y,
                          const uint16_t q_old[64], const uint16_t q_new[64], int16_t *out);
/* the block loop of examples/rotate/main.swift:164-190, one plane: source block s lands at offset + M s, where
 * M = ((matrix[0], matrix[1]), (matrix[2], matrix[3])) and mirrored axes start at the far end of the source plane
 * (main.swift:168-172); destination coefficient z = source coefficient zmap[z] * mul[z] (Block.transform, main.swift:13-99).
 * Destination blocks outside out_units are dropped, blocks nothing lands on are zero. */
int jpeg_sm100_transform_blocks(jpeg_sm100_ctx *ctx, const int16_t *coef, uint32_t units_x, uint32_t units_y,
                                const int32_t matrix[4], const uint8_t zmap[64], const int8_t mul[64],
                                int16_t *out, uint32_t out_units_x, uint32_t out_units_y);

/* ---- layer A with the image RESIDENT on the device ---------------------------------------------------------------------
 * JPEG.Context holds ONE Spectral for the whole file and pushes every scan into it (decode.swift:3565-3587, 3706-3725); the
 * stage functions then read it (Spectral.idct() decode.swift:4154 ...).  With the calls above every scan uploads and downloads
 * all planes.  A jpeg_sm100_spectral is the same image kept in HBM: the Swift side stores the handle in its Spectral (a final
 * class member, swift/JPEGSM100Shim.swift), pushes scans into it, runs the stages on it, and materialises host planes once,
 * on demand.  Host buffers in, host buffers out, synchronous, like the rest of layer A; a handle belongs to the ctx that made it.
 *   units_xy : n_planes x (units_x, units_y), the Spectral.Plane geometry (decode.swift:2456-2495).  New planes are zero
 *              (decode.swift:2241-2256). */
typedef struct jpeg_sm100_spectral jpeg_sm100_spectral;
int  jpeg_sm100_spectral_create(jpeg_sm100_ctx *ctx, uint32_t n_planes, const int32_t *units_xy, jpeg_sm100_spectral **out);
void jpeg_sm100_spectral_destroy(jpeg_sm100_ctx *ctx, jpeg_sm100_spectral *s);
/* Spectral.set(width:) / set(height:) (decode.swift:2456-2495): new geometry, the common region keeps its coefficients, the
 * rest is zero (DNL height redefinition decode.swift:3905-3924, cropping before a rotation examples/rotate/main.swift:101-199) */
int  jpeg_sm100_spectral_resize(jpeg_sm100_ctx *ctx, jpeg_sm100_spectral *s, const int32_t *units_xy);
/* host planes -> device (e.g. coefficients produced by the host) and device -> host (materialise Spectral.Plane buffers) */
int  jpeg_sm100_spectral_upload(jpeg_sm100_ctx *ctx, jpeg_sm100_spectral *s, const jpeg_sm100_plane_i16 *planes, uint32_t n_planes);
int  jpeg_sm100_spectral_download(jpeg_sm100_ctx *ctx, const jpeg_sm100_spectral *s, jpeg_sm100_plane_i16 *planes, uint32_t n_planes);
/* jpeg_sm100_decode_scan / jpeg_sm100_decode_scan_raw into the resident image: only the scan's bytes travel */
int  jpeg_sm100_spectral_decode_scan(jpeg_sm100_ctx *ctx, jpeg_sm100_spectral *s, const jpeg_sm100_scan_desc *scan,
                                     const uint8_t *ecs_concat, const uint64_t *ecs_offsets, uint32_t n_ecs, uint64_t interval,
                                     int extend, const jpeg_sm100_huff_table dc[4], const jpeg_sm100_huff_table ac[4]);
int  jpeg_sm100_spectral_decode_scan_raw(jpeg_sm100_ctx *ctx, jpeg_sm100_spectral *s, const jpeg_sm100_scan_desc *scan,
                                         const uint8_t *raw, uint64_t raw_len, uint64_t interval, int extend,
                                         const jpeg_sm100_huff_table dc[4], const jpeg_sm100_huff_table ac[4]);
/* jpeg_sm100_encode_scan from the resident image */
int  jpeg_sm100_spectral_encode_scan(jpeg_sm100_ctx *ctx, const jpeg_sm100_spectral *s, const jpeg_sm100_scan_desc *scan,
                                     uint64_t interval_mcus, jpeg_sm100_huff_table dc_out[4], jpeg_sm100_huff_table ac_out[4],
                                     uint8_t *ecs, uint64_t ecs_capacity, uint64_t *ecs_len);
/* Spectral.idct() (decode.swift:4154): every plane, host sample planes out (planes[p].samples: 64 * units_x * units_y uint16;
 * a NULL samples pointer skips that plane's download).  quanta: n_planes x 64 (zig-zag). */
int  jpeg_sm100_spectral_idct(jpeg_sm100_ctx *ctx, const jpeg_sm100_spectral *s, const uint16_t *quanta_zigzag, int precision,
                              jpeg_sm100_plane_u16 *planes, uint32_t n_planes);
/* jpeg_sm100_spectral_to_rgb8 from the resident image: idct().interleaved(cosite:).unpack(as: RGB.self) with one download */
int  jpeg_sm100_spectral_rgb8(jpeg_sm100_ctx *ctx, const jpeg_sm100_spectral *s, const uint16_t *quanta_zigzag,
                              const int32_t *factors_xy, uint32_t size_x, uint32_t size_y, int cosited, uint8_t *rgb);
/* bytes this ctx has moved host -> device and device -> host so far (layer A bookkeeping for benchmarks and tests) */
void jpeg_sm100_transfer_counts(jpeg_sm100_ctx *ctx, uint64_t *h2d_bytes, uint64_t *d2h_bytes);

/* =============================================================================================================
 * LAYER B -- device pointers, asynchronous on the ctx stream, batched over images of identical geometry
 * ============================================================================================================= */

/* coefficient planes of a batch: image i of plane p starts at coef + i * image_stride (int16 elements) */
typedef struct {
    uint32_t n_images, n_planes;
    struct { int16_t *coef; uint64_t image_stride; int32_t units_x, units_y, factor_x, factor_y; } plane[4];
} jpeg_sm100_dev_spectral;

/* sample planes of a batch (8-bit or 16-bit samples) */
typedef struct {
    uint32_t n_images, n_planes;
    int32_t  sample_bytes; /* 1 or 2 */
    struct { void *samples; uint64_t image_stride; /* in samples */ int32_t units_x, units_y, factor_x, factor_y; } plane[4];
} jpeg_sm100_dev_planar;

/* decode one scan of every image in the batch.
 *   d_ecs          : device, all unstuffed ECS bytes of all images (+ >= 8 bytes of slack after the last byte)
 *   d_ecs_offsets  : device, n_images * n_ecs + 1 byte offsets into d_ecs (image-major)
 *   tables         : HOST, n_images sets of (dc[4], ac[4]) -- tables[i*8 + 0..3] = dc, [i*8 + 4..7] = ac;
 *                    if tables_shared != 0 only set 0 is read and used for every image
 *   d_status       : device, n_images int32: 0 or the error of the lowest-index failing interval (may be NULL)
 *   extend         : JPEG_SM100_SCAN_EXTEND and/or JPEG_SM100_SCAN_FRESH (see above), 0 for neither
 * A baseline scan (band 0..<64) defines every coefficient of every block it decodes: blocks leave the decoder as whole
 * 128-byte lines, so whatever the planes held there before is replaced (for in-plane blocks of intervals that decode without
 * error); progressive scans touch only their band. */
int jpeg_sm100_dev_decode_scan(jpeg_sm100_ctx *ctx, const jpeg_sm100_scan_desc *scan,
                               const uint8_t *d_ecs, const uint64_t *d_ecs_offsets, uint32_t n_ecs,
                               uint64_t interval, int extend,
                               const jpeg_sm100_huff_table *tables, int tables_shared,
                               const jpeg_sm100_dev_spectral *spectral, int32_t *d_status);

/* N1 on device memory: lex the scan bytes of n_images images into the inputs of jpeg_sm100_dev_decode_scan.
 *   d_raw          : device, 16-byte aligned, readable for 16 bytes past the last image
 *   raw_offsets / raw_lengths : HOST, n_images each (any alignment)
 *   d_ecs          : device, 16-byte aligned, out: unstuffed bytes back to back; capacity >= sum(raw_lengths) + 16
 *   d_ecs_offsets  : device, out: n_images * n_ecs + 1
 *   d_status       : device, n_images int32: 0, ERR_RESTART_PHASE, ERR_ECS_COUNT or ERR_INVALID_ARGUMENT (foreign marker) */
int jpeg_sm100_dev_lex_scan(jpeg_sm100_ctx *ctx, const uint8_t *d_raw, const uint64_t *raw_offsets, const uint64_t *raw_lengths,
                            uint32_t n_images, uint32_t n_ecs, uint8_t *d_ecs, uint64_t *d_ecs_offsets, int32_t *d_status);

/* dequantise + IDCT every plane of every image.  quanta: HOST, n_planes x 64 (shared by all images). */
int jpeg_sm100_dev_idct(jpeg_sm100_ctx *ctx, const jpeg_sm100_dev_spectral *spectral,
                        const uint16_t *quanta_zigzag, int precision, const jpeg_sm100_dev_planar *planar);

/* upsample + YCbCr->RGB + pack; d_rgb: n_images * size_x * size_y * 3 bytes */
int jpeg_sm100_dev_planar_to_rgb8(jpeg_sm100_ctx *ctx, const jpeg_sm100_dev_planar *planar,
                                  uint32_t size_x, uint32_t size_y, int cosited, uint8_t *d_rgb);
/* coefficients -> RGB8 in ONE kernel where the geometry allows (three planes of 8-bit samples, 4:2:0 with centred chroma: the chain
 * Spectral.idct() -> Planar.interleaved(cosite: false) -> unpack(as: RGB.self), decode.swift:4154 -> 4182 -> jpeg.swift:441, without
 * the sample planes in between: 6 instead of 9 bytes of HBM traffic per pixel); any other geometry runs jpeg_sm100_dev_idct into
 * context-owned scratch planes followed by jpeg_sm100_dev_planar_to_rgb8.  Bit-identical to the staged calls either way.
 * The one-kernel path is opt-in (environment JPEG_SM100_FUSE=1, read per call): measured slower than the staged pair on B200.
 * The planes' factor_x / factor_y must be set.  quanta: HOST, n_planes x 64, zig-zag order. */
int jpeg_sm100_dev_spectral_to_rgb8(jpeg_sm100_ctx *ctx, const jpeg_sm100_dev_spectral *spectral, const uint16_t *quanta_zigzag,
                                    uint32_t size_x, uint32_t size_y, int cosited, uint8_t *d_rgb);
/* upsample + interleave (16-bit Rectangular.values), and the unpack kernels on device memory */
int jpeg_sm100_dev_interleave(jpeg_sm100_ctx *ctx, const jpeg_sm100_dev_planar *planar,
                              uint32_t size_x, uint32_t size_y, int cosited, uint16_t *d_interleaved);
int jpeg_sm100_dev_unpack_rgb8(jpeg_sm100_ctx *ctx, const uint16_t *d_interleaved, uint64_t n_pixels, int arity,
                               uint8_t *d_rgb);
int jpeg_sm100_dev_unpack_ycc8(jpeg_sm100_ctx *ctx, const uint16_t *d_interleaved, uint64_t n_pixels, int arity,
                               uint8_t *d_ycc);

/* encode mirror */
int jpeg_sm100_dev_rgb8_to_planar(jpeg_sm100_ctx *ctx, const uint8_t *d_rgb, uint32_t size_x, uint32_t size_y,
                                  const jpeg_sm100_dev_planar *planar);
int jpeg_sm100_dev_pack_rgb8(jpeg_sm100_ctx *ctx, const uint8_t *d_rgb, uint64_t n_pixels, int arity,
                             uint16_t *d_interleaved);
int jpeg_sm100_dev_decompose(jpeg_sm100_ctx *ctx, const uint16_t *d_interleaved, uint32_t size_x, uint32_t size_y,
                             const jpeg_sm100_dev_planar *planar);
int jpeg_sm100_dev_fdct(jpeg_sm100_ctx *ctx, const jpeg_sm100_dev_planar *planar,
                        const uint16_t *quanta_zigzag, int precision, const jpeg_sm100_dev_spectral *spectral);
/* entropy-encode one scan of every image.  Output per image i: stuffed bytes at d_ecs + i * ecs_image_stride,
 * length in d_ecs_len[i]; tables_out (HOST, n_images x 8, layout as in dev_decode_scan) filled after an internal
 * stream sync (optimal-table construction runs on the host between the statistics and the emit kernels). */
int jpeg_sm100_dev_encode_scan(jpeg_sm100_ctx *ctx, const jpeg_sm100_scan_desc *scan,
                               const jpeg_sm100_dev_spectral *spectral, uint64_t interval_mcus,
                               jpeg_sm100_huff_table *tables_out,
                               uint8_t *d_ecs, uint64_t ecs_image_stride, uint64_t *d_ecs_len);

/* N3 on device batches.  q_old / q_new: HOST, n_planes x 64.  in and out must have the same plane count and image count;
 * for requantize also the same units (in-place, out == in, is allowed); for transform_blocks `out` has its own geometry. */
int jpeg_sm100_dev_requantize(jpeg_sm100_ctx *ctx, const jpeg_sm100_dev_spectral *in, const uint16_t *q_old,
                              const uint16_t *q_new, const jpeg_sm100_dev_spectral *out);
int jpeg_sm100_dev_transform_blocks(jpeg_sm100_ctx *ctx, const jpeg_sm100_dev_spectral *in, const int32_t matrix[4],
                                    const uint8_t zmap[64], const int8_t mul[64], const jpeg_sm100_dev_spectral *out);

#ifdef __cplusplus
}
#endif
#endif /* JPEG_SM100_H */

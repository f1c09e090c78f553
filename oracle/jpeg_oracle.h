/*
 * jpeg_oracle.h -- CPU ORACLE for the JPEG block-transform hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load it.  The product path (libjpeg_sm100.so + jpeg_b200/host) never
 * links, imports or calls anything in oracle/.
 *
 * It is an op-for-op C restatement of the pure-Swift reference
 * (tayloraswift/jpeg @ 8fe8fda1, sources/jpeg/{decode,encode,jpeg,common}.swift);
 * every function cites the reference file:line it follows.  Parity is PINNED:
 * tests/test_oracle_golden.py checks it bit-for-bit against the reference's own
 * committed outputs (tests/regression/gold/ ycc + rgb, examples/decode-basic,
 * examples/decode-advanced plane dumps, 32 examples/encode-basic JPEG files,
 * tests/unit KATs).
 *
 * Build: gcc -O2 -ffp-contract=off (no -ffast-math): all float arithmetic is
 * IEEE binary32, evaluated left-to-right, never fused.
 */
#ifndef JPEG_ORACLE_H
#define JPEG_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* error codes (negative); mirror the JPEG.*Error cases the hot path raises */
enum {
    ORC_OK                          =   0,
    ORC_ERR_TRUNCATED_ECS           =  -1, /* DecodingError.truncatedEntropyCodedSegment  decode.swift:2778 */
    ORC_ERR_INVALID_COMPOSITE_VALUE =  -2, /* DecodingError.invalidCompositeValue         decode.swift:3109 */
    ORC_ERR_INVALID_BLOCK_RUN       =  -3, /* DecodingError.invalidCompositeBlockRun      decode.swift:2951 */
    ORC_ERR_UNDEFINED_DC            =  -4, /* DecodingError.undefinedScanHuffmanDCReference decode.swift:2888 */
    ORC_ERR_UNDEFINED_AC            =  -5, /* DecodingError.undefinedScanHuffmanACReference decode.swift:2894 */
    ORC_ERR_UNDEFINED_QUANTA        =  -6, /* DecodingError.undefinedScanQuantizationReference decode.swift:3467 */
    ORC_ERR_PRECONDITION            =  -7, /* the reference would trap (precondition / array bounds) */
    ORC_ERR_LEX                     = -10, /* JPEG.LexingError.*    */
    ORC_ERR_PARSE                   = -11, /* JPEG.ParsingError.*   */
    ORC_ERR_DECODE                  = -12, /* other JPEG.DecodingError.* (structure / progression) */
    ORC_ERR_UNSUPPORTED             = -13
};

#define ORC_INTERVAL_NONE INT64_MAX

/* ---- tables ------------------------------------------------------------- */

typedef struct {
    int     present;
    uint8_t counts[16];
    uint8_t values[256];
} orc_huff_spec;

/* zig-zag index of (k = horizontal frequency, h = vertical frequency); decode.swift:1289-1298 */
int  orc_zigzag(int k, int h);
/* T.81 EXTEND and its inverse; decode.swift:2742-2771 */
int  orc_extend(int binade, unsigned tail);
void orc_compact(int x, int *binade, unsigned *tail);
/* Huffman 2-level LUT (decode.swift:310-351, 1037-1265): returns 0 or ORC_ERR_PARSE; lookup gives (symbol,length) */
int  orc_huff_lookup(const orc_huff_spec *spec, unsigned codeword16, int *symbol, int *length);
/* Optimal table from frequencies (encode.swift:602-772, common.swift:127-296) */
void orc_huff_from_frequencies(const int64_t freq[256], orc_huff_spec *out);
/* canonical codes (encode.swift:664-680, 799-820): code[s], len[s] (len 0 = absent) */
void orc_huff_encoder(const orc_huff_spec *spec, uint16_t code[256], uint8_t len[256]);
/* CompressionLevel.quanta (encode.swift:286-333), zig-zag order */
void orc_quanta(double level, int chrominance, uint16_t out[64]);

/* ---- spectral image ------------------------------------------------------ */

typedef struct orc_spectral orc_spectral;

/* full container decode: lexer + parsers + JPEG.Context.decompress (decode.swift:130-190, 475-1005, 3728-3960) */
orc_spectral *orc_decompress(const uint8_t *jpeg, size_t n, int *err);
/* the same for a user-defined JPEG.Format in the style of examples/custom-color/main.swift:41-63 (recognised iff the frame's
 * component keys are exactly format_components and its precision is format_precision; planes in the listed order) */
orc_spectral *orc_decompress_format(const uint8_t *jpeg, size_t n, const int *format_components, int n_components,
                                    int format_precision, int *err);
/* JPEG.Context driven segment by segment, stopped after `scans` scans (examples/decode-online/main.swift) */
orc_spectral *orc_decompress_scans(const uint8_t *jpeg, size_t n, int scans, int *err);
int  orc_spectral_precision(const orc_spectral *s);
void orc_spectral_set_format(orc_spectral *s, const int *component_ids, int precision);
/* build an empty spectral image (Spectral.init(size:layout:...), decode.swift:2413) */
orc_spectral *orc_spectral_create(int size_x, int size_y, int ncomp, const int *factors_xy /*2*ncomp*/,
                                  int progressive);
void orc_spectral_free(orc_spectral *s);
void orc_spectral_info(const orc_spectral *s, int *size_xy, int *blocks_xy, int *ncomp, int *scale_xy, int *process);
void orc_spectral_plane_info(const orc_spectral *s, int p, int *units_xy, int *factor_xy, int *comp_id);
int16_t  *orc_spectral_coefficients(orc_spectral *s, int p);   /* 64*ux*uy int16, index 64*(ux*y+x)+z */
uint16_t *orc_spectral_quanta(orc_spectral *s, int p);         /* 64 u16, zig-zag order */
void orc_spectral_set_quanta(orc_spectral *s, int p, const uint16_t q[64]);

/* one scan, already lexed: Spectral.decode(ecss:interval:scan:tables:extend:) decode.swift:3476-3551.
 * comps[i] = plane index; dcsel/acsel = table slots. ecs_concat/offsets: n_ecs+1 offsets (unstuffed bytes). */
int orc_decode_scan(orc_spectral *s,
                    int band_lo, int band_hi, int bit_lo, int bit_hi /* <0 = .max */,
                    int ncomp, const int *comps, const int *dcsel, const int *acsel,
                    const orc_huff_spec dc[4], const orc_huff_spec ac[4],
                    const uint8_t *ecs_concat, const uint64_t *ecs_offsets, int n_ecs,
                    int64_t interval, int extend);

/* ---- transform stages ------------------------------------------------------ */

/* Spectral.Plane.idct (decode.swift:3984-4133): out = 8ux x 8uy uint16, row-major */
void orc_idct_plane(const int16_t *coef, int units_x, int units_y, const uint16_t quanta_zz[64],
                    int precision, uint16_t *out);
/* Planar.interleaved(cosite:) (decode.swift:4182-4276) */
void orc_interleave(const uint16_t *const *planes, const int *units_xy, const int *factors_xy, int ncomp,
                    int size_x, int size_y, int cosited, uint16_t *out);
/* RGB.unpack / YCbCr.unpack (jpeg.swift:441-453, 493-572) */
void orc_unpack_rgb(const uint16_t *interleaved, size_t npx, int ncomp, uint8_t *rgb);
void orc_unpack_ycc(const uint16_t *interleaved, size_t npx, int ncomp, uint8_t *ycc);
/* RGB.pack (jpeg.swift:463-478, 584-599) */
void orc_pack_rgb(const uint8_t *rgb, size_t npx, int ncomp, uint16_t *interleaved);
/* Rectangular.decomposed() (encode.swift:389-425) for one plane: out = 8ux x 8uy */
void orc_decompose_plane(const uint16_t *interleaved, int size_x, int size_y, int ncomp, int p,
                         int fx, int fy, int scale_x, int scale_y, uint16_t *out);
/* Spectral.Plane.fdct (encode.swift:80-248) */
void orc_fdct_plane(const uint16_t *samples, int units_x, int units_y, const uint16_t quanta_zz[64],
                    int precision, int16_t *coef);

/* convenience: whole pipelines */
int  orc_spectral_to_planes(orc_spectral *s, uint16_t **planes /* ncomp malloc'd */);
int  orc_decode_rgb(const uint8_t *jpeg, size_t n, int *size_xy, uint8_t **rgb /* malloc'd */, uint8_t **ycc);
void orc_free(void *p);

/* ---- entropy encode ------------------------------------------------------ */

/* Spectral.encode(scan:) (encode.swift:919-1620).  interval_mcus = 0 reproduces the reference (one ECS);
 * >0 is OUR extension (restart markers every interval_mcus MCUs, must be a multiple of the row width).
 * Returns malloc'd stuffed bytes incl. RSTn markers between intervals; tables written to dc_out/ac_out by slot. */
int orc_encode_scan(orc_spectral *s,
                    int band_lo, int band_hi, int bit_lo, int bit_hi,
                    int ncomp, const int *comps, const int *dcsel, const int *acsel,
                    int64_t interval_mcus,
                    orc_huff_spec dc_out[4], orc_huff_spec ac_out[4],
                    uint8_t **ecs, size_t *ecs_len);

#ifdef __cplusplus
}
#endif
#endif

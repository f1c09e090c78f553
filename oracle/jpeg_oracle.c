/*
 * jpeg_oracle.c -- CPU ORACLE (test infrastructure; see jpeg_oracle.h).
 *
 * Op-for-op restatement of tayloraswift/jpeg @ 8fe8fda1.  Citations are to
 * /root/reference/sources/jpeg/<file>.swift:<line>.
 *
 * Build with: gcc -O2 -ffp-contract=off -fno-fast-math
 */
#include "jpeg_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define API __attribute__((visibility("default")))

/* ========================================================================= */
/* zig-zag, amplitude coding                                                  */
/* ========================================================================= */

/* decode.swift:1289-1298  JPEG.Table.Quantization.z(k:h:) */
API int orc_zigzag(int x, int y)
{
    int p = (x + y < 8) ? 1 : 0, q = (x + y) & 1;
    int a = 72 * (p ^ 1), b = 2 * p - 1;
    int n = b * (x + y) - 14 * p + 15;
    int t = (n * (n + 1)) >> 1;
    return a + b * t - q * x - (q ^ 1) * y - 1;
}

static int ZZ[8][8]; /* ZZ[h][k] */
static int ZZ_ready = 0;
static void zz_init(void)
{
    if (ZZ_ready) return;
    for (int h = 0; h < 8; ++h)
        for (int k = 0; k < 8; ++k) ZZ[h][k] = orc_zigzag(k, h);
    ZZ_ready = 1;
}

/* decode.swift:2742-2754  Bitstream.extend(binade:_:as:); masking shifts (&>>, &<<) on UInt16 */
API int orc_extend(int binade, unsigned tail)
{
    uint16_t t    = (uint16_t) tail;
    uint16_t sign = (uint16_t) (t >> ((binade - 1) & 15));
    uint16_t high = (uint16_t) ((uint16_t) (0xffffu + sign) << (binade & 15));
    uint16_t low  = (uint16_t) (t + (sign ^ 1u));
    return (int) (int16_t) (high | low);
}

/* decode.swift:2757-2771  Bitstream.compact(_:) */
API void orc_compact(int xi, int *binade, unsigned *tail)
{
    int16_t  x   = (int16_t) xi;
    int      mag = x < 0 ? -(int) x : (int) x;
    int      b   = 0;
    while (b < 16 && (mag >> b)) ++b; /* 16 - leadingZeroBitCount(abs(x)) */
    uint16_t sign = (uint16_t) ((uint16_t) x >> 15);
    uint16_t mask = (uint16_t) ~((uint16_t) (0xffffu << (b & 15)));
    if (b == 16) mask = 0xffffu; /* unreachable for |x| <= 32767 */
    *binade = b;
    *tail   = (uint16_t) ((uint16_t) ((uint16_t) x - sign) & mask);
}

/* ========================================================================= */
/* Huffman decoder: two-level LUT                                             */
/* ========================================================================= */

typedef struct {
    int      ok;
    int      n, z, zeta;
    uint8_t *sym, *len; /* z entries */
} huffdec;

/* decode.swift:310-351  Table.Huffman.size(_:) */
static int huff_size(const uint8_t counts[16], int *n_out, int *z_out)
{
    int interior = 1;
    for (int l = 0; l < 8; ++l) {
        if (!(interior > 0)) return 0;
        interior = 2 * interior - counts[l];
    }
    int n = 256 - interior, z = n;
    for (int i = 0; i < 8; ++i) {
        if (!(interior > 0)) return 0;
        z += (int) counts[8 + i] << (7 - i);
        interior = 2 * interior - counts[8 + i];
    }
    if (!(interior > 0)) return 0;
    *n_out = n;
    *z_out = z;
    return 1;
}

/* decode.swift:1037-1240  Table.Huffman.decoder() */
static int huffdec_build(const orc_huff_spec *spec, huffdec *d)
{
    memset(d, 0, sizeof *d);
    int n, z;
    if (!huff_size(spec->counts, &n, &z)) return ORC_ERR_PARSE;
    d->n = n;
    d->z = z;
    d->zeta = z + n * 255;
    d->sym = (uint8_t *) malloc((size_t) z + 1);
    d->len = (uint8_t *) malloc((size_t) z + 1);
    int count = 0, base = 0;
    for (int l = 0; l < 16; ++l) {
        if (!(count < z)) break;
        int clones = (0x8080 >> l) & 0xff;
        for (int s = 0; s < spec->counts[l]; ++s) {
            for (int c = 0; c < clones && count < z; ++c) {
                d->sym[count] = spec->values[base + s];
                d->len[count] = (uint8_t) (l + 1);
                ++count;
            }
        }
        base += spec->counts[l];
    }
    d->ok = 1;
    return ORC_OK;
}
static void huffdec_free(huffdec *d)
{
    free(d->sym);
    free(d->len);
    memset(d, 0, sizeof *d);
}
/* decode.swift:1246-1265  Decoder[codeword] */
static inline void huffdec_lookup(const huffdec *d, unsigned cw, int *symbol, int *length)
{
    int i = (int) (cw >> 8);
    if (i < d->n) {
        *symbol = d->sym[i];
        *length = d->len[i];
    } else {
        int j = (int) cw;
        if (!(j < d->zeta)) {
            *symbol = 0;
            *length = 16;
            return;
        }
        *symbol = d->sym[j - d->n * 255];
        *length = d->len[j - d->n * 255];
    }
}

API int orc_huff_lookup(const orc_huff_spec *spec, unsigned cw, int *symbol, int *length)
{
    huffdec d;
    int     e = huffdec_build(spec, &d);
    if (e) return e;
    huffdec_lookup(&d, cw & 0xffff, symbol, length);
    huffdec_free(&d);
    return ORC_OK;
}

/* ========================================================================= */
/* Bitstream reader                                                           */
/* ========================================================================= */

/* jpeg.swift:1873-1916  Bitstream.init / [i, count:] / [i, as:]; data is padded with 1-bits */
typedef struct {
    const uint8_t *d;
    int64_t        nbytes;
    int64_t        count; /* bits */
} bits_t;

static inline unsigned bits_peek16(const bits_t *b, int64_t i)
{
    int64_t  B = i >> 3;
    int      sh = (int) (i & 7);
    uint32_t w = 0;
    for (int k = 0; k < 3; ++k) {
        int64_t  at = B + k;
        unsigned by = (at < b->nbytes) ? b->d[at] : 0xffu;
        w = (w << 8) | by;
    }
    return (w >> (8 - sh)) & 0xffffu;
}
/* [i, count: c] = front &>> (16 - c)  -- masking shift */
static inline unsigned bits_peek(const bits_t *b, int64_t i, int c) { return bits_peek16(b, i) >> ((16 - c) & 15); }

/* decode.swift:2773-2786 */
static inline int bits_refinement(const bits_t *b, int64_t *i, int *bit)
{
    if (!(*i < b->count)) return ORC_ERR_TRUNCATED_ECS;
    *bit = (int) (bits_peek16(b, *i) >> 15);
    *i += 1;
    return ORC_OK;
}
/* decode.swift:2788-2820 */
static inline int bits_composite_dc(const bits_t *b, int64_t *i, const huffdec *t, int *difference)
{
    if (!(*i < b->count)) return ORC_ERR_TRUNCATED_ECS;
    int symbol, length;
    huffdec_lookup(t, bits_peek16(b, *i), &symbol, &length);
    int binade = symbol;
    *i += length;
    if (!(binade > 0)) {
        *difference = 0;
        return ORC_OK;
    }
    if (!(*i + binade <= b->count)) return ORC_ERR_TRUNCATED_ECS;
    *difference = orc_extend(binade, bits_peek(b, *i, binade));
    *i += binade;
    return ORC_OK;
}
/* decode.swift:2822-2872 ; returns kind 0 = run(run, value), 1 = eob(run) */
static inline int bits_composite_ac(const bits_t *b, int64_t *i, const huffdec *t, int *kind, int *run, int *value)
{
    if (!(*i < b->count)) return ORC_ERR_TRUNCATED_ECS;
    int symbol, length;
    huffdec_lookup(t, bits_peek16(b, *i), &symbol, &length);
    int zeroes = symbol >> 4, binade = symbol & 0x0f;
    *i += length;
    if (binade == 0) {
        if (zeroes == 0) {
            *kind = 1;
            *run = 1;
            return ORC_OK;
        }
        if (zeroes <= 14) {
            if (!(*i + zeroes <= b->count)) return ORC_ERR_TRUNCATED_ECS;
            *kind = 1;
            *run = (1 << zeroes) | (int) bits_peek(b, *i, zeroes);
            *i += zeroes;
            return ORC_OK;
        }
        *kind = 0;
        *run = 15;
        *value = 0;
        return ORC_OK;
    }
    if (!(*i + binade <= b->count)) return ORC_ERR_TRUNCATED_ECS;
    *kind = 0;
    *run = zeroes;
    *value = orc_extend(binade, bits_peek(b, *i, binade));
    *i += binade;
    return ORC_OK;
}

/* ========================================================================= */
/* Spectral image                                                             */
/* ========================================================================= */

typedef struct {
    int      ux, uy, fx, fy, q, comp_id;
    int16_t *coef; /* 64*ux*uy */
} splane;

struct orc_spectral {
    int    sx, sy, bx, by, ncomp, scx, scy, process;
    int    precision;       /* Format.precision (8 for JPEG.Common) */
    splane pl[4];
    int    nquanta;
    uint16_t (*quanta)[64]; /* quanta[0] = default (zeros), decode.swift:1723-1728 */
    int    approx[4][64];   /* Layout.Progression, jpeg.swift:1581-1634 */
};

/* decode.swift:1364-1369 */
static int units_of(int size, int stride) { return size / stride + (size % stride != 0 ? 1 : 0); }

/* decode.swift:2232-2290  Plane.set(width:) */
static void plane_set_width(splane *p, int x)
{
    if (x == p->ux) return;
    size_t   count = (size_t) 64 * x * p->uy;
    int16_t *nw = (int16_t *) calloc(count ? count : 1, sizeof(int16_t));
    if (p->coef) {
        size_t so = (size_t) 64 * p->ux, sn = (size_t) 64 * x;
        size_t cp = so < sn ? so : sn;
        for (int y = 0; y < p->uy; ++y) memcpy(nw + y * sn, p->coef + y * so, cp * sizeof(int16_t));
    }
    free(p->coef);
    p->coef = nw;
    p->ux = x;
}
/* decode.swift:2291-2312  Plane.set(height:) */
static void plane_set_height(splane *p, int y)
{
    if (y == p->uy) return;
    size_t   count = (size_t) 64 * p->ux * y, old = (size_t) 64 * p->ux * p->uy;
    int16_t *nw = (int16_t *) calloc(count ? count : 1, sizeof(int16_t));
    if (p->coef) memcpy(nw, p->coef, (count < old ? count : old) * sizeof(int16_t));
    free(p->coef);
    p->coef = nw;
    p->uy = y;
}
/* decode.swift:2456-2470 */
static void spectral_set_width(orc_spectral *s, int x)
{
    s->bx = units_of(x, 8 * s->scx);
    s->sx = x;
    for (int p = 0; p < s->ncomp; ++p) plane_set_width(&s->pl[p], units_of(x * s->pl[p].fx, 8 * s->scx));
}
/* decode.swift:2484-2495 */
static void spectral_set_height(orc_spectral *s, int y)
{
    s->by = units_of(y, 8 * s->scy);
    s->sy = y;
    for (int p = 0; p < s->ncomp; ++p) plane_set_height(&s->pl[p], units_of(y * s->pl[p].fy, 8 * s->scy));
}

static int spectral_push_quanta(orc_spectral *s, const uint16_t q[64])
{
    s->quanta = (uint16_t(*)[64]) realloc(s->quanta, sizeof(uint16_t[64]) * (size_t) (s->nquanta + 1));
    memcpy(s->quanta[s->nquanta], q, sizeof(uint16_t[64]));
    return s->nquanta++;
}

API orc_spectral *orc_spectral_create(int size_x, int size_y, int ncomp, const int *factors_xy, int progressive)
{
    zz_init();
    orc_spectral *s = (orc_spectral *) calloc(1, sizeof *s);
    s->ncomp = ncomp;
    s->precision = 8;
    s->process = progressive ? 2 : 0;
    for (int p = 0; p < ncomp; ++p) {
        s->pl[p].fx = factors_xy[2 * p];
        s->pl[p].fy = factors_xy[2 * p + 1];
        s->pl[p].comp_id = p + 1;
        if (s->pl[p].fx > s->scx) s->scx = s->pl[p].fx;
        if (s->pl[p].fy > s->scy) s->scy = s->pl[p].fy;
    }
    uint16_t zero[64] = {0};
    spectral_push_quanta(s, zero);
    for (int c = 0; c < 4; ++c)
        for (int z = 0; z < 64; ++z) s->approx[c][z] = INT32_MAX;
    spectral_set_width(s, size_x);
    spectral_set_height(s, size_y);
    return s;
}
API void orc_spectral_free(orc_spectral *s)
{
    if (!s) return;
    for (int p = 0; p < 4; ++p) free(s->pl[p].coef);
    free(s->quanta);
    free(s);
}
API void orc_spectral_info(const orc_spectral *s, int *size_xy, int *blocks_xy, int *ncomp, int *scale_xy, int *process)
{
    if (size_xy) size_xy[0] = s->sx, size_xy[1] = s->sy;
    if (blocks_xy) blocks_xy[0] = s->bx, blocks_xy[1] = s->by;
    if (ncomp) *ncomp = s->ncomp;
    if (scale_xy) scale_xy[0] = s->scx, scale_xy[1] = s->scy;
    if (process) *process = s->process;
}
API void orc_spectral_plane_info(const orc_spectral *s, int p, int *units_xy, int *factor_xy, int *comp_id)
{
    if (units_xy) units_xy[0] = s->pl[p].ux, units_xy[1] = s->pl[p].uy;
    if (factor_xy) factor_xy[0] = s->pl[p].fx, factor_xy[1] = s->pl[p].fy;
    if (comp_id) *comp_id = s->pl[p].comp_id;
}
API int16_t  *orc_spectral_coefficients(orc_spectral *s, int p) { return s->pl[p].coef; }
API uint16_t *orc_spectral_quanta(orc_spectral *s, int p) { return s->quanta[s->pl[p].q]; }
API void      orc_spectral_set_quanta(orc_spectral *s, int p, const uint16_t q[64])
{
    s->pl[p].q = spectral_push_quanta(s, q);
}

/* Plane[x:y:z:] get/set with the bounds guard, decode.swift:1455-1479 */
static inline int16_t pget(const splane *p, int x, int y, int z)
{
    if (!(x >= 0 && x < p->ux && y >= 0 && y < p->uy)) return 0;
    return p->coef[64 * ((size_t) p->ux * y + x) + z];
}
static inline void pset(splane *p, int x, int y, int z, int16_t v)
{
    if (!(x >= 0 && x < p->ux && y >= 0 && y < p->uy)) return;
    p->coef[64 * ((size_t) p->ux * y + x) + z] = v;
}

/* ========================================================================= */
/* Scan decoders                                                              */
/* ========================================================================= */

typedef struct {
    int64_t lo, hi; /* MCU/block range start ..< start + interval */
} blkrange;

static int64_t sat_add(int64_t a, int64_t b) { return (a > INT64_MAX - b) ? INT64_MAX : a + b; }

/* rows helper for the sequential / DC-first decoders:
 * (lo / w ..< hi / w).clamped(to: 0 ..< (extend ? .max : limit))   decode.swift:2897-2899, 3205-3207 */
static void rows_clamped(blkrange b, int w, int extend, int limit, int64_t *r0, int64_t *r1)
{
    int64_t a = b.lo / w, c = b.hi / w;
    int64_t lim = extend ? INT64_MAX : (int64_t) limit;
    if (a > lim) a = lim;
    if (c > lim) c = lim;
    if (a < 0) a = 0;
    if (c < a) c = a;
    *r0 = a;
    *r1 = c;
}
/* rows helper for the Range2-driven decoders (DC refine / AC scans):
 *   rows = lo / w ..< min(hi / w, limit);  for (x, y) in (0, rows.lower) ..< (w, rows.upper)
 * General.Range2 (common.swift:383-409) yields ONE row when the y range is empty and traps when
 * lower > upper; both are reproduced. */
static int rows_range2(blkrange b, int w, int limit, int64_t *r0, int64_t *nrows)
{
    int64_t a = b.lo / w, c = b.hi / w;
    if (c > limit) c = limit;
    if (a > c) return ORC_ERR_PRECONDITION;
    *r0 = a;
    *nrows = (c > a) ? c - a : 1;
    return ORC_OK;
}

/* decode.swift:2880-2956  Plane.decode (sequential, non-interleaved) */
static int plane_decode_sequential(splane *pl, const bits_t *bits, blkrange blocks, const huffdec *dc, const huffdec *ac,
                                   int extend)
{
    int64_t r0, r1;
    rows_clamped(blocks, pl->ux, extend, pl->uy, &r0, &r1);
    int64_t b = 0;
    int16_t pred = 0;
    for (int64_t y = r0; y < r1; ++y) {
        if (extend) {
            if (!(b < bits->count && bits_peek16(bits, b) != 0xffff)) break;
            if (y >= pl->uy) plane_set_height(pl, (int) y + 1);
        }
        for (int x = 0; x < pl->ux; ++x) {
            int diff, e;
            if ((e = bits_composite_dc(bits, &b, dc, &diff))) return e;
            pred = (int16_t) (pred + (int16_t) diff);
            pset(pl, x, (int) y, 0, pred);
            int z = 1;
            while (z < 64) {
                int kind, run, v;
                if ((e = bits_composite_ac(bits, &b, ac, &kind, &run, &v))) return e;
                if (kind == 0) {
                    z += run;
                    if (!(z < 64)) break;
                    pset(pl, x, (int) y, z, (int16_t) v);
                    z += 1;
                } else if (run == 1) {
                    break;
                } else {
                    return ORC_ERR_INVALID_BLOCK_RUN;
                }
            }
        }
    }
    return ORC_OK;
}

/* decode.swift:2960-3004  Plane.decode (DC first, non-interleaved) */
static int plane_decode_dc_first(splane *pl, const bits_t *bits, blkrange blocks, int a, const huffdec *dc, int extend)
{
    int64_t r0, r1;
    rows_clamped(blocks, pl->ux, extend, pl->uy, &r0, &r1);
    int64_t b = 0;
    int16_t pred = 0;
    for (int64_t y = r0; y < r1; ++y) {
        if (extend) {
            if (!(b < bits->count && bits_peek16(bits, b) != 0xffff)) break;
            if (y >= pl->uy) plane_set_height(pl, (int) y + 1);
        }
        for (int x = 0; x < pl->ux; ++x) {
            int diff, e;
            if ((e = bits_composite_dc(bits, &b, dc, &diff))) return e;
            pred = (int16_t) (pred + (int16_t) diff);
            pset(pl, x, (int) y, 0, (int16_t) ((uint16_t) pred << a));
        }
    }
    return ORC_OK;
}

/* decode.swift:3007-3018  Plane.decode (DC refine, non-interleaved) */
static int plane_decode_dc_refine(splane *pl, const bits_t *bits, blkrange blocks, int a)
{
    int64_t r0, nrows;
    int     e;
    if ((e = rows_range2(blocks, pl->ux, pl->uy, &r0, &nrows))) return e;
    int64_t b = 0;
    for (int64_t y = r0; y < r0 + nrows; ++y)
        for (int x = 0; x < pl->ux; ++x) {
            int bit;
            if ((e = bits_refinement(bits, &b, &bit))) return e;
            pset(pl, x, (int) y, 0, (int16_t) (pget(pl, x, (int) y, 0) | (int16_t) ((uint16_t) bit << a)));
        }
    return ORC_OK;
}

/* decode.swift:3021-3069  Plane.decode (AC first) */
static int plane_decode_ac_first(splane *pl, const bits_t *bits, blkrange blocks, int band_lo, int band_hi, int a,
                                 const huffdec *ac)
{
    int64_t r0, nrows;
    int     e;
    if ((e = rows_range2(blocks, pl->ux, pl->uy, &r0, &nrows))) return e;
    int64_t b = 0;
    int     skip = 0;
    for (int64_t y = r0; y < r0 + nrows; ++y)
        for (int x = 0; x < pl->ux; ++x) {
            int z = band_lo;
            while (z < band_hi) {
                if (skip != 0) {
                    skip -= 1;
                    break;
                }
                int kind, run, v;
                if ((e = bits_composite_ac(bits, &b, ac, &kind, &run, &v))) return e;
                if (kind == 0) {
                    z += run;
                    if (!(z < band_hi)) break;
                    pset(pl, x, (int) y, z, (int16_t) ((uint16_t) (int16_t) v << a));
                    z += 1;
                } else {
                    skip = run - 1;
                    break;
                }
            }
        }
    return ORC_OK;
}

/* decode.swift:3072-3152  Plane.decode (AC refine) */
static int plane_decode_ac_refine(splane *pl, const bits_t *bits, blkrange blocks, int band_lo, int band_hi, int a,
                                  const huffdec *ac)
{
    int64_t r0, nrows;
    int     e;
    if ((e = rows_range2(blocks, pl->ux, pl->uy, &r0, &nrows))) return e;
    int64_t b = 0;
    int     skip = 0;
    for (int64_t y = r0; y < r0 + nrows; ++y)
        for (int x = 0; x < pl->ux; ++x) {
            int z = band_lo;
            while (z < band_hi) { /* frequency: */
                int     zeroes;
                int16_t delta;
                if (skip > 0) {
                    zeroes = 64;
                    delta = 0;
                    skip -= 1;
                } else {
                    int kind, run, v;
                    if ((e = bits_composite_ac(bits, &b, ac, &kind, &run, &v))) return e;
                    if (kind == 0) {
                        if (!(v >= -1 && v <= 1)) return ORC_ERR_INVALID_COMPOSITE_VALUE;
                        zeroes = run;
                        delta = (int16_t) v;
                    } else {
                        zeroes = 64;
                        delta = 0;
                        skip = run - 1;
                    }
                }
                int skipped = 0, placed = 0;
                do {
                    int16_t unrefined = pget(pl, x, (int) y, z);
                    if (unrefined == 0) {
                        if (!(skipped < zeroes)) {
                            pset(pl, x, (int) y, z, (int16_t) ((uint16_t) delta << a));
                            z += 1;
                            placed = 1;
                            break; /* continue frequency */
                        }
                        skipped += 1;
                    } else {
                        int bit;
                        if ((e = bits_refinement(bits, &b, &bit))) return e;
                        int16_t d = (int16_t) ((unrefined < 0 ? -1 : 1) * bit);
                        pset(pl, x, (int) y, z, (int16_t) (unrefined + (int16_t) ((uint16_t) d << a)));
                    }
                    z += 1;
                } while (z < band_hi);
                if (!placed) break; /* break frequency */
            }
        }
    return ORC_OK;
}

typedef struct {
    int     p; /* plane index */
    int     fx, fy;
    huffdec dc, ac;
} descriptor;

/* decode.swift:3158-3291  Spectral.decode (sequential, interleaved) */
static int spectral_decode_sequential(orc_spectral *s, const bits_t *bits, blkrange blocks, int ncomp, descriptor *d,
                                      int extend)
{
    if (ncomp == 1) return plane_decode_sequential(&s->pl[d[0].p], bits, blocks, &d[0].dc, &d[0].ac, extend);
    int64_t r0, r1;
    rows_clamped(blocks, s->bx, extend, s->by, &r0, &r1);
    int64_t b = 0;
    int16_t pred[4] = {0, 0, 0, 0};
    for (int64_t my = r0; my < r1; ++my) {
        if (extend) {
            if (!(b < bits->count && bits_peek16(bits, b) != 0xffff)) break;
            for (int c = 0; c < ncomp; ++c) {
                int height = (int) (my + 1) * d[c].fy;
                if (height > s->pl[d[c].p].uy) plane_set_height(&s->pl[d[c].p], height);
            }
        }
        for (int mx = 0; mx < s->bx; ++mx)
            for (int c = 0; c < ncomp; ++c) {
                splane *pl = &s->pl[d[c].p];
                for (int y = (int) my * d[c].fy; y < (int) my * d[c].fy + d[c].fy; ++y)
                    for (int x = mx * d[c].fx; x < mx * d[c].fx + d[c].fx; ++x) {
                        int diff, e;
                        if ((e = bits_composite_dc(bits, &b, &d[c].dc, &diff))) return e;
                        pred[c] = (int16_t) (pred[c] + (int16_t) diff);
                        pset(pl, x, y, 0, pred[c]);
                        int z = 1;
                        while (z < 64) {
                            int kind, run, v;
                            if ((e = bits_composite_ac(bits, &b, &d[c].ac, &kind, &run, &v))) return e;
                            if (kind == 0) {
                                z += run;
                                if (!(z < 64)) break;
                                pset(pl, x, y, z, (int16_t) v);
                                z += 1;
                            } else if (run == 1) {
                                break;
                            } else {
                                return ORC_ERR_INVALID_BLOCK_RUN;
                            }
                        }
                    }
            }
    }
    return ORC_OK;
}

/* decode.swift:3295-3392  Spectral.decode (DC first, interleaved) */
static int spectral_decode_dc_first(orc_spectral *s, const bits_t *bits, blkrange blocks, int a, int ncomp,
                                    descriptor *d, int extend)
{
    if (ncomp == 1) return plane_decode_dc_first(&s->pl[d[0].p], bits, blocks, a, &d[0].dc, extend);
    int64_t r0, r1;
    rows_clamped(blocks, s->bx, extend, s->by, &r0, &r1);
    int64_t b = 0;
    int16_t pred[4] = {0, 0, 0, 0};
    for (int64_t my = r0; my < r1; ++my) {
        if (extend) {
            if (!(b < bits->count && bits_peek16(bits, b) != 0xffff)) break;
            for (int c = 0; c < ncomp; ++c) {
                int height = (int) (my + 1) * d[c].fy;
                if (height > s->pl[d[c].p].uy) plane_set_height(&s->pl[d[c].p], height);
            }
        }
        for (int mx = 0; mx < s->bx; ++mx)
            for (int c = 0; c < ncomp; ++c) {
                splane *pl = &s->pl[d[c].p];
                for (int y = (int) my * d[c].fy; y < (int) my * d[c].fy + d[c].fy; ++y)
                    for (int x = mx * d[c].fx; x < mx * d[c].fx + d[c].fx; ++x) {
                        int diff, e;
                        if ((e = bits_composite_dc(bits, &b, &d[c].dc, &diff))) return e;
                        pred[c] = (int16_t) (pred[c] + (int16_t) diff);
                        pset(pl, x, y, 0, (int16_t) ((uint16_t) pred[c] << a));
                    }
            }
    }
    return ORC_OK;
}

/* decode.swift:3395-3445  Spectral.decode (DC refine, interleaved) */
static int spectral_decode_dc_refine(orc_spectral *s, const bits_t *bits, blkrange blocks, int a, int ncomp,
                                     descriptor *d)
{
    if (ncomp == 1) return plane_decode_dc_refine(&s->pl[d[0].p], bits, blocks, a);
    int64_t r0, nrows;
    int     e;
    if ((e = rows_range2(blocks, s->bx, s->by, &r0, &nrows))) return e;
    int64_t b = 0;
    for (int64_t my = r0; my < r0 + nrows; ++my)
        for (int mx = 0; mx < s->bx; ++mx)
            for (int c = 0; c < ncomp; ++c) {
                splane *pl = &s->pl[d[c].p];
                for (int y = (int) my * d[c].fy; y < (int) my * d[c].fy + d[c].fy; ++y)
                    for (int x = mx * d[c].fx; x < mx * d[c].fx + d[c].fx; ++x) {
                        int bit;
                        if ((e = bits_refinement(bits, &b, &bit))) return e;
                        pset(pl, x, y, 0, (int16_t) (pget(pl, x, y, 0) | (int16_t) ((uint16_t) bit << a)));
                    }
            }
    return ORC_OK;
}

/* decode.swift:3476-3551  Spectral.decode(ecss:interval:scan:tables:extend:) */
API int orc_decode_scan(orc_spectral *s, int band_lo, int band_hi, int bit_lo, int bit_hi, int ncomp, const int *comps,
                        const int *dcsel, const int *acsel, const orc_huff_spec dc[4], const orc_huff_spec ac[4],
                        const uint8_t *ecs_concat, const uint64_t *ecs_offsets, int n_ecs, int64_t interval, int extend)
{
    zz_init();
    int initial = bit_hi < 0;
    int kind; /* 0 sequential, 1 dc first, 2 dc refine, 3 ac first, 4 ac refine */
    if (band_lo == 0 && band_hi == 64) {
        if (!initial) return ORC_ERR_PRECONDITION; /* fatalError("unreachable") decode.swift:3512 */
        kind = 0;
    } else if (band_lo == 0 && band_hi == 1)
        kind = initial ? 1 : 2;
    else
        kind = initial ? 3 : 4;

    descriptor d[4];
    memset(d, 0, sizeof d);
    int err = ORC_OK;
    /* table lookup + LUT build happens inside each per-ECS decode call in the reference
       (decode.swift:2884-2895, 3186-3203); an empty ecss list therefore raises nothing */
    if (n_ecs > 0) {
        for (int c = 0; c < ncomp && !err; ++c) {
            d[c].p = comps[c];
            d[c].fx = s->pl[comps[c]].fx;
            d[c].fy = s->pl[comps[c]].fy;
            if (kind == 0 || kind == 1) {
                if (!dc[dcsel[c]].present) {
                    err = ORC_ERR_UNDEFINED_DC;
                    break;
                }
                if ((err = huffdec_build(&dc[dcsel[c]], &d[c].dc))) break;
            }
            if (kind == 0 || kind == 3 || kind == 4) {
                if (!ac[acsel[c]].present) {
                    err = ORC_ERR_UNDEFINED_AC;
                    break;
                }
                if ((err = huffdec_build(&ac[acsel[c]], &d[c].ac))) break;
            }
        }
    }
    int64_t start = 0;
    for (int e = 0; e < n_ecs && !err; ++e) {
        bits_t bits;
        bits.d = ecs_concat + ecs_offsets[e];
        bits.nbytes = (int64_t) (ecs_offsets[e + 1] - ecs_offsets[e]);
        bits.count = 8 * bits.nbytes;
        blkrange blocks = {start, sat_add(start, interval)};
        switch (kind) {
        case 0: err = spectral_decode_sequential(s, &bits, blocks, ncomp, d, extend); break;
        case 1: err = spectral_decode_dc_first(s, &bits, blocks, bit_lo, ncomp, d, extend); break;
        case 2: err = spectral_decode_dc_refine(s, &bits, blocks, bit_lo, ncomp, d); break;
        case 3: err = plane_decode_ac_first(&s->pl[d[0].p], &bits, blocks, band_lo, band_hi, bit_lo, &d[0].ac); break;
        case 4: err = plane_decode_ac_refine(&s->pl[d[0].p], &bits, blocks, band_lo, band_hi, bit_lo, &d[0].ac); break;
        }
        if (interval == INT64_MAX) break; /* stride(from: 0, to: .max, by: .max) yields one element */
        start += interval;
    }
    for (int c = 0; c < 4; ++c) {
        huffdec_free(&d[c].dc);
        huffdec_free(&d[c].ac);
    }
    return err;
}

/* ========================================================================= */
/* Container: lexer + parsers + Context.decompress                            */
/* ========================================================================= */

typedef struct {
    const uint8_t *d;
    size_t         n, pos;
} lexer;

typedef struct {
    uint8_t *p;
    size_t   n, cap;
} bytebuf;
static void bb_push(bytebuf *b, uint8_t v)
{
    if (b->n == b->cap) {
        b->cap = b->cap ? 2 * b->cap : 4096;
        b->p = (uint8_t *) realloc(b->p, b->cap);
    }
    b->p[b->n++] = v;
}

/* jpeg.swift:735-809  Marker.init?(code:) -- returns 0 if invalid */
static int marker_valid(uint8_t c)
{
    if (c >= 0xc0 && c <= 0xcf) return c != 0xc8;
    if (c >= 0xd0 && c <= 0xef) return 1;
    return c == 0xfe;
}

/* decode.swift:130-190  segment(prefix:).  ecs (unstuffed) appended to `ecs` if prefix. */
static int lex_segment(lexer *lx, int prefix, bytebuf *ecs, uint8_t *marker, const uint8_t **body, size_t *body_len)
{
    while (lx->pos < lx->n) {
        uint8_t byte = lx->d[lx->pos++];
        if (byte != 0xff) {
            if (!prefix) return ORC_ERR_LEX; /* invalidMarkerSegmentPrefix */
            bb_push(ecs, byte);
            continue;
        }
        int stuffed = 0;
        do {
            if (lx->pos >= lx->n) return ORC_ERR_LEX; /* truncatedMarkerSegmentType */
            byte = lx->d[lx->pos++];
            if (byte == 0x00) {
                if (!prefix) return ORC_ERR_LEX;
                bb_push(ecs, 0xff);
                stuffed = 1;
                break;
            }
        } while (byte == 0xff);
        if (stuffed) continue;
        if (!marker_valid(byte)) return ORC_ERR_LEX; /* invalidMarkerSegmentType */
        *marker = byte;
        /* decode.swift:63-90 tail(type:) */
        if (byte == 0xd8 || byte == 0xd9 || (byte >= 0xd0 && byte <= 0xd7)) {
            *body = NULL;
            *body_len = 0;
            return ORC_OK;
        }
        if (lx->pos + 2 > lx->n) return ORC_ERR_LEX;
        size_t length = ((size_t) lx->d[lx->pos] << 8) | lx->d[lx->pos + 1];
        lx->pos += 2;
        if (length < 2) return ORC_ERR_LEX;
        if (lx->pos + (length - 2) > lx->n) return ORC_ERR_LEX;
        *body = lx->d + lx->pos;
        *body_len = length - 2;
        lx->pos += length - 2;
        return ORC_OK;
    }
    return ORC_ERR_LEX; /* truncatedEntropyCodedSegment (lexing) */
}

typedef struct {
    orc_huff_spec dc[4], ac[4];
    int           qslot[4]; /* index into spectral quanta list, -1 = empty */
    int64_t       interval; /* -1 = nil */
    orc_spectral *s;
    int           nframe_comp;
} context;

/* decode.swift:475-556  Table.parse(huffman:) */
static int parse_huffman(const uint8_t *data, size_t n, orc_huff_spec dc[4], orc_huff_spec ac[4], orc_huff_spec *pend_dc,
                         int *npend_dc, orc_huff_spec *pend_ac, int *npend_ac, int *pend_dc_t, int *pend_ac_t)
{
    (void) dc;
    (void) ac;
    size_t base = 0;
    while (base < n) {
        if (n < base + 17) return ORC_ERR_PARSE;
        size_t count = 0;
        for (int i = 0; i < 16; ++i) count += data[base + 1 + i];
        if (n < base + 17 + count) return ORC_ERR_PARSE;
        int cls = data[base] >> 4, target = data[base] & 0x0f;
        if (cls != 0 && cls != 1) return ORC_ERR_PARSE; /* invalidHuffmanTypeCode */
        if (target > 3) return ORC_ERR_PARSE;           /* invalidHuffmanTargetCode */
        orc_huff_spec t;
        memset(&t, 0, sizeof t);
        t.present = 1;
        memcpy(t.counts, data + base + 1, 16);
        /* a DHT may list up to 16*255 leaves; only 256 distinct values are meaningful */
        if (count > 256) return ORC_ERR_PARSE;
        memcpy(t.values, data + base + 17, count);
        int nn, zz;
        if (!huff_size(t.counts, &nn, &zz)) return ORC_ERR_PARSE; /* invalidHuffmanTable */
        if (cls == 0) {
            pend_dc[*npend_dc] = t;
            pend_dc_t[(*npend_dc)++] = target;
        } else {
            pend_ac[*npend_ac] = t;
            pend_ac_t[(*npend_ac)++] = target;
        }
        base += 17 + count;
        if (*npend_dc >= 60 || *npend_ac >= 60) return ORC_ERR_UNSUPPORTED;
    }
    return ORC_OK;
}

/* decode.swift:568-615  Table.parse(quantization:) + Context.push(quanta:) 3664-3672 */
/* Table.parse(quantization:) decode.swift:480-530; 16-bit tables: Quantization.init(precision:values:target:) decode.swift:431-438.
 * prec_out[i] = 0 (8-bit) / 1 (16-bit); whether a 16-bit table is acceptable depends on the format (see quanta_allowed). */
static int parse_and_push_quanta(const uint8_t *data, size_t n, uint16_t (*out)[64], int *targets, int *prec_out, int *nout)
{
    size_t base = 0;
    while (base < n) {
        int target = data[base] & 0x0f, prec = data[base] >> 4;
        if (target > 3) return ORC_ERR_PARSE;
        if (*nout >= 16) return ORC_ERR_UNSUPPORTED;
        if (prec == 0) {
            if (n < base + 65) return ORC_ERR_PARSE;
            for (int i = 0; i < 64; ++i) out[*nout][i] = data[base + 1 + i];
            base += 65;
        } else if (prec == 1) {
            if (n < base + 129) return ORC_ERR_PARSE;
            for (int i = 0; i < 64; ++i)
                out[*nout][i] = (uint16_t) ((data[base + 1 + 2 * i] << 8) | data[base + 2 + 2 * i]);
            base += 129;
        } else
            return ORC_ERR_PARSE;
        prec_out[*nout] = prec;
        targets[(*nout)++] = target;
    }
    return ORC_OK;
}
/* Spectral.push(qi:quanta:) decode.swift:2546-2557: "an 8-bit dct-based process shall not use a 16-bit quantization table"
 * -> DecodingError.invalidScanQuantizationPrecision */
static int quanta_allowed(int format_precision, int table_prec) { return table_prec == 0 || format_precision > 8; }

typedef struct {
    int id, fx, fy, qsel;
} framecomp;

/* jpeg.swift:1597-1634  Progression.update */
static int progression_update(orc_spectral *s, int band_lo, int band_hi, int bit_lo, int bit_hi, int ncomp,
                              const int *comps)
{
    int hi = bit_hi < 0 ? INT32_MAX : bit_hi;
    for (int c = 0; c < ncomp; ++c) {
        int *ap = s->approx[comps[c]];
        if (!(ap[0] < INT32_MAX || band_lo == 0)) return ORC_ERR_DECODE;
        for (int z = band_lo; z < band_hi; ++z) {
            if (!(hi == ap[z] && bit_lo < ap[z])) return ORC_ERR_DECODE;
            ap[z] = bit_lo;
        }
    }
    return ORC_OK;
}

/* decode.swift:3728-3960  Context.decompress(stream:), generic over JPEG.Format (jpeg.swift:300-340).
 * fmt_ids == NULL: JPEG.Common (jpeg.swift:370-397).  Otherwise a user-defined format in the style of
 * examples/custom-color/main.swift:41-63: recognised iff the frame's component keys are exactly fmt_ids (as a set) and
 * its precision is fmt_precision; planes are ordered as fmt_ids lists them (Layout.init(format:process:components:),
 * jpeg.swift:1286-1329). */
static orc_spectral *decompress_impl(const uint8_t *jpeg, size_t n, const int *fmt_ids, int fmt_n, int fmt_precision, int max_scans,
                                     int *err_out)
{
    zz_init();
    int           err = ORC_OK;
    lexer         lx = {jpeg, n, 0};
    uint8_t       marker = 0;
    const uint8_t *body = NULL;
    size_t        blen = 0;
    bytebuf       none = {0, 0, 0};
    orc_spectral *s = NULL;
    context       ctx;
    memset(&ctx, 0, sizeof ctx);
    ctx.interval = -1;
    for (int i = 0; i < 4; ++i) ctx.qslot[i] = -1;
    bytebuf   ecs = {0, 0, 0};
    uint64_t *offs = NULL;

#define FAIL(code)                                                                                                     \
    do {                                                                                                               \
        err = (code);                                                                                                  \
        goto done;                                                                                                     \
    } while (0)
#define NEXT()                                                                                                         \
    do {                                                                                                               \
        if ((err = lex_segment(&lx, 0, &none, &marker, &body, &blen))) goto done;                                      \
    } while (0)

    NEXT();
    if (marker != 0xd8) FAIL(ORC_ERR_DECODE); /* missingStartOfImage */
    NEXT();
    /* preamble: APPn / COM (metadata is opaque to the hot path; JFIF/EXIF payloads are not validated here) */
    while ((marker >= 0xe0 && marker <= 0xef) || marker == 0xfe) NEXT();

    /* definitions until SOF */
    orc_huff_spec pend_dc[64], pend_ac[64];
    int           pend_dc_t[64], pend_ac_t[64], npdc = 0, npac = 0;
    uint16_t      pend_q[16][64];
    int           pend_q_t[16], pend_q_prec[16], npq = 0;
    int64_t       pend_interval = -2; /* -2 = not seen */
    framecomp     fc[4];
    int           nfc = 0, process = -1, precision = 0, fw = 0, fh = 0;
    for (;;) {
        if (marker >= 0xc0 && marker <= 0xcf && marker != 0xc4 && marker != 0xc8 && marker != 0xcc) {
            /* decode.swift:784-836 Frame.parse + validate 700-768 */
            if (blen < 6) FAIL(ORC_ERR_PARSE);
            precision = body[0];
            fh = (body[1] << 8) | body[2];
            fw = (body[3] << 8) | body[4];
            int count = body[5];
            if (blen != (size_t) (3 * count + 6)) FAIL(ORC_ERR_PARSE);
            if (marker == 0xc0) process = 0;
            else if (marker == 0xc1) process = 1;
            else if (marker == 0xc2) process = 2;
            else process = 3 + marker; /* unsupported: lossless / differential / arithmetic */
            if (count > 4 && process <= 2) {
                /* Common.recognize returns nil for anything but 1 or 3 components (jpeg.swift:370-397) */
                FAIL(ORC_ERR_DECODE);
            }
            for (int i = 0; i < count && i < 4; ++i) {
                const uint8_t *b3 = body + 6 + 3 * i;
                fc[i].id = b3[0];
                fc[i].fx = b3[1] >> 4;
                fc[i].fy = b3[1] & 0x0f;
                fc[i].qsel = b3[2] & 0x0f;
                if (fc[i].qsel > 3) FAIL(ORC_ERR_PARSE);
                for (int j = 0; j < i; ++j)
                    if (fc[j].id == fc[i].id) FAIL(ORC_ERR_PARSE); /* duplicateFrameComponentIndex */
            }
            nfc = count;
            if (!(fw > 0)) FAIL(ORC_ERR_PARSE);
            for (int i = 0; i < nfc; ++i) {
                if (!(fc[i].fx >= 1 && fc[i].fx <= 4 && fc[i].fy >= 1 && fc[i].fy <= 4)) FAIL(ORC_ERR_PARSE);
                if (process == 0 && fc[i].qsel > 1) FAIL(ORC_ERR_PARSE);
            }
            if (process == 0 && precision != 8) FAIL(ORC_ERR_PARSE);
            if ((process == 1 || process == 2) && precision != 8 && precision != 12) FAIL(ORC_ERR_PARSE);
            if (process > 2) FAIL(ORC_ERR_DECODE); /* unsupportedFrameCodingProcess decode.swift:2369 */
            if (nfc < 1) FAIL(ORC_ERR_PARSE);
            NEXT();
            break;
        }
        switch (marker) {
        case 0xdb:
            if ((err = parse_and_push_quanta(body, blen, pend_q, pend_q_t, pend_q_prec, &npq))) goto done;
            break;
        case 0xc4:
            if ((err = parse_huffman(body, blen, NULL, NULL, pend_dc, &npdc, pend_ac, &npac, pend_dc_t, pend_ac_t)))
                goto done;
            break;
        case 0xdd:
            if (blen != 2) FAIL(ORC_ERR_PARSE);
            pend_interval = (body[0] << 8) | body[1];
            break;
        case 0xda: FAIL(ORC_ERR_DECODE); /* prematureScanHeaderSegment */
        case 0xdc: FAIL(ORC_ERR_DECODE);
        case 0xd9: FAIL(ORC_ERR_DECODE);
        case 0xd8: FAIL(ORC_ERR_DECODE);
        default:
            if (marker >= 0xd0 && marker <= 0xd7) FAIL(ORC_ERR_DECODE); /* unexpectedRestart */
            break; /* APPn, COM, DAC, DHP, EXP: ignored */
        }
        NEXT();
    }

    /* Context.init(frame:) -> Spectral.decode(frame:) decode.swift:2359-2388; Common.recognize jpeg.swift:370-397 */
    if (fmt_ids) {
        /* Format.recognize(_:precision:) in the style of examples/custom-color/main.swift:43-53 */
        if (precision != fmt_precision || nfc != fmt_n) FAIL(ORC_ERR_DECODE); /* unrecognizedColorFormat */
        framecomp ordered[4];
        for (int i = 0; i < fmt_n; ++i) {
            int k = -1;
            for (int j = 0; j < nfc; ++j)
                if (fc[j].id == fmt_ids[i]) k = j;
            if (k < 0) FAIL(ORC_ERR_DECODE);
            ordered[i] = fc[k];
        }
        for (int i = 0; i < fmt_n; ++i) fc[i] = ordered[i];
    } else {
        if (precision != 8) FAIL(ORC_ERR_DECODE); /* Common.recognize: precision 8 only (jpeg.swift:370-397) */
        if (!(nfc == 1 || nfc == 3)) FAIL(ORC_ERR_DECODE); /* unrecognizedColorFormat */
        /* planes are ordered by sorted component key (format.components) */
        for (int i = 0; i < nfc; ++i)
            for (int j = i + 1; j < nfc; ++j)
                if (fc[j].id < fc[i].id) {
                    framecomp t = fc[i];
                    fc[i] = fc[j];
                    fc[j] = t;
                }
        if (nfc == 3 && !(fc[1].id == fc[0].id + 1 && fc[2].id == fc[0].id + 2)) FAIL(ORC_ERR_DECODE);
    }
    {
        int factors[8];
        for (int i = 0; i < nfc; ++i) factors[2 * i] = fc[i].fx, factors[2 * i + 1] = fc[i].fy;
        s = orc_spectral_create(fw, fh, nfc, factors, process == 2);
        s->process = process;
        s->precision = precision;
        for (int i = 0; i < nfc; ++i) s->pl[i].comp_id = fc[i].id;
    }
    ctx.s = s;
    for (int i = 0; i < npdc; ++i) ctx.dc[pend_dc_t[i]] = pend_dc[i];
    for (int i = 0; i < npac; ++i) ctx.ac[pend_ac_t[i]] = pend_ac[i];
    for (int i = 0; i < npq; ++i) {
        if (!quanta_allowed(precision, pend_q_prec[i])) FAIL(ORC_ERR_DECODE); /* invalidScanQuantizationPrecision */
        ctx.qslot[pend_q_t[i]] = spectral_push_quanta(s, pend_q[i]);
    }
    if (pend_interval != -2) ctx.interval = pend_interval == 0 ? -1 : pend_interval;

    int first = 1, nscans = 0;
    for (;;) {
        if (marker >= 0xc0 && marker <= 0xcf && marker != 0xc4 && marker != 0xc8 && marker != 0xcc)
            FAIL(ORC_ERR_DECODE); /* duplicateFrameHeaderSegment */
        switch (marker) {
        case 0xdb: {
            npq = 0;
            if ((err = parse_and_push_quanta(body, blen, pend_q, pend_q_t, pend_q_prec, &npq))) goto done;
            for (int i = 0; i < npq; ++i) {
                if (!quanta_allowed(precision, pend_q_prec[i])) FAIL(ORC_ERR_DECODE);
                ctx.qslot[pend_q_t[i]] = spectral_push_quanta(s, pend_q[i]);
            }
            break;
        }
        case 0xc4: {
            npdc = npac = 0;
            if ((err = parse_huffman(body, blen, NULL, NULL, pend_dc, &npdc, pend_ac, &npac, pend_dc_t, pend_ac_t)))
                goto done;
            for (int i = 0; i < npdc; ++i) ctx.dc[pend_dc_t[i]] = pend_dc[i];
            for (int i = 0; i < npac; ++i) ctx.ac[pend_ac_t[i]] = pend_ac[i];
            break;
        }
        case 0xda: {
            /* decode.swift:946-1004 Header.Scan.parse + validate 870-930 */
            if (blen < 4) FAIL(ORC_ERR_PARSE);
            int count = body[0];
            if (blen != (size_t) (2 * count + 4)) FAIL(ORC_ERR_PARSE);
            if (count > 4) FAIL(ORC_ERR_PARSE);
            int cid[4], dcs[4], acs[4];
            for (int i = 0; i < count; ++i) {
                cid[i] = body[1 + 2 * i];
                dcs[i] = body[2 + 2 * i] >> 4;
                acs[i] = body[2 + 2 * i] & 0x0f;
                if (dcs[i] > 3 || acs[i] > 3) FAIL(ORC_ERR_PARSE);
                if (process == 0 && (dcs[i] > 1 || acs[i] > 1)) FAIL(ORC_ERR_PARSE);
            }
            int band_lo = body[2 * count + 1], band_hi = body[2 * count + 2] + 1;
            int bit_lo = body[2 * count + 3] & 0x0f;
            int bit_hi = (body[2 * count + 3] & 0xf0) == 0 ? -1 : body[2 * count + 3] >> 4;
            if (!(band_lo < band_hi && (bit_hi < 0 || bit_lo < bit_hi))) FAIL(ORC_ERR_PARSE);
            {
                int ok = 0;
                if (process != 2) ok = band_lo == 0 && band_hi == 64 && bit_lo == 0 && bit_hi < 0 && count >= 1;
                else if (band_lo == 0 && band_hi == 1) ok = (bit_hi < 0 || bit_hi == bit_lo + 1) && count >= 1;
                else if (band_lo >= 1 && band_hi >= 2 && band_hi <= 64)
                    ok = (bit_hi < 0 || bit_hi == bit_lo + 1) && count == 1;
                if (!ok) FAIL(ORC_ERR_PARSE);
            }
            /* lex ECS + RSTn  decode.swift:3895-3933 */
            ecs.n = 0;
            size_t noff = 0, capoff = 64;
            free(offs);
            offs = (uint64_t *) malloc(sizeof(uint64_t) * capoff);
            offs[noff++] = 0;
            for (int index = 0;; ++index) {
                if ((err = lex_segment(&lx, 1, &ecs, &marker, &body, &blen))) goto done;
                if (noff == capoff) offs = (uint64_t *) realloc(offs, sizeof(uint64_t) * (capoff *= 2));
                offs[noff++] = ecs.n;
                if (!(marker >= 0xd0 && marker <= 0xd7)) break;
                if ((marker & 0x0f) != index % 8) FAIL(ORC_ERR_DECODE); /* invalidRestartPhase */
            }
            int n_ecs = (int) noff - 1;
            /* Context.push(scan:ecss:extend:) decode.swift:3706-3725 */
            int64_t interval;
            if (ctx.interval > 0) interval = ctx.interval;
            else if (n_ecs == 1) interval = INT64_MAX;
            else FAIL(ORC_ERR_DECODE); /* missingRestartIntervalSegment */
            /* progression.update uses header components (by ci); Layout.push(scan:) jpeg.swift:1506-1554 */
            int comps[4], volume = 0;
            for (int i = 0; i < count; ++i) {
                int c = -1;
                for (int p = 0; p < s->ncomp; ++p)
                    if (s->pl[p].comp_id == cid[i]) c = p;
                comps[i] = c;
            }
            {
                int known[4], nk = 0;
                for (int i = 0; i < count; ++i)
                    if (comps[i] >= 0) known[nk++] = comps[i];
                if ((err = progression_update(s, band_lo, band_hi, bit_lo, bit_hi, nk, known))) goto done;
            }
            for (int i = 0; i < count; ++i) {
                if (comps[i] < 0) FAIL(ORC_ERR_DECODE); /* undefinedScanComponentReference */
                volume += s->pl[comps[i]].fx * s->pl[comps[i]].fy;
            }
            if (!((volume >= 0 && volume <= 10) || count == 1)) FAIL(ORC_ERR_DECODE);
            /* dequantize: bind quanta at each component's first DC / sequential scan  decode.swift:3451-3498 */
            if (bit_hi < 0 && band_lo == 0) {
                for (int i = 0; i < count; ++i) {
                    int sel = -1;
                    for (int k = 0; k < nfc; ++k)
                        if (fc[k].id == cid[i]) sel = fc[k].qsel;
                    if (ctx.qslot[sel] < 0) FAIL(ORC_ERR_UNDEFINED_QUANTA);
                    s->pl[comps[i]].q = ctx.qslot[sel];
                }
            }
            if ((err = orc_decode_scan(s, band_lo, band_hi, bit_lo, bit_hi, count, comps, dcs, acs, ctx.dc, ctx.ac, ecs.p,
                                       offs, n_ecs, interval, first)))
                goto done;
            if (first) {
                int height;
                if (marker == 0xdc) {
                    if (blen != 2) FAIL(ORC_ERR_PARSE);
                    height = (body[0] << 8) | body[1];
                    if (!(height > 0)) FAIL(ORC_ERR_PRECONDITION);
                    NEXT();
                } else if (fh > 0)
                    height = fh;
                else
                    FAIL(ORC_ERR_DECODE); /* missingHeightRedefinitionSegment */
                spectral_set_height(s, height);
                first = 0;
            }
            /* online decoding (examples/decode-online/main.swift:252-282): the caller looks at context.spectral after every scan */
            if (max_scans > 0 && ++nscans == max_scans) goto done;
            continue; /* `continue scans`: marker already holds the next segment */
        }
        case 0xdd:
            if (blen != 2) FAIL(ORC_ERR_PARSE);
            ctx.interval = ((body[0] << 8) | body[1]) == 0 ? -1 : ((body[0] << 8) | body[1]);
            break;
        case 0xd9: goto done; /* EOI */
        case 0xd8: FAIL(ORC_ERR_DECODE);
        case 0xdc: FAIL(ORC_ERR_DECODE); /* unexpectedHeightRedefinitionSegment */
        default:
            if (marker >= 0xd0 && marker <= 0xd7) FAIL(ORC_ERR_DECODE);
            break;
        }
        NEXT();
    }
done:
    free(ecs.p);
    free(offs);
    free(none.p);
    if (err_out) *err_out = err;
    if (err) {
        orc_spectral_free(s);
        return NULL;
    }
    return s;
#undef FAIL
#undef NEXT
}
API orc_spectral *orc_decompress(const uint8_t *jpeg, size_t n, int *err_out)
{
    return decompress_impl(jpeg, n, NULL, 0, 8, -1, err_out);
}
API orc_spectral *orc_decompress_format(const uint8_t *jpeg, size_t n, const int *format_components, int n_components,
                                        int format_precision, int *err_out)
{
    if (!format_components || n_components < 1 || n_components > 4) {
        if (err_out) *err_out = ORC_ERR_PRECONDITION;
        return NULL;
    }
    return decompress_impl(jpeg, n, format_components, n_components, format_precision, -1, err_out);
}
/* JPEG.Context driven segment by segment (decode.swift:3565-3725), stopped after `scans` scans: the state
 * examples/decode-online/main.swift hands to its capture closure */
API orc_spectral *orc_decompress_scans(const uint8_t *jpeg, size_t n, int scans, int *err_out)
{
    return decompress_impl(jpeg, n, NULL, 0, 8, scans, err_out);
}
API int  orc_spectral_precision(const orc_spectral *s) { return s->precision; }
/* a blank image of a user-defined format: component keys in plane order, sample precision */
API void orc_spectral_set_format(orc_spectral *s, const int *component_ids, int precision)
{
    for (int p = 0; p < s->ncomp; ++p) s->pl[p].comp_id = component_ids[p];
    s->precision = precision;
}

/* ========================================================================= */
/* IDCT                                                                       */
/* ========================================================================= */

static const float AAN_R[8] = {1.0f,         1.387039845f, 1.306562965f, 1.175875602f,
                               1.0f,         0.785694958f, 0.541196100f, 0.275899379f};

/* decode.swift:3984-4017 modulate(quanta:scale:): q[h][k] = (r[k] * r[h]) * (scale * Q[zz(k,h)]) */
static void modulate(const uint16_t quanta_zz[64], float scale, float q[8][8])
{
    zz_init();
    for (int h = 0; h < 8; ++h)
        for (int k = 0; k < 8; ++k) {
            float hv = AAN_R[k] * AAN_R[h];
            float row = scale * (float) quanta_zz[ZZ[h][k]];
            q[h][k] = hv * row;
        }
}

/* decode.swift:4042-4093 idct8: operates on 8 vectors h[0..7] (each 8 lanes) */
static void idct8(float h[8][8], float shift, float out[8][8])
{
    for (int j = 0; j < 8; ++j) {
        float h0 = h[0][j], h1 = h[1][j], h2 = h[2][j], h3 = h[3][j];
        float h4 = h[4][j], h5 = h[5][j], h6 = h[6][j], h7 = h[7][j];
        float a0 = (shift + h0) + h4;
        float a1 = (shift + h0) - h4;
        float b = h2 + h6;
        float c = 1.414213562f * (h2 - h6) - b;
        float r0 = a0 + b, r1 = a1 + c, r2 = a1 - c, r3 = a0 - b;
        float d0 = h5 - h3, d1 = h1 + h7, d2 = h1 - h7, d3 = h5 + h3;
        float f = 1.414213562f * (d1 - d3);
        float l = 1.847759065f * (d0 + d2);
        float m0 = l - d2 * 1.082392200f;
        float m1 = l - d0 * 2.613125930f;
        float s0 = d1 + d3;
        float s1 = m1 - s0;
        float s2 = f - s1;
        float s3 = m0 - s2;
        out[0][j] = r0 + s0;
        out[1][j] = r1 + s1;
        out[2][j] = r2 + s2;
        out[3][j] = r3 + s3;
        out[4][j] = r3 - s3;
        out[5][j] = r2 - s2;
        out[6][j] = r1 - s1;
        out[7][j] = r0 - s0;
    }
}
static void transpose8(float a[8][8], float t[8][8])
{
    for (int i = 0; i < 8; ++i)
        for (int j = 0; j < 8; ++j) t[i][j] = a[j][i];
}

/* decode.swift:4101-4133 Plane.idct(quanta:precision:) */
API void orc_idct_plane(const int16_t *coef, int ux, int uy, const uint16_t quanta_zz[64], int precision, uint16_t *out)
{
    zz_init();
    float q[8][8];
    modulate(quanta_zz, 0.125f, q);
    const size_t stride = (size_t) 8 * ux;
    const float  level = ldexpf(1.0f, precision - 1) + 0.5f;
    const float  limit = ldexpf(1.0f, precision) - 1.0f;
    for (int y = 0; y < uy; ++y)
        for (int x = 0; x < ux; ++x) {
            const int16_t *blk = coef + 64 * ((size_t) ux * y + x);
            float h[8][8], t[8][8], f[8][8], g[8][8];
            /* load: decode.swift:4020-4039  quanta * float(coef) */
            for (int hh = 0; hh < 8; ++hh)
                for (int k = 0; k < 8; ++k) h[hh][k] = q[hh][k] * (float) blk[ZZ[hh][k]];
            idct8(h, 0.0f, t);
            transpose8(t, f);
            idct8(f, level, t);
            transpose8(t, g);
            for (int i = 0; i < 8; ++i)
                for (int j = 0; j < 8; ++j) {
                    float v = g[i][j];
                    v = v < 0.0f ? 0.0f : (v > limit ? limit : v); /* clamped(lowerBound:upperBound:) */
                    if (!(v == v)) v = 0.0f;
                    out[((size_t) 8 * y + i) * stride + 8 * x + j] = (uint16_t) v; /* truncating */
                }
        }
}

/* ========================================================================= */
/* Upsample + interleave, colour                                              */
/* ========================================================================= */

/* decode.swift:4182-4276 Planar.interleaved(cosite:) */
API void orc_interleave(const uint16_t *const *planes, const int *units_xy, const int *factors_xy, int ncomp, int sx,
                        int sy, int cosited, uint16_t *out)
{
    if (ncomp == 1) {
        const size_t pw = (size_t) 8 * units_xy[0];
        for (int y = 0; y < sy; ++y)
            for (int x = 0; x < sx; ++x) out[(size_t) y * sx + x] = planes[0][x + pw * y];
        return;
    }
    int scx = 0, scy = 0;
    for (int p = 0; p < ncomp; ++p) {
        if (factors_xy[2 * p] > scx) scx = factors_xy[2 * p];
        if (factors_xy[2 * p + 1] > scy) scy = factors_xy[2 * p + 1];
    }
    for (int p = 0; p < ncomp; ++p) {
        const uint16_t *pl = planes[p];
        const int       fx = factors_xy[2 * p], fy = factors_xy[2 * p + 1];
        const size_t    pw = (size_t) 8 * units_xy[2 * p];
        const int       ph = 8 * units_xy[2 * p + 1];
        if (fx == scx && fy == scy) {
            for (int y = 0; y < sy; ++y)
                for (int x = 0; x < sx; ++x) out[((size_t) y * sx + x) * ncomp + p] = pl[x + pw * y];
            continue;
        }
        int ax, ay, bx, by, cx, cy;
        if (cosited) {
            ax = ay = 0;
            bx = fx, by = fy;
            cx = scx, cy = scy;
        } else {
            ax = fx - scx, ay = fy - scy;
            bx = 2 * fx, by = 2 * fy;
            cx = 2 * scx, cy = 2 * scy;
        }
        const int dx = (int) pw - 1, dy = ph - 1;
        for (int y = 0; y < sy; ++y)
            for (int x = 0; x < sx; ++x) {
                int   nx = ax + bx * x, ny = ay + by * y;
                int   ix = nx / cx, rx = nx % cx; /* C truncating division == quotientAndRemainder */
                int   iy = ny / cy, ry = ny % cy;
                int   jx = ix + 1 < dx ? ix + 1 : dx, jy = iy + 1 < dy ? iy + 1 : dy;
                float tx = (float) rx / (float) cx, ty = (float) ry / (float) cy;
                tx = tx < 1.0f ? tx : 1.0f;
                tx = tx > 0.0f ? tx : 0.0f;
                ty = ty < 1.0f ? ty : 1.0f;
                ty = ty > 0.0f ? ty : 0.0f;
                float u00 = (float) pl[ix + pw * iy], u01 = (float) pl[jx + pw * iy];
                float u10 = (float) pl[ix + pw * jy], u11 = (float) pl[jx + pw * jy];
                float v0 = u00 * (1.0f - tx) + u01 * tx;
                float v1 = u10 * (1.0f - tx) + u11 * tx;
                float w = v0 * (1.0f - ty) + v1 * ty;
                out[((size_t) y * sx + x) * ncomp + p] = (uint16_t) roundf(w); /* .rounded(): half away */
            }
    }
}

static inline uint8_t clamp_u8(float x)
{
    /* jpeg.swift:344-354  .init(max(0, min(x, 255))) : truncating */
    float v = x < 255.0f ? x : 255.0f;
    v = v > 0.0f ? v : 0.0f;
    return (uint8_t) v;
}

/* jpeg.swift:441-453 YCbCr.rgb ; 551-572 RGB.unpack */
API void orc_unpack_rgb(const uint16_t *il, size_t npx, int ncomp, uint8_t *rgb)
{
    for (size_t i = 0; i < npx; ++i) {
        float Y, cb, cr;
        if (ncomp == 1) {
            Y = (float) (uint8_t) il[i];
            cb = cr = 128.0f;
        } else {
            Y = (float) (uint8_t) il[3 * i];
            cb = (float) (uint8_t) il[3 * i + 1];
            cr = (float) (uint8_t) il[3 * i + 2];
        }
        float db = cb - 128.0f, dr = cr - 128.0f;
        float r = (Y + 0.00000f * db) + 1.40200f * dr;
        float g = (Y + -0.34414f * db) + -0.71414f * dr;
        float b = (Y + 1.77200f * db) + 0.00000f * dr;
        rgb[3 * i] = clamp_u8(r);
        rgb[3 * i + 1] = clamp_u8(g);
        rgb[3 * i + 2] = clamp_u8(b);
    }
}
/* jpeg.swift:493-513 YCbCr.unpack */
API void orc_unpack_ycc(const uint16_t *il, size_t npx, int ncomp, uint8_t *ycc)
{
    for (size_t i = 0; i < npx; ++i) {
        if (ncomp == 1) {
            ycc[3 * i] = (uint8_t) il[i];
            ycc[3 * i + 1] = ycc[3 * i + 2] = 128;
        } else {
            ycc[3 * i] = (uint8_t) il[3 * i];
            ycc[3 * i + 1] = (uint8_t) il[3 * i + 1];
            ycc[3 * i + 2] = (uint8_t) il[3 * i + 2];
        }
    }
}
/* jpeg.swift:463-478 RGB.ycc ; 584-599 RGB.pack */
API void orc_pack_rgb(const uint8_t *rgb, size_t npx, int ncomp, uint16_t *il)
{
    for (size_t i = 0; i < npx; ++i) {
        float R = (float) rgb[3 * i], G = (float) rgb[3 * i + 1], B = (float) rgb[3 * i + 2];
        float y = ((0.0f + 0.2990f * R) + 0.5870f * G) + 0.1140f * B;
        float cb = ((128.0f + -0.1687f * R) + -0.3313f * G) + 0.5000f * B;
        float cr = ((128.0f + 0.5000f * R) + -0.4187f * G) + -0.0813f * B;
        if (ncomp == 1)
            il[i] = clamp_u8(y);
        else {
            il[3 * i] = clamp_u8(y);
            il[3 * i + 1] = clamp_u8(cb);
            il[3 * i + 2] = clamp_u8(cr);
        }
    }
}

/* encode.swift:389-425 Rectangular.decomposed() */
API void orc_decompose_plane(const uint16_t *il, int sx, int sy, int ncomp, int p, int fx, int fy, int scx, int scy,
                             uint16_t *out)
{
    const int   ux = units_of(sx * fx, 8 * scx), uy = units_of(sy * fy, 8 * scy);
    const int   rx = scx / fx, ry = scy / fy;
    const float magnitude = (float) (rx * ry);
    for (int y = 0; y < 8 * uy; ++y)
        for (int x = 0; x < 8 * ux; ++x) {
            int     bx = x * scx / fx, by = y * scy / fy;
            int64_t sum = 0;
            for (int yy = by; yy < by + ry; ++yy)
                for (int xx = bx; xx < bx + rx; ++xx) {
                    int ix = xx < sx - 1 ? xx : sx - 1, iy = yy < sy - 1 ? yy : sy - 1;
                    sum += il[((size_t) sx * iy + ix) * ncomp + p];
                }
            out[(size_t) 8 * ux * y + x] = (uint16_t) ((float) sum / magnitude);
        }
}

/* ========================================================================= */
/* FDCT + quantise                                                            */
/* ========================================================================= */

/* encode.swift:123-188 fdct8 on 8 vectors */
static void fdct8(float g[8][8], float shift, float out[8][8])
{
    for (int j = 0; j < 8; ++j) {
        float g0 = g[0][j], g1 = g[1][j], g2 = g[2][j], g3 = g[3][j];
        float g4 = g[4][j], g5 = g[5][j], g6 = g[6][j], g7 = g[7][j];
        float a0 = g0 + g7, a1 = g1 + g6, a2 = g2 + g5, a3 = g3 + g4;
        float b0 = a0 + a3, b1 = a1 + a2, b2 = a1 - a2, b3 = a0 - a3;
        float c = 0.707106781f * (b2 + b3);
        float r0 = (b0 + b1) - shift, r1 = b3 + c, r2 = b0 - b1, r3 = b3 - c;
        float d0 = g3 - g4, d1 = g2 - g5, d2 = g1 - g6, d3 = g0 - g7;
        float f0 = d0 + d1, f1 = d1 + d2, f2 = d2 + d3;
        float k = 0.707106781f * f1;
        float l = 0.382683433f * (f0 - f2);
        float m0 = l + f0 * 0.541196100f;
        float m1 = l + f2 * 1.306562965f;
        float n0 = d3 + k, n1 = d3 - k;
        float s0 = n0 + m1, s1 = n1 - m0, s2 = n1 + m0, s3 = n0 - m1;
        out[0][j] = r0;
        out[1][j] = s0;
        out[2][j] = r1;
        out[3][j] = s1;
        out[4][j] = r2;
        out[5][j] = s2;
        out[6][j] = r3;
        out[7][j] = s3;
    }
}

/* encode.swift:199-248 Spectral.Plane.fdct */
API void orc_fdct_plane(const uint16_t *samples, int ux, int uy, const uint16_t quanta_zz[64], int precision,
                        int16_t *coef)
{
    zz_init();
    float q[8][8];
    modulate(quanta_zz, 8.0f, q);
    const float  level = ldexpf(1.0f, precision - 1) * 8.0f;
    const float  limit = ldexpf(1.0f, precision) - 1.0f;
    const size_t pw = (size_t) 8 * ux;
    for (int y = 0; y < uy; ++y)
        for (int x = 0; x < ux; ++x) {
            float g[8][8], t[8][8], f[8][8], h[8][8];
            for (int hh = 0; hh < 8; ++hh)
                for (int k = 0; k < 8; ++k) {
                    float v = (float) samples[(size_t) (8 * y + hh) * pw + 8 * x + k];
                    g[hh][k] = v < limit ? v : limit; /* pointwiseMin(limit, ...) encode.swift:86 */
                }
            transpose8(g, t);
            fdct8(t, level, f);
            transpose8(f, t);
            fdct8(t, 0.0f, h);
            int16_t *blk = coef + 64 * ((size_t) ux * y + x);
            for (int hh = 0; hh < 8; ++hh)
                for (int k = 0; k < 8; ++k) {
                    float v = h[hh][k] / q[hh][k];
                    blk[ZZ[hh][k]] = (int16_t) roundf(v); /* rounding: .toNearestOrAwayFromZero */
                }
        }
}

/* encode.swift:286-333 CompressionLevel.quanta */
API void orc_quanta(double t, int chrominance, uint16_t out[64])
{
    static const uint16_t lum[64] = {16, 11, 10, 16, 124, 140, 151, 161, 12, 12, 14, 19, 126, 158, 160, 155,
                                     14, 13, 16, 24, 140, 157, 169, 156, 14, 17, 22, 29, 151, 187, 180, 162,
                                     18, 22, 37, 56, 168, 109, 103, 177, 24, 35, 55, 64, 181, 104, 113, 192,
                                     49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 199};
    static const uint16_t chr[64] = {17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99,
                                     24, 26, 56, 99, 99, 99, 99, 99, 47, 66, 99, 99, 99, 99, 99, 99,
                                     99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99,
                                     99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99};
    zz_init();
    const uint16_t *key = chrominance ? chr : lum;
    for (int h = 0; h < 8; ++h)
        for (int k = 0; k < 8; ++k) {
            double v = round(1.0 * (1 - t) + (double) key[8 * h + k] * t); /* .rounded(): half away */
            v = v < 255.0 ? v : 255.0;
            v = v > 1.0 ? v : 1.0;
            out[ZZ[h][k]] = (uint16_t) v;
        }
}

/* ========================================================================= */
/* Huffman table construction                                                 */
/* ========================================================================= */

typedef struct {
    int64_t key;
    int     node;
} heapitem;
typedef struct {
    int left, right; /* -1 = leaf */
} treenode;

/* common.swift:127-296 General.Heap (1-based binary min-heap, strict <) */
typedef struct {
    heapitem *a; /* a[1..n] */
    int       n;
} heap_t;
static void heap_swap(heap_t *h, int i, int j)
{
    heapitem t = h->a[i];
    h->a[i] = h->a[j];
    h->a[j] = t;
}
static void heap_sift_down(heap_t *h, int i)
{
    for (;;) {
        int l = i << 1, r = (i << 1) + 1, c;
        if (!(l < h->n + 1)) return;
        if (!(r < h->n + 1)) {
            if (h->a[l].key < h->a[i].key) c = l;
            else return;
        } else {
            c = h->a[r].key < h->a[l].key ? r : l;
            if (!(h->a[c].key < h->a[i].key)) return;
        }
        heap_swap(h, i, c);
        i = c;
    }
}
static void heap_sift_up(heap_t *h, int i)
{
    for (;;) {
        int p = i >> 1;
        if (!(p >= 1)) return;
        if (!(h->a[i].key < h->a[p].key)) return;
        heap_swap(h, i, p);
        i = p;
    }
}
static void heap_enqueue(heap_t *h, int64_t key, int node)
{
    h->n += 1;
    h->a[h->n].key = key;
    h->a[h->n].node = node;
    heap_sift_up(h, h->n);
}
static int heap_dequeue(heap_t *h, heapitem *out)
{
    if (h->n == 0) return 0;
    if (h->n == 1) {
        *out = h->a[1];
        h->n = 0;
        return 1;
    }
    heap_swap(h, 1, h->n);
    *out = h->a[h->n];
    h->n -= 1;
    heap_sift_down(h, 1);
    return 1;
}

/* encode.swift:602-661 limit(height:of:) */
static int limit_levels(int height, int *levels, int count)
{
    if (!(count > height)) {
        levels[count - 1] -= 1;
        return count;
    }
    int unhoused = 0;
    for (int l = count - 1; l >= height; --l) {
        int pairs = levels[l] >> 1;
        unhoused += pairs;
        levels[l - 1] += pairs;
    }
    count = height;
    int split = height - 2;
    while (unhoused > 0) {
        if (!(levels[split] > 0)) {
            split -= 1;
            continue;
        }
        int resettled = levels[split] < unhoused ? levels[split] : unhoused;
        unhoused -= resettled;
        levels[split] -= resettled;
        levels[split + 1] += 2 * resettled;
        if (split < height - 2) split += 1;
    }
    levels[height - 1] -= 1;
    return count;
}

/* encode.swift:701-772 Table.Huffman.init(frequencies:target:) */
API void orc_huff_from_frequencies(const int64_t freq[256], orc_huff_spec *out)
{
    memset(out, 0, sizeof *out);
    int     sym[256], n = 0;
    int64_t f[256];
    for (int v = 0; v < 256; ++v)
        if (freq[v] > 0) {
            sym[n] = v;
            f[n] = freq[v];
            ++n;
        }
    if (n == 0) return;
    /* stable sort, frequency descending (insertion sort is stable) */
    for (int i = 1; i < n; ++i) {
        int     s = sym[i], j = i;
        int64_t k = f[i];
        while (j > 0 && f[j - 1] < k) {
            f[j] = f[j - 1];
            sym[j] = sym[j - 1];
            --j;
        }
        f[j] = k;
        sym[j] = s;
    }
    treenode tree[1024];
    int      ntree = 0;
    heapitem store[600];
    heap_t   heap = {store, 0};
    /* heap from sorted.reversed() via heapify */
    for (int i = 0; i < n; ++i) {
        tree[ntree].left = tree[ntree].right = -1;
        store[i + 1].key = f[n - 1 - i];
        store[i + 1].node = ntree++;
    }
    heap.n = n;
    {
        int halfway = (n >> 1) + 1; /* parent(endIndex - 1) + 1 with endIndex = n + 1 */
        for (int i = halfway - 1; i >= 1; --i) heap_sift_down(&heap, i);
    }
    tree[ntree].left = tree[ntree].right = -1;
    heap_enqueue(&heap, 0, ntree++);
    heapitem first, second;
    int      root = -1;
    while (heap_dequeue(&heap, &first)) {
        if (!heap_dequeue(&heap, &second)) {
            root = first.node;
            break;
        }
        tree[ntree].left = first.node;
        tree[ntree].right = second.node;
        heap_enqueue(&heap, first.key + second.key, ntree++);
    }
    /* levels(): BFS leaf counts per depth (encode.swift:576-595), root level dropped */
    int  levels[300], nlevels = 0;
    int *queue = (int *) malloc(sizeof(int) * 1024), *next = (int *) malloc(sizeof(int) * 1024);
    int  nq = 1;
    queue[0] = root;
    while (nq > 0) {
        int leaves = 0, nn = 0;
        for (int i = 0; i < nq; ++i) {
            if (tree[queue[i]].left < 0) leaves += 1;
            else {
                next[nn++] = tree[queue[i]].left;
                next[nn++] = tree[queue[i]].right;
            }
        }
        levels[nlevels++] = leaves;
        int *t = queue;
        queue = next;
        next = t;
        nq = nn;
    }
    free(queue);
    free(next);
    int lim[300];
    for (int i = 1; i < nlevels; ++i) lim[i - 1] = levels[i];
    int nl = limit_levels(16, lim, nlevels - 1);
    int base = 0;
    for (int l = 0; l < nl && l < 16; ++l) {
        out->counts[l] = (uint8_t) lim[l];
        for (int i = 0; i < lim[l]; ++i) out->values[base + i] = (uint8_t) sym[base + i];
        base += lim[l];
    }
    out->present = 1;
}

/* encode.swift:664-680 assign + 799-820 encoder() */
API void orc_huff_encoder(const orc_huff_spec *spec, uint16_t code[256], uint8_t len[256])
{
    memset(code, 0, 256 * sizeof(uint16_t));
    memset(len, 0, 256);
    uint16_t counter = 0;
    int      base = 0;
    for (int l = 0; l < 16; ++l) {
        for (int i = 0; i < spec->counts[l]; ++i) {
            code[spec->values[base + i]] = counter;
            len[spec->values[base + i]] = (uint8_t) (l + 1);
            counter += 1;
        }
        base += spec->counts[l];
        counter <<= 1;
    }
}

/* ========================================================================= */
/* Scan encoders                                                              */
/* ========================================================================= */

/* token stream: class 0 = DC-table symbol, 1 = AC-table symbol, 2 = raw bits, 3 = interval boundary */
typedef struct {
    uint8_t  cls, sel, symbol, nbits;
    uint16_t bits;
} token;
typedef struct {
    token *t;
    size_t n, cap;
} tokbuf;
static void tok_push(tokbuf *b, int cls, int sel, int symbol, int nbits, unsigned bits)
{
    if (b->n == b->cap) {
        b->cap = b->cap ? 2 * b->cap : 1 << 16;
        b->t = (token *) realloc(b->t, b->cap * sizeof(token));
    }
    token k = {(uint8_t) cls, (uint8_t) sel, (uint8_t) symbol, (uint8_t) nbits, (uint16_t) bits};
    b->t[b->n++] = k;
}
/* encode.swift:850-857 Composite.DC.decomposed */
static void tok_dc(tokbuf *b, int sel, int difference)
{
    int      binade;
    unsigned tail;
    orc_compact(difference, &binade, &tail);
    tok_push(b, 0, sel, binade, binade, tail);
}
/* encode.swift:859-879 Composite.AC.decomposed */
static void tok_ac_run(tokbuf *b, int sel, int zeroes, int value)
{
    int      binade;
    unsigned tail;
    orc_compact(value, &binade, &tail);
    tok_push(b, 1, sel, (zeroes << 4) | binade, binade, tail);
}
static void tok_ac_eob(tokbuf *b, int sel, int run)
{
    int binade = 0;
    while ((run >> (binade + 1)) != 0) ++binade; /* bitWidth - lzcnt - 1 */
    unsigned tail = (unsigned) (~(1 << binade) & run);
    tok_push(b, 1, sel, binade << 4, binade, tail);
}

/* encode.swift:919-959 Plane.encode(x:y:predecessor:) */
static void encode_block_sequential(tokbuf *b, const splane *pl, int x, int y, int16_t *pred, int dsel, int asel)
{
    int16_t c0 = pget(pl, x, y, 0);
    tok_dc(b, dsel, (int16_t) (c0 - *pred));
    *pred = c0;
    int zeroes = 0;
    for (int z = 1; z < 64; ++z) {
        int16_t c = pget(pl, x, y, z);
        if (c == 0) {
            if (zeroes == 15) {
                tok_ac_run(b, asel, 15, 0);
                zeroes = 0;
            } else
                zeroes += 1;
        } else {
            tok_ac_run(b, asel, zeroes, c);
            zeroes = 0;
        }
    }
    if (zeroes > 0) tok_ac_eob(b, asel, 1);
}

/* EOB-run state for the progressive AC encoders: index of the pending .eob token, or -1 */
typedef struct {
    long   last_eob;   /* token index of the last composite if it is an .eob, else -1 */
    int    count;      /* its run */
} eobstate;

static void eob_merge_or_push(tokbuf *b, eobstate *st, int sel)
{
    /* encode.swift:1092-1099 / 1176-1184 */
    if (st->last_eob >= 0 && st->count < 4096) {
        st->count += 1;
        int binade = 0;
        while ((st->count >> (binade + 1)) != 0) ++binade;
        token *t = &b->t[st->last_eob];
        t->symbol = (uint8_t) (binade << 4);
        t->nbits = (uint8_t) binade;
        t->bits = (uint16_t) (~(1 << binade) & st->count);
    } else {
        st->last_eob = (long) b->n;
        st->count = 1;
        tok_ac_eob(b, sel, 1);
    }
}

/* encode.swift:1060-1117 AC first: one block */
static void encode_block_ac_first(tokbuf *b, eobstate *st, const splane *pl, int x, int y, int lo, int hi, int a, int sel)
{
    int zeroes = 0;
    for (int z = lo; z < hi; ++z) {
        int16_t c = pget(pl, x, y, z);
        int16_t sign = c < 0 ? -1 : 1, magnitude = (int16_t) (c < 0 ? -c : c);
        int16_t high = (int16_t) (sign * (magnitude >> a));
        if (high == 0)
            zeroes += 1;
        else {
            for (int i = 0; i < zeroes / 16; ++i) tok_ac_run(b, sel, 15, 0);
            tok_ac_run(b, sel, zeroes % 16, high);
            st->last_eob = -1;
            zeroes = 0;
        }
    }
    if (zeroes > 0) eob_merge_or_push(b, st, sel);
}

/* encode.swift:1119-1205 AC refine: one block.  Refinement bits ride behind their composite as class-2 tokens;
 * when an EOB run is extended the new bits are appended after the bits already queued behind that EOB. */
static void encode_block_ac_refine(tokbuf *b, eobstate *st, const splane *pl, int x, int y, int lo, int hi, int a,
                                   int sel)
{
    const int16_t mask = (int16_t) (uint16_t) (0xffffu << (a + 1));
    uint8_t       staged[64], agg[64 * 4];
    int           nstaged = 0;
    /* refinements: list of staged lists, one per 16 zeros */
    uint8_t ref[8][64];
    int     nref = 0, reflen[8];
    int     zeroes = 0;
    for (int z = lo; z < hi; ++z) {
        int16_t c = pget(pl, x, y, z);
        int16_t sign = c < 0 ? -1 : 1, magnitude = (int16_t) (c < 0 ? -c : c);
        int16_t product = (int16_t) (magnitude & mask), remainder = (int16_t) (magnitude & ~mask);
        int16_t low = (int16_t) (sign * (remainder >> a));
        if (product == 0) {
            if (low == 0) {
                zeroes += 1;
                if (zeroes % 16 == 0) {
                    memcpy(ref[nref], staged, (size_t) nstaged);
                    reflen[nref++] = nstaged;
                    nstaged = 0;
                }
            } else {
                for (int r = 0; r < nref; ++r) {
                    tok_ac_run(b, sel, 15, 0);
                    for (int i = 0; i < reflen[r]; ++i) tok_push(b, 2, 0, 0, 1, ref[r][i]);
                }
                tok_ac_run(b, sel, zeroes % 16, low);
                for (int i = 0; i < nstaged; ++i) tok_push(b, 2, 0, 0, 1, staged[i]);
                st->last_eob = -1;
                nref = 0;
                nstaged = 0;
                zeroes = 0;
            }
        } else
            staged[nstaged++] = (uint8_t) (low != 0);
    }
    memcpy(ref[nref], staged, (size_t) nstaged);
    reflen[nref++] = nstaged;
    int nagg = 0;
    for (int r = 0; r < nref; ++r)
        for (int i = 0; i < reflen[r]; ++i) agg[nagg++] = ref[r][i];
    if (zeroes > 0 || nagg > 0) {
        eob_merge_or_push(b, st, sel);
        for (int i = 0; i < nagg; ++i) tok_push(b, 2, 0, 0, 1, agg[i]);
    }
}

/* bit writer: jpeg.swift:1920-2007 append + bytes(escaping:with:) */
typedef struct {
    bytebuf  out;
    uint32_t acc;
    int      nacc;
} bitwriter;
static void bw_put(bitwriter *w, unsigned bits, int count)
{
    for (int i = count - 1; i >= 0; --i) {
        w->acc = (w->acc << 1) | ((bits >> i) & 1u);
        if (++w->nacc == 8) {
            uint8_t by = (uint8_t) w->acc;
            bb_push(&w->out, by);
            if (by == 0xff) bb_push(&w->out, 0x00);
            w->acc = 0;
            w->nacc = 0;
        }
    }
}
static void bw_flush(bitwriter *w)
{
    if (w->nacc > 0) bw_put(w, 0xffu, 8 - w->nacc); /* pad with 1-bits */
}

/* encode.swift:1559-1620 Spectral.encode(scan:) and the ten per-kind encoders it dispatches to */
API int orc_encode_scan(orc_spectral *s, int band_lo, int band_hi, int bit_lo, int bit_hi, int ncomp, const int *comps,
                        const int *dcsel, const int *acsel, int64_t interval_mcus, orc_huff_spec dc_out[4],
                        orc_huff_spec ac_out[4], uint8_t **ecs, size_t *ecs_len)
{
    zz_init();
    int initial = bit_hi < 0, kind;
    if (band_lo == 0 && band_hi == 64) {
        if (!initial) return ORC_ERR_PRECONDITION;
        kind = 0;
    } else if (band_lo == 0 && band_hi == 1)
        kind = initial ? 1 : 2;
    else
        kind = initial ? 3 : 4;
    if (kind >= 3 && ncomp != 1) return ORC_ERR_PRECONDITION;
    const int a = bit_lo;

    /* iteration space: interleaved scans walk MCUs of the image; single-component scans walk the plane's units */
    const int interleaved = ncomp > 1;
    const int W = interleaved ? s->bx : s->pl[comps[0]].ux;
    const int H = interleaved ? s->by : s->pl[comps[0]].uy;
    int64_t   rows_per_interval = H > 0 ? H : 1;
    if (interval_mcus > 0) {
        if (interval_mcus % W != 0) return ORC_ERR_UNSUPPORTED;
        rows_per_interval = interval_mcus / W;
    }

    tokbuf   tb = {0, 0, 0};
    eobstate st = {-1, 0};
    int16_t  pred[4] = {0, 0, 0, 0};
    for (int my = 0; my < H; ++my) {
        if (my > 0 && my % rows_per_interval == 0) {
            tok_push(&tb, 3, 0, 0, 0, 0);
            st.last_eob = -1;
            pred[0] = pred[1] = pred[2] = pred[3] = 0;
        }
        for (int mx = 0; mx < W; ++mx)
            for (int c = 0; c < ncomp; ++c) {
                const splane *pl = &s->pl[comps[c]];
                const int     fx = interleaved ? pl->fx : 1, fy = interleaved ? pl->fy : 1;
                for (int y = my * fy; y < my * fy + fy; ++y)
                    for (int x = mx * fx; x < mx * fx + fx; ++x) {
                        switch (kind) {
                        case 0: encode_block_sequential(&tb, pl, x, y, &pred[c], dcsel[c], acsel[c]); break;
                        case 1: { /* encode.swift:1013-1044, 1386-1510 */
                            int16_t high = (int16_t) (pget(pl, x, y, 0) >> a);
                            tok_dc(&tb, dcsel[c], (int16_t) (high - pred[c]));
                            pred[c] = high;
                            break;
                        }
                        case 2: /* encode.swift:1046-1058, 1513-1557 */
                            tok_push(&tb, 2, 0, 0, 1, (unsigned) ((pget(pl, x, y, 0) >> a) & 1));
                            break;
                        case 3: encode_block_ac_first(&tb, &st, pl, x, y, band_lo, band_hi, a, acsel[c]); break;
                        case 4: encode_block_ac_refine(&tb, &st, pl, x, y, band_lo, band_hi, a, acsel[c]); break;
                        }
                    }
            }
    }

    /* frequencies per (class, selector) -> optimal tables */
    int64_t freq[2][4][256];
    memset(freq, 0, sizeof freq);
    int used[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
    for (int c = 0; c < ncomp; ++c) {
        if (kind == 0 || kind == 1) used[0][dcsel[c]] = 1;
        if (kind == 0 || kind == 3 || kind == 4) used[1][acsel[c]] = 1;
    }
    for (size_t i = 0; i < tb.n; ++i)
        if (tb.t[i].cls < 2) freq[tb.t[i].cls][tb.t[i].sel][tb.t[i].symbol] += 1;
    uint16_t code[2][4][256];
    uint8_t  len[2][4][256];
    for (int k = 0; k < 4; ++k) {
        memset(&dc_out[k], 0, sizeof(orc_huff_spec));
        memset(&ac_out[k], 0, sizeof(orc_huff_spec));
        if (used[0][k]) {
            orc_huff_from_frequencies(freq[0][k], &dc_out[k]);
            orc_huff_encoder(&dc_out[k], code[0][k], len[0][k]);
        }
        if (used[1][k]) {
            orc_huff_from_frequencies(freq[1][k], &ac_out[k]);
            orc_huff_encoder(&ac_out[k], code[1][k], len[1][k]);
        }
    }
    /* emit */
    bitwriter w;
    memset(&w, 0, sizeof w);
    int rst = 0;
    for (size_t i = 0; i < tb.n; ++i) {
        const token *t = &tb.t[i];
        if (t->cls == 3) {
            bw_flush(&w);
            bb_push(&w.out, 0xff);
            bb_push(&w.out, (uint8_t) (0xd0 + (rst++ & 7)));
            continue;
        }
        if (t->cls < 2) bw_put(&w, code[t->cls][t->sel][t->symbol], len[t->cls][t->sel][t->symbol]);
        bw_put(&w, t->bits, t->nbits);
    }
    bw_flush(&w);
    free(tb.t);
    *ecs = w.out.p;
    *ecs_len = w.out.n;
    return ORC_OK;
}

/* ========================================================================= */
/* Convenience pipelines                                                      */
/* ========================================================================= */

API int orc_spectral_to_planes(orc_spectral *s, uint16_t **planes)
{
    for (int p = 0; p < s->ncomp; ++p) {
        size_t n = (size_t) 64 * s->pl[p].ux * s->pl[p].uy;
        planes[p] = (uint16_t *) malloc((n ? n : 1) * sizeof(uint16_t));
        orc_idct_plane(s->pl[p].coef, s->pl[p].ux, s->pl[p].uy, s->quanta[s->pl[p].q], s->precision, planes[p]);
    }
    return ORC_OK;
}

API int orc_decode_rgb(const uint8_t *jpeg, size_t n, int *size_xy, uint8_t **rgb, uint8_t **ycc)
{
    int           err = 0;
    orc_spectral *s = orc_decompress(jpeg, n, &err);
    if (!s) return err;
    uint16_t *planes[4] = {0, 0, 0, 0};
    orc_spectral_to_planes(s, planes);
    int units[8], factors[8];
    for (int p = 0; p < s->ncomp; ++p) {
        units[2 * p] = s->pl[p].ux, units[2 * p + 1] = s->pl[p].uy;
        factors[2 * p] = s->pl[p].fx, factors[2 * p + 1] = s->pl[p].fy;
    }
    size_t    npx = (size_t) s->sx * s->sy;
    uint16_t *il = (uint16_t *) malloc((npx ? npx : 1) * (size_t) s->ncomp * sizeof(uint16_t));
    orc_interleave((const uint16_t *const *) planes, units, factors, s->ncomp, s->sx, s->sy, 0, il);
    if (rgb) {
        *rgb = (uint8_t *) malloc(3 * npx + 1);
        orc_unpack_rgb(il, npx, s->ncomp, *rgb);
    }
    if (ycc) {
        *ycc = (uint8_t *) malloc(3 * npx + 1);
        orc_unpack_ycc(il, npx, s->ncomp, *ycc);
    }
    size_xy[0] = s->sx;
    size_xy[1] = s->sy;
    for (int p = 0; p < 4; ++p) free(planes[p]);
    free(il);
    orc_spectral_free(s);
    return ORC_OK;
}

API void orc_free(void *p) { free(p); }

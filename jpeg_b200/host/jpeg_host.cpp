// jpeg_host.cpp -- see jpeg_host.hpp.  Host bookkeeping in C++; every sample / coefficient is produced by libjpeg_sm100.so.
#include "jpeg_host.hpp"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <memory>
#include <mutex>

namespace jpeg {

static inline int units_of(int size, int stride) { return size / stride + (size % stride != 0 ? 1 : 0); }

// ---------------------------------------------------------------------------------------------------------------
// Device
// ---------------------------------------------------------------------------------------------------------------
Device::Device(int index)
{
    const int rc = jpeg_sm100_create(index, &ctx_);
    if (rc != JPEG_SM100_OK) throw Error(Error::Kind::device, std::string("jpeg_sm100_create: ") + jpeg_sm100_error_string(rc), rc);
}
Device::~Device()
{
    if (ctx_) jpeg_sm100_destroy(ctx_);
}
Device &Device::shared()
{
    static std::unique_ptr<Device> d;
    static std::once_flag          once;
    std::call_once(once, [] { d.reset(new Device(0)); });
    return *d;
}
void Device::check(int status) const
{
    if (status == JPEG_SM100_OK) return;
    std::string what = jpeg_sm100_error_string(status);
    if (status == JPEG_SM100_ERR_CUDA) {
        what += ": ";
        what += jpeg_sm100_last_cuda_error(ctx_);
        throw Error(Error::Kind::device, what, status);
    }
    // -1..-8 are the DecodingError / ParsingError cases of the hot path (error.swift:495-667)
    const Error::Kind k = status == JPEG_SM100_ERR_INVALID_HUFFMAN ? Error::Kind::parsing
                          : status > -20                           ? Error::Kind::decoding
                                                                   : Error::Kind::device;
    throw Error(k, what, status);
}

// JPEG.Table.Quantization.z(k:h:) decode.swift:1289-1298
static int zigzag_of(int k, int h)
{
    const int p = (k + h < 8) ? 1 : 0, q = (k + h) & 1;
    const int a = 72 * (p ^ 1), b = 2 * p - 1, n = b * (k + h) - 14 * p + 15;
    return a + b * ((n * (n + 1)) >> 1) - q * k - (q ^ 1) * h - 1;
}

Table::Quantization CompressionLevel::quanta() const
{
    static const uint8_t lum[64] = {16, 11, 10, 16, 124, 140, 151, 161, 12, 12, 14, 19, 126, 158, 160, 155,
                                    14, 13, 16, 24, 140, 157, 169, 156, 14, 17, 22, 29, 151, 187, 180, 162,
                                    18, 22, 37, 56, 168, 109, 103, 177, 24, 35, 55, 64, 181, 104, 113, 192,
                                    49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 199};
    static const uint8_t chr[32] = {17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99,
                                    24, 26, 56, 99, 99, 99, 99, 99, 47, 66, 99, 99, 99, 99, 99, 99};
    Table::Quantization q{};
    for (int h = 0; h < 8; ++h)
        for (int k = 0; k < 8; ++k) {
            const int    i = 8 * h + k;
            const double key = kind == Kind::luminance ? lum[i] : (i < 32 ? chr[i] : 99);
            const double v = std::round(1.0 * (1 - level) + key * level);  // Double.rounded(): ties away from zero
            q[zigzag_of(k, h)] = (uint16_t) std::max(1.0, std::min(v, 255.0));
        }
    return q;
}

bool Format::recognize(std::vector<int> keys, int frame_precision) const
{
    std::vector<int> mine = components;
    std::sort(keys.begin(), keys.end());
    std::sort(mine.begin(), mine.end());
    return keys == mine && frame_precision == precision;
}

namespace Data {

// ---------------------------------------------------------------------------------------------------------------
// Spectral
// ---------------------------------------------------------------------------------------------------------------
Spectral::Spectral(std::pair<int, int> size_, const std::vector<std::pair<int, int>> &factors, const std::vector<int> &keys,
                   bool progressive_, Device *device_, int precision_)
    : progressive(progressive_), precision(precision_), device(device_ ? device_ : &Device::shared())
{
    scale = {1, 1};
    for (auto &f : factors) {
        scale.first = std::max(scale.first, f.first);
        scale.second = std::max(scale.second, f.second);
    }
    quanta.push_back(Table::Quantization{});
    planes.resize(factors.size());
    for (size_t i = 0; i < factors.size(); ++i) {
        planes[i].factor = factors[i];
        planes[i].component = keys.empty() ? (int) i + 1 : keys[i];
    }
    set_size(size_);
}

void Spectral::set_size(std::pair<int, int> s)
{
    size = s;
    blocks = {units_of(s.first, 8 * scale.first), units_of(s.second, 8 * scale.second)};
    for (auto &p : planes) {
        const int ux = units_of(s.first * p.factor.first, 8 * scale.first);
        const int uy = units_of(s.second * p.factor.second, 8 * scale.second);
        std::vector<int16_t> fresh((size_t) 64 * ux * uy, 0);
        const int            oy = std::min(uy, p.units.second), ox = std::min(ux, p.units.first);
        for (int y = 0; y < oy; ++y)
            std::memcpy(&fresh[(size_t) 64 * ux * y], &p.coefficients[(size_t) 64 * p.units.first * y], sizeof(int16_t) * 64 * ox);
        p.coefficients.swap(fresh);
        p.units = {ux, uy};
    }
}

// ---- N3: spectral-domain operations ----------------------------------------------------------------------------------
Spectral Spectral::requantized(const std::vector<Table::Quantization> &nq) const
{
    if (nq.size() != quanta.size()) throw Error(Error::Kind::decoding, "requantized: one table per quantisation slot expected");
    Spectral out = *this;
    out.quanta = nq;
    for (size_t p = 0; p < planes.size(); ++p) {
        const Plane &src = planes[p];
        if (src.coefficients.empty()) continue;
        device->check(jpeg_sm100_requantize(device->ctx(), src.coefficients.data(), (uint32_t) src.units.first, (uint32_t) src.units.second,
                                            quanta[src.q].data(), nq[src.q].data(), out.planes[p].coefficients.data()));
    }
    return out;
}

Spectral Spectral::rotated(Rotation r) const
{
    // Block.transform (examples/rotate/main.swift:13-99): per destination zig-zag index the source index and the sign
    struct Co { int z, mul; };
    std::array<Co, 64> blank, t, res;
    for (int y = 0; y < 8; ++y)
        for (int x = 0; x < 8; ++x) blank[8 * y + x] = {zigzag_of(x, y), 1};
    auto transpose = [](const std::array<Co, 64> &a) { std::array<Co, 64> o; for (int y = 0; y < 8; ++y) for (int x = 0; x < 8; ++x) o[8 * y + x] = a[8 * x + y]; return o; };
    auto reflect_v = [](const std::array<Co, 64> &a) { std::array<Co, 64> o; for (int y = 0; y < 8; ++y) for (int x = 0; x < 8; ++x) o[8 * y + x] = {a[8 * y + x].z, a[8 * y + x].mul * (1 - 2 * (y & 1))}; return o; };
    auto reflect_h = [](const std::array<Co, 64> &a) { std::array<Co, 64> o; for (int y = 0; y < 8; ++y) for (int x = 0; x < 8; ++x) o[8 * y + x] = {a[8 * y + x].z, a[8 * y + x].mul * (1 - 2 * (x & 1))}; return o; };
    int32_t matrix[4];
    int     w = size.first, h = size.second;
    switch (r) {
    case Rotation::ii: t = transpose(blank), res = reflect_v(t), matrix[0] = 0, matrix[1] = 1, matrix[2] = -1, matrix[3] = 0, w -= w % (8 * scale.first); break;
    case Rotation::iii: t = reflect_h(blank), res = reflect_v(t), matrix[0] = -1, matrix[1] = 0, matrix[2] = 0, matrix[3] = -1, w -= w % (8 * scale.first), h -= h % (8 * scale.second); break;
    default: t = transpose(blank), res = reflect_h(t), matrix[0] = 0, matrix[1] = -1, matrix[2] = 1, matrix[3] = 0, h -= h % (8 * scale.second); break;
    }
    uint8_t zmap[64];
    int8_t  mul[64];
    for (int hh = 0; hh < 8; ++hh)
        for (int k = 0; k < 8; ++k) zmap[zigzag_of(k, hh)] = (uint8_t) res[8 * hh + k].z, mul[zigzag_of(k, hh)] = (int8_t) res[8 * hh + k].mul;
    Spectral src = *this;
    src.set_size({w, h});  // original.set(width:) / set(height:): whole MCUs along the mirrored axes (main.swift:117-139)
    Spectral out = src;
    out.set_size(r == Rotation::iii ? std::make_pair(w, h) : std::make_pair(h, w));
    for (auto &q : out.quanta) {
        const Table::Quantization old = q;
        for (int z = 0; z < 64; ++z) q[z] = old[zmap[z]];
    }
    for (size_t p = 0; p < planes.size(); ++p) {
        const Plane &a = src.planes[p];
        Plane       &b = out.planes[p];
        std::fill(b.coefficients.begin(), b.coefficients.end(), (int16_t) 0);
        if (b.coefficients.empty()) continue;
        device->check(jpeg_sm100_transform_blocks(device->ctx(), a.coefficients.data(), (uint32_t) a.units.first, (uint32_t) a.units.second, matrix,
                                                  zmap, mul, b.coefficients.data(), (uint32_t) b.units.first, (uint32_t) b.units.second));
    }
    return out;
}

static jpeg_sm100_scan_desc scan_desc(const Spectral &s, const Scan &scan)
{
    jpeg_sm100_scan_desc d{};
    d.band_lo = scan.band.first;
    d.band_hi = scan.band.second;
    d.bit_lo = scan.bits.first;
    d.bit_hi = scan.bits.second < 0 ? JPEG_SM100_BITS_MAX : scan.bits.second;
    d.n_comp = (int) scan.components.size();
    for (int i = 0; i < d.n_comp; ++i) {
        const auto &c = scan.components[i];
        d.comp[i].plane = c.c;
        d.comp[i].factor_x = s.planes[c.c].factor.first;
        d.comp[i].factor_y = s.planes[c.c].factor.second;
        d.comp[i].dc = c.dc;
        d.comp[i].ac = c.ac;
    }
    d.blocks_x = s.blocks.first;
    d.blocks_y = s.blocks.second;
    return d;
}

static std::vector<jpeg_sm100_plane_i16> plane_views(const Spectral &s)
{
    std::vector<jpeg_sm100_plane_i16> v(s.planes.size());
    for (size_t i = 0; i < v.size(); ++i) {
        v[i].coef = const_cast<int16_t *>(s.planes[i].coefficients.data());
        v[i].units_x = s.planes[i].units.first;
        v[i].units_y = s.planes[i].units.second;
    }
    return v;
}

static void slots_to_array(const Table::HuffmanSlots &in, jpeg_sm100_huff_table out[4])
{
    for (int i = 0; i < 4; ++i) {
        if (in[i]) {
            out[i] = *in[i];
            out[i].present = 1;
        } else
            std::memset(&out[i], 0, sizeof out[i]);
    }
}

void Spectral::decode(const std::vector<std::vector<uint8_t>> &ecss, int64_t interval, const Scan &scan, const Table::HuffmanSlots &dc,
                      const Table::HuffmanSlots &ac, bool extend)
{
    std::vector<uint64_t> offsets(ecss.size() + 1, 0);
    for (size_t i = 0; i < ecss.size(); ++i) offsets[i + 1] = offsets[i] + ecss[i].size();
    std::vector<uint8_t> flat(offsets.back() + 8, 0);
    for (size_t i = 0; i < ecss.size(); ++i)
        if (!ecss[i].empty()) std::memcpy(&flat[offsets[i]], ecss[i].data(), ecss[i].size());
    jpeg_sm100_huff_table d[4], a[4];
    slots_to_array(dc, d);
    slots_to_array(ac, a);
    const jpeg_sm100_scan_desc desc = scan_desc(*this, scan);
    auto                       views = plane_views(*this);
    device->check(jpeg_sm100_decode_scan(device->ctx(), &desc, flat.data(), offsets.data(), (uint32_t) ecss.size(),
                                         interval < 0 ? JPEG_SM100_INTERVAL_NONE : (uint64_t) interval, extend ? 1 : 0, d, a,
                                         views.data(), (uint32_t) views.size()));
}

void Spectral::decode_raw(const uint8_t *raw, size_t n, int64_t interval, const Scan &scan, const Table::HuffmanSlots &dc,
                          const Table::HuffmanSlots &ac, bool extend)
{
    jpeg_sm100_huff_table d[4], a[4];
    slots_to_array(dc, d);
    slots_to_array(ac, a);
    const jpeg_sm100_scan_desc desc = scan_desc(*this, scan);
    auto                       views = plane_views(*this);
    device->check(jpeg_sm100_decode_scan_raw(device->ctx(), &desc, raw, n, interval < 0 ? JPEG_SM100_INTERVAL_NONE : (uint64_t) interval,
                                             extend ? 1 : 0, d, a, views.data(), (uint32_t) views.size()));
}

std::vector<uint8_t> Spectral::encode(const Scan &scan, Table::HuffmanSlots &dc, Table::HuffmanSlots &ac, uint64_t interval_mcus) const
{
    const jpeg_sm100_scan_desc desc = scan_desc(*this, scan);
    auto                       views = plane_views(*this);
    jpeg_sm100_huff_table      d[4] = {}, a[4] = {};
    size_t                     cap = 4096;
    for (auto &p : planes) cap += p.coefficients.size() * 4;
    std::vector<uint8_t> out(cap);
    uint64_t             n = 0;
    device->check(jpeg_sm100_encode_scan(device->ctx(), &desc, views.data(), (uint32_t) views.size(), interval_mcus, d, a, out.data(),
                                         cap, &n));
    out.resize(n);
    for (int i = 0; i < 4; ++i) {
        dc[i] = d[i].present ? std::optional<Table::Huffman>(d[i]) : std::nullopt;
        ac[i] = a[i].present ? std::optional<Table::Huffman>(a[i]) : std::nullopt;
    }
    return out;
}

Planar Spectral::idct() const
{
    Planar out;
    out.size = size;
    out.device = device;
    out.precision = precision;
    out.planes.resize(planes.size());
    for (size_t i = 0; i < planes.size(); ++i) {
        const auto &p = planes[i];
        auto       &o = out.planes[i];
        o.units = p.units;
        o.factor = p.factor;
        o.samples.assign((size_t) 64 * p.units.first * p.units.second, 0);
        device->check(jpeg_sm100_idct(device->ctx(), p.coefficients.data(), p.units.first, p.units.second, quanta[p.q].data(),
                                      precision, o.samples.data()));
    }
    return out;
}

std::vector<uint8_t> Spectral::to_rgb8(bool cosite) const
{
    auto                  views = plane_views(*this);
    std::vector<uint16_t> q(64 * planes.size());
    std::vector<int32_t>  f(2 * planes.size());
    for (size_t i = 0; i < planes.size(); ++i) {
        std::memcpy(&q[64 * i], quanta[planes[i].q].data(), 128);
        f[2 * i] = planes[i].factor.first;
        f[2 * i + 1] = planes[i].factor.second;
    }
    std::vector<uint8_t> rgb((size_t) 3 * size.first * size.second);
    device->check(jpeg_sm100_spectral_to_rgb8(device->ctx(), views.data(), (uint32_t) views.size(), q.data(), f.data(), size.first,
                                              size.second, cosite ? 1 : 0, rgb.data()));
    return rgb;
}

// ---------------------------------------------------------------------------------------------------------------
// Planar / Rectangular
// ---------------------------------------------------------------------------------------------------------------
static std::vector<jpeg_sm100_plane_u16> plane_views(const Planar &pl)
{
    std::vector<jpeg_sm100_plane_u16> v(pl.planes.size());
    for (size_t i = 0; i < v.size(); ++i) {
        v[i].samples = const_cast<uint16_t *>(pl.planes[i].samples.data());
        v[i].units_x = pl.planes[i].units.first;
        v[i].units_y = pl.planes[i].units.second;
        v[i].factor_x = pl.planes[i].factor.first;
        v[i].factor_y = pl.planes[i].factor.second;
    }
    return v;
}

Rectangular Planar::interleaved(bool cosite) const
{
    Rectangular r;
    r.size = size;
    r.device = device;
    r.precision = precision;
    for (auto &p : planes) r.factors.push_back(p.factor);
    r.values.assign((size_t) size.first * size.second * planes.size(), 0);
    auto views = plane_views(*this);
    device->check(jpeg_sm100_interleave(device->ctx(), views.data(), (uint32_t) views.size(), size.first, size.second, cosite ? 1 : 0,
                                        r.values.data()));
    return r;
}

Spectral Planar::fdct(const std::vector<Table::Quantization> &q, const std::vector<int> &keys) const
{
    std::vector<std::pair<int, int>> factors;
    for (auto &p : planes) factors.push_back(p.factor);
    Spectral s(size, factors, keys, false, device, precision);
    for (size_t i = 0; i < planes.size(); ++i) {
        const auto &p = planes[i];
        device->check(jpeg_sm100_fdct(device->ctx(), p.samples.data(), p.units.first, p.units.second, q[i].data(), precision,
                                      s.planes[i].coefficients.data()));
        s.quanta.push_back(q[i]);
        s.planes[i].q = (int) s.quanta.size() - 1;
    }
    return s;
}

std::vector<uint8_t> Rectangular::unpack_rgb() const
{
    std::vector<uint8_t> out((size_t) 3 * size.first * size.second);
    device->check(jpeg_sm100_unpack_rgb8(device->ctx(), values.data(), (uint64_t) size.first * size.second, stride(), out.data()));
    return out;
}
std::vector<uint8_t> Rectangular::unpack_ycc() const
{
    std::vector<uint8_t> out((size_t) 3 * size.first * size.second);
    device->check(jpeg_sm100_unpack_ycc8(device->ctx(), values.data(), (uint64_t) size.first * size.second, stride(), out.data()));
    return out;
}
Rectangular Rectangular::pack(std::pair<int, int> size, const std::vector<std::pair<int, int>> &factors, const uint8_t *rgb, Device *dev)
{
    Rectangular r;
    r.size = size;
    r.factors = factors;
    r.device = dev ? dev : &Device::shared();
    r.values.assign((size_t) size.first * size.second * factors.size(), 0);
    r.device->check(jpeg_sm100_pack_rgb8(r.device->ctx(), rgb, (uint64_t) size.first * size.second, (int) factors.size(), r.values.data()));
    return r;
}
Planar Rectangular::decomposed() const
{
    Planar pl;
    pl.size = size;
    pl.device = device;
    pl.precision = precision;
    int sx = 1, sy = 1;
    for (auto &f : factors) {
        sx = std::max(sx, f.first);
        sy = std::max(sy, f.second);
    }
    pl.planes.resize(factors.size());
    for (size_t i = 0; i < factors.size(); ++i) {
        auto &p = pl.planes[i];
        p.factor = factors[i];
        p.units = {units_of(size.first * factors[i].first, 8 * sx), units_of(size.second * factors[i].second, 8 * sy)};
        p.samples.assign((size_t) 64 * p.units.first * p.units.second, 0);
    }
    auto views = plane_views(pl);
    device->check(jpeg_sm100_decompose(device->ctx(), values.data(), size.first, size.second, views.data(), (uint32_t) views.size()));
    return pl;
}

// ---------------------------------------------------------------------------------------------------------------
// container: lexer (decode.swift:130-190), table / header parsers (475-1005), Context (3554-3961)
// ---------------------------------------------------------------------------------------------------------------
namespace {

struct Segment {
    std::vector<uint8_t> ecs;  // prefix (only with prefix = true)
    int                  marker = 0;
    const uint8_t       *body = nullptr;
    size_t               len = 0;
};

bool marker_valid(int c) { return (0xC0 <= c && c <= 0xCF && c != 0xC8) || (0xD0 <= c && c <= 0xEF) || c == 0xFE; }
bool is_frame(int m) { return 0xC0 <= m && m <= 0xCF && m != 0xC4 && m != 0xC8 && m != 0xCC; }
bool is_restart(int m) { return 0xD0 <= m && m <= 0xD7; }

[[noreturn]] void lexing(const char *w) { throw Error(Error::Kind::lexing, w); }
[[noreturn]] void parsing(const char *w) { throw Error(Error::Kind::parsing, w); }
[[noreturn]] void decoding(const char *w) { throw Error(Error::Kind::decoding, w); }

struct Lexer {
    const uint8_t *d;
    size_t         n, pos = 0;

    Segment segment(bool prefix = false)
    {
        Segment s;
        while (pos < n) {
            const uint8_t *ff = (const uint8_t *) std::memchr(d + pos, 0xFF, n - pos);
            if (!ff) break;
            const size_t k = ff - d;
            if (k > pos) {
                if (!prefix) lexing("invalidMarkerSegmentPrefix");
                s.ecs.insert(s.ecs.end(), d + pos, d + k);
            }
            pos = k + 1;
            bool stuffed = false;
            int  b = 0;
            for (;;) {
                if (pos >= n) lexing("truncatedMarkerSegmentType");
                b = d[pos++];
                if (b == 0x00) {
                    if (!prefix) lexing("invalidMarkerSegmentPrefix");
                    s.ecs.push_back(0xFF);
                    stuffed = true;
                    break;
                }
                if (b != 0xFF) break;
            }
            if (stuffed) continue;
            if (!marker_valid(b)) lexing("invalidMarkerSegmentType");
            s.marker = b;
            if (b == 0xD8 || b == 0xD9 || is_restart(b)) return s;
            if (pos + 2 > n) lexing("truncatedMarkerSegmentHeader");
            const size_t ln = ((size_t) d[pos] << 8) | d[pos + 1];
            pos += 2;
            if (ln < 2) lexing("invalidMarkerSegmentLength");
            if (pos + ln - 2 > n) lexing("truncatedMarkerSegmentBody");
            s.body = d + pos;
            s.len = ln - 2;
            pos += ln - 2;
            return s;
        }
        lexing("truncatedEntropyCodedSegment");
    }
};

struct Component {
    int key, fx, fy, tq;
};

void parse_dht(const Segment &s, Table::HuffmanSlots &dc, Table::HuffmanSlots &ac)
{
    size_t base = 0;
    while (base < s.len) {
        if (s.len < base + 17) parsing("mismatchedHuffmanSegmentSize");
        size_t total = 0;
        for (int i = 0; i < 16; ++i) total += s.body[base + 1 + i];
        if (s.len < base + 17 + total) parsing("mismatchedHuffmanSegmentSize");
        const int cls = s.body[base] >> 4, tgt = s.body[base] & 15;
        if (cls > 1) parsing("invalidHuffmanTypeCode");
        if (tgt > 3) parsing("invalidHuffmanTargetCode");
        if (total > 256) parsing("invalidHuffmanTable");
        Table::Huffman t{};
        t.present = 1;
        std::memcpy(t.counts, s.body + base + 1, 16);
        std::memcpy(t.values, s.body + base + 17, total);
        (cls == 0 ? dc : ac)[tgt] = t;
        base += 17 + total;
    }
}

// Table.parse(quantization:) decode.swift:480-530; 16-bit tables (Pq = 1): 64 big-endian UInt16, decode.swift:431-438
struct PendingQuanta {
    int                 target;
    Table::Quantization values;
    bool                wide;  // Pq = 1
};
void parse_dqt(const Segment &s, std::vector<PendingQuanta> &out)
{
    size_t base = 0;
    while (base < s.len) {
        const int tgt = s.body[base] & 15, prec = s.body[base] >> 4;
        if (tgt > 3) parsing("invalidQuantizationTargetCode");
        Table::Quantization q;
        if (prec == 0) {
            if (s.len < base + 65) parsing("mismatchedQuantizationSegmentSize");
            for (int i = 0; i < 64; ++i) q[i] = s.body[base + 1 + i];
            base += 65;
        } else if (prec == 1) {
            if (s.len < base + 129) parsing("mismatchedQuantizationSegmentSize");
            for (int i = 0; i < 64; ++i) q[i] = (uint16_t) ((s.body[base + 1 + 2 * i] << 8) | s.body[base + 2 + 2 * i]);
            base += 129;
        } else
            parsing("invalidQuantizationPrecisionCode");
        out.push_back({tgt, q, prec == 1});
    }
}
// Spectral.push(qi:quanta:) decode.swift:2546-2557: "an 8-bit dct-based process shall not use a 16-bit quantization table"
void push_quanta(Spectral &s, int qslot[4], const std::vector<PendingQuanta> &tables)
{
    for (auto &t : tables) {
        if (t.wide && s.precision <= 8) decoding("invalidScanQuantizationPrecision");
        s.quanta.push_back(t.values);
        qslot[t.target] = (int) s.quanta.size() - 1;
    }
}

int64_t parse_dri(const Segment &s)
{
    if (s.len != 2) parsing("mismatchedRestartIntervalSegmentSize");
    const int v = (s.body[0] << 8) | s.body[1];
    return v ? v : -1;
}

void put_segment(std::vector<uint8_t> &out, int marker, const std::vector<uint8_t> &tail = {})
{
    out.push_back(0xFF);
    out.push_back((uint8_t) marker);
    if (marker == 0xD8 || marker == 0xD9) return;
    const size_t ln = tail.size() + 2;
    out.push_back((uint8_t) (ln >> 8));
    out.push_back((uint8_t) ln);
    out.insert(out.end(), tail.begin(), tail.end());
}

}  // namespace

Spectral Spectral::decompress(const uint8_t *data, size_t n, Device *dev, bool gpu_lexer, const Format *format)
{
    Lexer   lx{data, n};
    Segment sg = lx.segment();
    if (sg.marker != 0xD8) decoding("missingStartOfImage");
    sg = lx.segment();
    while ((0xE0 <= sg.marker && sg.marker <= 0xEF) || sg.marker == 0xFE) sg = lx.segment();

    Table::HuffmanSlots                              dc{}, ac{};
    std::vector<PendingQuanta>                       pending;
    int64_t                                          interval = -1;
    std::vector<Component>                           comps;
    int                                              process = -1, fw = 0, fh = 0, precision = 0;
    for (;;) {
        const int m = sg.marker;
        if (is_frame(m)) {
            if (sg.len < 6) parsing("mismatchedFrameHeaderSegmentSize");
            const int count = sg.body[5];
            precision = sg.body[0];
            fh = (sg.body[1] << 8) | sg.body[2];
            fw = (sg.body[3] << 8) | sg.body[4];
            if (sg.len != (size_t) 3 * count + 6) parsing("mismatchedFrameHeaderSegmentSize");
            process = m == 0xC0 ? 0 : m == 0xC1 ? 1 : m == 0xC2 ? 2 : -1;
            for (int i = 0; i < count; ++i) {
                const uint8_t *c = sg.body + 6 + 3 * i;
                if ((c[2] & 15) > 3) parsing("invalidFrameQuantizationSelectorCode");
                for (auto &o : comps)
                    if (o.key == c[0]) parsing("duplicateFrameComponentIndex");
                comps.push_back({c[0], c[1] >> 4, c[1] & 15, c[2] & 15});
            }
            if (fw <= 0) parsing("invalidFrameWidth");
            for (auto &c : comps) {
                if (c.fx < 1 || c.fx > 4 || c.fy < 1 || c.fy > 4) parsing("invalidFrameComponentSamplingFactor");
                if (process == 0 && c.tq > 1) parsing("invalidFrameQuantizationSelector");
            }
            if (process < 0) decoding("unsupportedFrameCodingProcess");
            if ((process == 0 && precision != 8) || (precision != 8 && precision != 12)) parsing("invalidFramePrecision");  // decode.swift:700-768
            sg = lx.segment();
            break;
        }
        if (m == 0xDB)
            parse_dqt(sg, pending);
        else if (m == 0xC4)
            parse_dht(sg, dc, ac);
        else if (m == 0xDD)
            interval = parse_dri(sg);
        else if (m == 0xDA || m == 0xDC || m == 0xD9 || m == 0xD8 || is_restart(m))
            decoding("prematureSegment");
        sg = lx.segment();
    }

    if (format) {  // Format.recognize; planes in the order format.components lists them (jpeg.swift:1286-1329)
        std::vector<int> keys;
        for (auto &c : comps) keys.push_back(c.key);
        if (!format->recognize(keys, precision)) decoding("unrecognizedColorFormat");
        auto rank = [&](int key) { return std::find(format->components.begin(), format->components.end(), key) - format->components.begin(); };
        std::sort(comps.begin(), comps.end(), [&](const Component &a, const Component &b) { return rank(a.key) < rank(b.key); });
    } else {  // JPEG.Common.recognize (jpeg.swift:370-397)
        std::sort(comps.begin(), comps.end(), [](const Component &a, const Component &b) { return a.key < b.key; });
        if (precision != 8 ||
            !(comps.size() == 1 || (comps.size() == 3 && comps[1].key == comps[0].key + 1 && comps[2].key == comps[0].key + 2)))
            decoding("unrecognizedColorFormat");
    }
    std::vector<std::pair<int, int>> factors;
    std::vector<int>                 keys;
    for (auto &c : comps) {
        factors.push_back({c.fx, c.fy});
        keys.push_back(c.key);
    }
    Spectral s({fw, fh}, factors, keys, process == 2, dev, precision);
    int      qslot[4] = {-1, -1, -1, -1};
    push_quanta(s, qslot, pending);
    // JPEG.Layout progression state (jpeg.swift:1581-1634): approximation bit per coefficient, -2 = not yet seen, -1 = .max
    std::vector<std::array<int, 64>> approx(comps.size());
    for (auto &a : approx) a.fill(-2);
    auto plane_of = [&](int key) {
        for (size_t i = 0; i < comps.size(); ++i)
            if (comps[i].key == key) return (int) i;
        return -1;
    };

    bool first = true;
    for (;;) {
        const int m = sg.marker;
        if (is_frame(m)) decoding("duplicateFrameHeaderSegment");
        if (m == 0xDB) {
            std::vector<PendingQuanta> qs;
            parse_dqt(sg, qs);
            push_quanta(s, qslot, qs);
        } else if (m == 0xC4)
            parse_dht(sg, dc, ac);
        else if (m == 0xDA) {
            if (sg.len < 4 || sg.body[0] > 4 || sg.len != (size_t) 2 * sg.body[0] + 4) parsing("mismatchedScanHeaderSegmentSize");
            const int count = sg.body[0];
            struct H {
                int key, dc, ac;
            } hdr[4];
            for (int i = 0; i < count; ++i) {
                hdr[i] = {sg.body[1 + 2 * i], sg.body[2 + 2 * i] >> 4, sg.body[2 + 2 * i] & 15};
                if (hdr[i].dc > 3 || hdr[i].ac > 3 || (process == 0 && (hdr[i].dc > 1 || hdr[i].ac > 1))) parsing("invalidScanHuffmanSelector");
            }
            Scan scan;
            scan.band = {sg.body[2 * count + 1], sg.body[2 * count + 2] + 1};
            const int lo = sg.body[2 * count + 3] & 15, hi = sg.body[2 * count + 3] >> 4;
            scan.bits = {lo, hi == 0 ? -1 : hi};
            const auto &band = scan.band;
            if (!(band.first < band.second && (hi == 0 || lo < hi))) parsing("invalidScanProgressiveSubset");
            bool ok;
            if (process != 2)
                ok = band == std::make_pair(0, 64) && lo == 0 && hi == 0 && count >= 1;
            else if (band == std::make_pair(0, 1))
                ok = (hi == 0 || hi == lo + 1) && count >= 1;
            else
                ok = band.first >= 1 && band.second >= 2 && band.second <= 64 && (hi == 0 || hi == lo + 1) && count == 1;
            if (!ok) parsing("invalidScanProgressiveSubset");

            std::vector<std::vector<uint8_t>> ecss;
            const uint8_t                    *raw = lx.d + lx.pos;
            size_t                            raw_len = 0;
            int64_t                           ival;
            if (gpu_lexer) {
                // the host only finds where the scan ends (the FF of the next non-RSTn marker); unstuffing, splitting and
                // the restart-phase check run on the GPU inside decode_raw
                size_t p = lx.pos;
                for (;;) {
                    const uint8_t *ff = p < n ? (const uint8_t *) std::memchr(data + p, 0xFF, n - p) : nullptr;
                    if (!ff) lexing("truncatedEntropyCodedSegment");
                    size_t q = (size_t) (ff - data) + 1;
                    while (q < n && data[q] == 0xFF) ++q;
                    if (q >= n) lexing("truncatedMarkerSegmentType");
                    if (data[q] == 0x00 || is_restart(data[q])) {
                        p = q + 1;
                        continue;
                    }
                    raw_len = (size_t) (ff - data) - lx.pos;
                    lx.pos = (size_t) (ff - data);
                    break;
                }
                sg = lx.segment();  // the marker that ended the scan
                ival = interval;    // -1: none; decode_raw raises missingRestartIntervalSegment if it meets an RSTn then
            } else {
                for (int index = 0;; ++index) {
                    sg = lx.segment(true);
                    ecss.push_back(std::move(sg.ecs));
                    if (!is_restart(sg.marker)) break;
                    if ((sg.marker & 15) != index % 8) decoding("invalidRestartPhase");
                }
                if (interval >= 0)
                    ival = interval;
                else if (ecss.size() == 1)
                    ival = -1;
                else
                    decoding("missingRestartIntervalSegment");
            }

            for (int i = 0; i < count; ++i) {  // Progression.update (jpeg.swift:1597-1634)
                const int p = plane_of(hdr[i].key);
                if (p < 0) continue;
                auto &ap = approx[p];
                if (!(ap[0] != -2 || band.first == 0)) decoding("invalidSpectralSelectionProgression");
                for (int z = band.first; z < band.second; ++z) {
                    // first visit: scan.bits.upper must be .max; later: upper == previous lower and lower < previous lower
                    const bool fresh = ap[z] == -2;
                    if (!((fresh && hi == 0) || (!fresh && hi == ap[z] && lo < ap[z]))) decoding("invalidSuccessiveApproximationProgression");
                    ap[z] = lo;
                }
            }
            int volume = 0;
            for (int i = 0; i < count; ++i) {
                const int p = plane_of(hdr[i].key);
                if (p < 0) decoding("undefinedScanComponentReference");
                volume += s.planes[p].factor.first * s.planes[p].factor.second;
                scan.components.push_back({p, hdr[i].dc, hdr[i].ac});
            }
            if (!(volume <= 10 || count == 1)) decoding("invalidScanSamplingVolume");
            if (hi == 0 && band.first == 0)  // dequantize: decode.swift:3451-3498
                for (int i = 0; i < count; ++i) {
                    const int p = plane_of(hdr[i].key), sel = comps[p].tq;
                    if (qslot[sel] < 0) decoding("undefinedScanQuantizationReference");
                    s.planes[p].q = qslot[sel];
                }
            if (first && fh == 0) decoding("unsupported: DNL-defined height must be resolved before the first scan is placed");
            if (gpu_lexer)
                s.decode_raw(raw, raw_len, ival, scan, dc, ac, first);
            else
                s.decode(ecss, ival, scan, dc, ac, first);
            s.scans.push_back(scan);
            if (first) {
                if (sg.marker == 0xDC) {
                    if (sg.len != 2) parsing("mismatchedHeightRedefinitionSegmentSize");
                    s.set_size({fw, (sg.body[0] << 8) | sg.body[1]});
                    sg = lx.segment();
                }
                first = false;
            }
            continue;
        } else if (m == 0xDD)
            interval = parse_dri(sg);
        else if (m == 0xD9)
            return s;
        else if (m == 0xD8 || m == 0xDC || is_restart(m))
            decoding("unexpectedSegment");
        sg = lx.segment();
    }
}

// Spectral.compress(stream:) encode.swift:1918-1972.  Table slots are assigned explicitly (quantisation: plane 0 ->
// slot 0, others -> slot 1; Huffman: as named by each scan) -- the reference derives them from scan lifetimes through
// a Dictionary whose iteration order is per-process random (jpeg.swift:1388-1441), so its byte layout is not a target.
std::vector<uint8_t> Spectral::compress(uint64_t interval_mcus, bool jfif) const
{
    std::vector<uint8_t> out;
    put_segment(out, 0xD8);
    if (jfif) put_segment(out, 0xE0, {'J', 'F', 'I', 'F', 0, 1, 2, 2, 0, 1, 0, 1, 0, 0});
    std::vector<std::pair<int, int>> qs;  // (quanta index, slot)
    for (size_t i = 0; i < planes.size(); ++i) {
        bool seen = false;
        for (auto &e : qs) seen |= e.first == planes[i].q;
        if (!seen) qs.push_back({planes[i].q, (int) std::min<size_t>(i, 1)});
    }
    auto slot_of = [&](int q) {
        for (auto &e : qs)
            if (e.first == q) return e.second;
        return 0;
    };
    std::vector<uint8_t> sof = {(uint8_t) precision, (uint8_t) (size.second >> 8), (uint8_t) size.second, (uint8_t) (size.first >> 8), (uint8_t) size.first,
                                (uint8_t) planes.size()};
    for (auto &p : planes) {
        sof.push_back((uint8_t) p.component);
        sof.push_back((uint8_t) ((p.factor.first << 4) | p.factor.second));
        sof.push_back((uint8_t) slot_of(p.q));
    }
    put_segment(out, progressive ? 0xC2 : 0xC0, sof);
    std::sort(qs.begin(), qs.end(), [](auto &a, auto &b) { return a.second < b.second; });
    std::vector<uint8_t> dqt;
    for (auto &e : qs) {
        if (precision > 8) {  // decode.swift:2528-2532: formats deeper than 8 bits write 16-bit tables
            dqt.push_back((uint8_t) (0x10 | e.second));
            for (int i = 0; i < 64; ++i) {
                dqt.push_back((uint8_t) (quanta[e.first][i] >> 8));
                dqt.push_back((uint8_t) quanta[e.first][i]);
            }
        } else {
            dqt.push_back((uint8_t) e.second);
            for (int i = 0; i < 64; ++i) dqt.push_back((uint8_t) quanta[e.first][i]);
        }
    }
    put_segment(out, 0xDB, dqt);
    if (interval_mcus) put_segment(out, 0xDD, {(uint8_t) (interval_mcus >> 8), (uint8_t) interval_mcus});
    for (auto &sc : scans) {
        Table::HuffmanSlots  dc{}, ac{};
        std::vector<uint8_t> ecs = encode(sc, dc, ac, interval_mcus);
        std::vector<uint8_t> dht;
        for (int cls = 0; cls < 2; ++cls)
            for (int slot = 0; slot < 4; ++slot) {
                const auto &t = (cls == 0 ? dc : ac)[slot];
                if (!t) continue;
                dht.push_back((uint8_t) ((cls << 4) | slot));
                size_t total = 0;
                for (int i = 0; i < 16; ++i) {
                    dht.push_back(t->counts[i]);
                    total += t->counts[i];
                }
                dht.insert(dht.end(), t->values, t->values + total);
            }
        if (!dht.empty()) put_segment(out, 0xC4, dht);
        std::vector<uint8_t> sos = {(uint8_t) sc.components.size()};
        for (auto &c : sc.components) {
            sos.push_back((uint8_t) planes[c.c].component);
            sos.push_back((uint8_t) ((c.dc << 4) | c.ac));
        }
        sos.push_back((uint8_t) sc.band.first);
        sos.push_back((uint8_t) (sc.band.second - 1));
        sos.push_back((uint8_t) (((sc.bits.second < 0 ? 0 : sc.bits.second) << 4) | sc.bits.first));
        put_segment(out, 0xDA, sos);
        out.insert(out.end(), ecs.begin(), ecs.end());
    }
    put_segment(out, 0xD9);
    return out;
}

}  // namespace Data
}  // namespace jpeg

// ---------------------------------------------------------------------------------------------------------------
// flat C facade over the classes above, so the parity tests (ctypes) can drive the C++ host exactly as they drive
// the reference-shaped Python mirror.  Buffers returned through out-pointers are released with jpegh_free.
// ---------------------------------------------------------------------------------------------------------------
#include <cstdlib>

#define JPEGH_API extern "C" __attribute__((visibility("default")))

namespace {
int fail(const std::exception &e, char *err, size_t cap)
{
    if (err && cap) {
        std::strncpy(err, e.what(), cap - 1);
        err[cap - 1] = 0;
    }
    const auto *je = dynamic_cast<const jpeg::Error *>(&e);
    if (!je) return -1000;
    if (je->code) return je->code;
    return je->kind == jpeg::Error::Kind::lexing ? -201 : je->kind == jpeg::Error::Kind::parsing ? -202 : -203;
}
template <class T> T *dup(const std::vector<T> &v)
{
    T *p = (T *) std::malloc(std::max<size_t>(1, v.size() * sizeof(T)));
    if (!v.empty()) std::memcpy(p, v.data(), v.size() * sizeof(T));
    return p;
}
}  // namespace

JPEGH_API void jpegh_free(void *p) { std::free(p); }

// mode 0: staged  idct().interleaved(cosite).unpack(as: RGB)   1: same, YCbCr   2: fused to_rgb8;  +4: host lexer instead of the GPU's
JPEGH_API int jpegh_decompress_pixels8(const uint8_t *data, size_t n, int mode, int cosite, uint8_t **pixels, int32_t *w, int32_t *h,
                                       char *err, size_t errcap)
{
    try {
        auto                 s = jpeg::Data::Spectral::decompress(data, n, nullptr, (mode & 4) == 0);
        mode &= 3;
        std::vector<uint8_t> px;
        if (mode == 2)
            px = s.to_rgb8(cosite != 0);
        else {
            auto r = s.idct().interleaved(cosite != 0);
            px = mode == 0 ? r.unpack_rgb() : r.unpack_ycc();
        }
        *pixels = dup(px);
        *w = s.size.first;
        *h = s.size.second;
        return 0;
    } catch (const std::exception &e) {
        return fail(e, err, errcap);
    }
}

JPEGH_API int jpegh_decompress_coefficients(const uint8_t *data, size_t n, int32_t plane, int16_t **coef, int32_t *ux, int32_t *uy,
                                            uint16_t quanta[64], char *err, size_t errcap)
{
    try {
        auto s = jpeg::Data::Spectral::decompress(data, n);
        if (plane < 0 || plane >= (int) s.planes.size()) throw jpeg::Error(jpeg::Error::Kind::decoding, "no such plane");
        *coef = dup(s.planes[plane].coefficients);
        *ux = s.planes[plane].units.first;
        *uy = s.planes[plane].units.second;
        std::memcpy(quanta, s.quanta[s.planes[plane].q].data(), 128);
        return 0;
    } catch (const std::exception &e) {
        return fail(e, err, errcap);
    }
}

// decompress, then compress with the same progression (examples/recompress): a decodable file with identical coefficients
JPEGH_API int jpegh_recompress(const uint8_t *data, size_t n, uint64_t interval_mcus, uint8_t **out, size_t *out_n, char *err, size_t errcap)
{
    try {
        auto s = jpeg::Data::Spectral::decompress(data, n);
        auto b = s.compress(interval_mcus);
        *out = dup(b);
        *out_n = b.size();
        return 0;
    } catch (const std::exception &e) {
        return fail(e, err, errcap);
    }
}

// JPEG.CompressionLevel.luminance(level).quanta / .chrominance(level).quanta (encode.swift:286-333); no device involved
JPEGH_API void jpegh_compression_level_quanta(int32_t chrominance, double level, uint16_t quanta[64])
{
    const auto q = (chrominance ? jpeg::CompressionLevel::chrominance(level) : jpeg::CompressionLevel::luminance(level)).quanta();
    std::memcpy(quanta, q.data(), 128);
}

// The same two calls for a user-defined format (examples/custom-color): components = the format's keys in plane order.
// jpegh_recompress_format: Spectral<Format>.decompress -> compress (no JFIF segment when jfif == 0);
// jpegh_decompress_samples16: Rectangular<Format>.decompress -> the interleaved 16-bit values, stride = n_components.
JPEGH_API int jpegh_recompress_format(const uint8_t *data, size_t n, const int32_t *components, int32_t n_components, int32_t precision,
                                      int32_t jfif, uint8_t **out, size_t *out_n, char *err, size_t errcap)
{
    try {
        jpeg::Format f{std::vector<int>(components, components + n_components), precision};
        auto         s = jpeg::Data::Spectral::decompress(data, n, nullptr, true, &f);
        auto         b = s.compress(0, jfif != 0);
        *out = dup(b);
        *out_n = b.size();
        return 0;
    } catch (const std::exception &e) {
        return fail(e, err, errcap);
    }
}
JPEGH_API int jpegh_decompress_samples16(const uint8_t *data, size_t n, const int32_t *components, int32_t n_components,
                                         int32_t precision, uint16_t **values, int32_t *w, int32_t *h, char *err, size_t errcap)
{
    try {
        jpeg::Format f{std::vector<int>(components, components + n_components), precision};
        auto         s = jpeg::Data::Spectral::decompress(data, n, nullptr, true, &f);
        auto         r = s.idct().interleaved(false);
        *values = dup(r.values);
        *w = s.size.first;
        *h = s.size.second;
        return 0;
    } catch (const std::exception &e) {
        return fail(e, err, errcap);
    }
}

// examples/rotate: decompress -> rotate losslessly in the coefficient domain -> compress.  rotation: 2 (ii), 3 (iii), 4 (iv)
JPEGH_API int jpegh_rotate(const uint8_t *data, size_t n, int rotation, uint8_t **out, size_t *out_n, char *err, size_t errcap)
{
    try {
        auto s = jpeg::Data::Spectral::decompress(data, n);
        auto r = s.rotated(rotation == 2 ? jpeg::Data::Spectral::Rotation::ii
                                         : rotation == 3 ? jpeg::Data::Spectral::Rotation::iii : jpeg::Data::Spectral::Rotation::iv);
        auto b = r.compress(0);
        *out = dup(b);
        *out_n = b.size();
        return 0;
    } catch (const std::exception &e) {
        return fail(e, err, errcap);
    }
}

// Rectangular.pack -> decomposed() -> fdct(quanta:) -> compress with an explicit progression (examples/encode-basic,
// encode-advanced).  scans: n_scans x 17 int32 = band_lo, band_hi, bit_lo, bit_hi (-1 = .max), n_comp, then 4 x (plane, dc, ac)
JPEGH_API int jpegh_compress_rgb8(const uint8_t *rgb, int32_t w, int32_t h, int32_t n_planes, const int32_t *factors_xy,
                                  const uint16_t *quanta /* n_planes x 64 */, int32_t progressive, const int32_t *scans, int32_t n_scans,
                                  uint64_t interval_mcus, uint8_t **out, size_t *out_n, char *err, size_t errcap)
{
    try {
        std::vector<std::pair<int, int>>       f;
        std::vector<jpeg::Table::Quantization> q(n_planes);
        for (int i = 0; i < n_planes; ++i) {
            f.push_back({factors_xy[2 * i], factors_xy[2 * i + 1]});
            std::memcpy(q[i].data(), quanta + 64 * i, 128);
        }
        auto r = jpeg::Data::Rectangular::pack({w, h}, f, rgb);
        auto s = r.decomposed().fdct(q);
        s.progressive = progressive != 0;
        // planes with identical tables share a quanta entry (-> one DQT slot), like the reference's layout
        for (int i = 2; i < n_planes; ++i)
            if (q[i] == q[1]) s.planes[i].q = s.planes[1].q;
        for (int k = 0; k < n_scans; ++k) {
            const int32_t *e = scans + 17 * k;
            jpeg::Scan     sc;
            sc.band = {e[0], e[1]};
            sc.bits = {e[2], e[3]};
            for (int i = 0; i < e[4]; ++i) sc.components.push_back({e[5 + 3 * i], e[6 + 3 * i], e[7 + 3 * i]});
            s.scans.push_back(sc);
        }
        auto b = s.compress(interval_mcus);
        *out = dup(b);
        *out_n = b.size();
        return 0;
    } catch (const std::exception &e) {
        return fail(e, err, errcap);
    }
}

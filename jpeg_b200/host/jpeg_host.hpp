// jpeg_host.hpp -- C++ host mirror of the reference's staged interface (tayloraswift/jpeg is compiled Swift; there is no
// Swift toolchain here, so the host side above the C-ABI is written in C++).  Same names, argument meaning and error
// behaviour as the Swift types; every hot-path body is a call into libjpeg_sm100.so (include/jpeg_sm100.h).
//
//   jpeg::Data::Spectral::decompress(bytes)     JPEG.Data.Spectral.decompress(stream:)   decode.swift:4315 / Context 3728
//     .decode(ecss, interval, scan, tables, ..)   Spectral.decode(ecss:interval:scan:...)   decode.swift:3476   -> jpeg_sm100_decode_scan
//     .idct()                                     Spectral.idct()                           decode.swift:4154   -> jpeg_sm100_idct
//   jpeg::Data::Planar::interleaved(cosite)      Planar.interleaved(cosite:)               decode.swift:4182   -> jpeg_sm100_interleave
//   jpeg::Data::Rectangular::unpack_rgb/_ycc()   Rectangular.unpack(as:)                   decode.swift:4294   -> jpeg_sm100_unpack_*8
//   jpeg::Data::Rectangular::pack(rgb, ..)       Rectangular.pack(size:layout:...)         encode.swift:456    -> jpeg_sm100_pack_rgb8
//     .decomposed()                               Rectangular.decomposed()                  encode.swift:389    -> jpeg_sm100_decompose
//   jpeg::Data::Planar::fdct(quanta)             Planar.fdct(quanta:)                      encode.swift:353    -> jpeg_sm100_fdct
//   jpeg::Data::Spectral::encode(scan)           Spectral.encode(scan:)                    encode.swift:1559   -> jpeg_sm100_encode_scan
//     .compress()                                 Spectral.compress(stream:)                encode.swift:1918
//     .requantized(quanta) / .rotated(r)          examples/recompress, examples/rotate (N3)             -> jpeg_sm100_requantize / _transform_blocks
//
// Container lexing / parsing / serialisation (decode.swift:53-1005, 3554-3961; encode.swift:1623-1972) is host
// bookkeeping the reference keeps in Swift; it is restated here so whole files can be driven through the GPU path.
// No oracle, no CPU fallback: without libjpeg_sm100.so and a B200 every stage throws jpeg::Error.
#pragma once

#include <array>
#include <cstdint>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/jpeg_sm100.h"

namespace jpeg {

// JPEG.LexingError / ParsingError / DecodingError (error.swift), collapsed to a category + the reference's case name
struct Error : std::runtime_error {
    enum class Kind { lexing, parsing, decoding, device };
    Kind kind;
    int  code;  // jpeg_sm100 status for Kind::device / hot-path decoding errors, else 0
    Error(Kind k, const std::string &what, int c = 0) : std::runtime_error(what), kind(k), code(c) {}
};

// one GPU context (device + stream); shared by default, like JPEG.SM100.shared in the Swift shim
class Device {
public:
    explicit Device(int index = 0);
    ~Device();
    Device(const Device &) = delete;
    Device &operator=(const Device &) = delete;
    jpeg_sm100_ctx *ctx() const { return ctx_; }
    void            check(int status) const;  // status -> jpeg::Error, mapping the DecodingError cases
    static Device  &shared();

private:
    jpeg_sm100_ctx *ctx_ = nullptr;
};

namespace Table {
using Huffman = jpeg_sm100_huff_table;                          // BITS + HUFFVAL (jpeg.swift:998-1012)
using HuffmanSlots = std::array<std::optional<Huffman>, 4>;     // (Delegate?, Delegate?, Delegate?, Delegate?)
using Quantization = std::array<uint16_t, 64>;                  // zig-zag order (jpeg.swift:1022-1073)
}  // namespace Table

// JPEG.Scan (jpeg.swift:1135-1197): band, bits, components as (plane index, dc selector, ac selector)
struct Scan {
    struct Component {
        int c, dc, ac;
    };
    std::pair<int, int>    band;  // lower ..< upper
    std::pair<int, int>    bits;  // lower ..< upper, upper = -1 for .max
    std::vector<Component> components;
};

// JPEG.CompressionLevel (encode.swift:260-333): quantum values for a quality parameter (0.0 = all ones, 1.0 = the keyframe
// table), in zig-zag order.  Host-side Double arithmetic.
struct CompressionLevel {
    enum class Kind { luminance, chrominance } kind;
    double level;
    static CompressionLevel luminance(double level) { return {Kind::luminance, level}; }
    static CompressionLevel chrominance(double level) { return {Kind::chrominance, level}; }
    Table::Quantization     quanta() const;
};

// A user-defined JPEG.Format (jpeg.swift:300-340) in the style of examples/custom-color/main.swift:41-63: recognised iff the
// frame's component keys are exactly `components` and its precision is `precision`; `components` orders the planes.
// Where a `const Format *` is expected, nullptr means JPEG.Common (jpeg.swift:370-397).
struct Format {
    std::vector<int> components;
    int              precision = 8;
    bool             recognize(std::vector<int> keys, int frame_precision) const;
};

namespace Data {

class Planar;
class Rectangular;

// JPEG.Data.Spectral<JPEG.Common> (decode.swift:1397-1519)
class Spectral {
public:
    struct Plane {
        std::pair<int, int>  units{0, 0};
        std::pair<int, int>  factor{1, 1};
        int                  q = 0;        // index into `quanta`
        int                  component = 0;  // component key
        std::vector<int16_t> coefficients;   // 64 * (units.x * y + x) + z
    };

    Spectral(std::pair<int, int> size, const std::vector<std::pair<int, int>> &factors, const std::vector<int> &keys = {},
             bool progressive = false, Device *device = nullptr, int precision = 8);

    std::pair<int, int>              size{0, 0}, blocks{0, 0}, scale{1, 1};
    bool                             progressive = false;
    int                              precision = 8;  // Format.precision: 8 for JPEG.Common, 12 for deeper user-defined formats
    std::vector<Plane>               planes;
    std::vector<Table::Quantization> quanta;  // quanta[0] = default (zeros), decode.swift:1723
    std::vector<Scan>                scans;   // the progression this image was decoded from / is encoded with

    void set_size(std::pair<int, int> size);  // set(width:) + set(height:) decode.swift:2456-2495

    // decode.swift:3476: one scan, already lexed (unstuffed, split at RSTn); interval < 0 = no DRI
    void decode(const std::vector<std::vector<uint8_t>> &ecss, int64_t interval, const Scan &scan,
                const Table::HuffmanSlots &dc, const Table::HuffmanSlots &ac, bool extend);
    // the same scan from its raw bytes (stuffed, RSTn-delimited): lexing (decode.swift:130-190, 3895-3933) happens on the GPU
    void decode_raw(const uint8_t *raw, size_t n, int64_t interval, const Scan &scan, const Table::HuffmanSlots &dc,
                    const Table::HuffmanSlots &ac, bool extend);
    // encode.swift:1559: returns the stuffed entropy-coded segment; tables by slot; interval_mcus = 0 -> reference form
    std::vector<uint8_t> encode(const Scan &scan, Table::HuffmanSlots &dc, Table::HuffmanSlots &ac, uint64_t interval_mcus = 0) const;

    // spectral-domain operations (N3): the loops of examples/recompress/main.swift:35-58 and examples/rotate/main.swift:101-199
    Spectral requantized(const std::vector<Table::Quantization> &new_quanta) const;  // one table per entry of `quanta`
    enum class Rotation { ii, iii, iv };                                             // quadrant the x axis is rotated into
    Spectral rotated(Rotation r) const;

    Planar               idct() const;                       // decode.swift:4154
    std::vector<uint8_t> to_rgb8(bool cosite = false) const;  // fused idct().interleaved().unpack(as: RGB.self)

    // decode.swift:3728-3960; gpu_lexer = false keeps byte unstuffing / RSTn splitting on the host (Bytestream.segment)
    static Spectral      decompress(const uint8_t *data, size_t n, Device *device = nullptr, bool gpu_lexer = true,
                                    const Format *format = nullptr);
    std::vector<uint8_t> compress(uint64_t interval_mcus = 0, bool jfif = true) const;         // encode.swift:1918-1972

    Device *device;
};

// JPEG.Data.Planar<JPEG.Common> (decode.swift:1543-1632): uint16 samples, x + 8 * units.x * y
class Planar {
public:
    struct Plane {
        std::pair<int, int>   units, factor;
        std::vector<uint16_t> samples;
    };
    std::pair<int, int> size;
    std::vector<Plane>  planes;
    Device             *device;
    int                 precision = 8;

    Rectangular interleaved(bool cosite = false) const;                          // decode.swift:4182
    // encode.swift:353 (one table per plane); keys: the component key of every plane (default 1, 2, ...)
    Spectral    fdct(const std::vector<Table::Quantization> &quanta, const std::vector<int> &keys = {}) const;
};

// JPEG.Data.Rectangular<JPEG.Common> (decode.swift:1650-1718): uint16 values, (y * size.x + x) * stride + p
class Rectangular {
public:
    std::pair<int, int>               size;
    std::vector<std::pair<int, int>>  factors;
    std::vector<uint16_t>             values;
    Device                           *device;
    int                               precision = 8;

    int                  stride() const { return (int) factors.size(); }
    std::vector<uint8_t> unpack_rgb() const;  // [JPEG.RGB]   jpeg.swift:551
    std::vector<uint8_t> unpack_ycc() const;  // [JPEG.YCbCr] jpeg.swift:493
    static Rectangular   pack(std::pair<int, int> size, const std::vector<std::pair<int, int>> &factors, const uint8_t *rgb,
                              Device *device = nullptr);  // encode.swift:456 with RGB pixels
    Planar               decomposed() const;              // encode.swift:389
};

}  // namespace Data
}  // namespace jpeg

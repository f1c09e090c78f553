// JPEGSM100Shim.swift -- the reference-side binding of libjpeg_sm100.so.
//
// This file goes INSIDE the `JPEG` module of tayloraswift/jpeg (sources/jpeg/), next to decode.swift, because the
// plane buffers it hands to C are `private`/internal there (decode.swift:1433, 1575; set(values:units:) 2225;
// Planar.Plane.init(_:units:factor:) 4137).  It cannot be compiled in the build container (no Swift toolchain);
// the identical symbols are exercised through ctypes by tests/ (jpeg_b200/lib.py, jpeg_b200/host.py).
//
// Build: add a system-library target `CJPEGSM100` whose module.modulemap is
//     module CJPEGSM100 [system] { header "jpeg_sm100.h" link "jpeg_sm100" export * }
// and `dependencies: ["CJPEGSM100"]` on the `JPEG` target in Package.swift (Package.swift:27).
//
// What changes in the reference: the BODIES of five functions become one-line calls into this file
// (see INTEGRATION.md for the diff); every public type and signature stays as it is.

import CJPEGSM100

extension JPEG
{
    /// One GPU context per thread of use (the library is re-entrant per ctx, a ctx is not thread-safe).
    final class SM100
    {
        let ctx:OpaquePointer
        init(device:Int32 = 0) throws
        {
            var ctx:OpaquePointer?
            let status:Int32 = jpeg_sm100_create(device, &ctx)
            guard status == 0, let ctx:OpaquePointer = ctx
            else
            {
                throw JPEG.SM100Error.cuda(status)
            }
            self.ctx = ctx
        }
        deinit
        {
            jpeg_sm100_destroy(self.ctx)
        }

        static let shared:SM100 = try! .init()
    }

    enum SM100Error:Swift.Error
    {
        case cuda(Int32)
        case usage(Int32)
        case batch(Int32)
    }
}

extension JPEG.SM100
{
    /// C status -> the error the Swift decoder would have thrown at the same point.
    static func check(_ status:Int32) throws
    {
        switch status
        {
        case 0:
            return
        case -1:
            throw JPEG.DecodingError.truncatedEntropyCodedSegment
        case -2:
            throw JPEG.DecodingError.invalidCompositeValue(0, expected: -1 ... 1)
        case -3:
            throw JPEG.DecodingError.invalidCompositeBlockRun(0, expected: 1 ... 1)
        case -4:
            throw JPEG.DecodingError.undefinedScanHuffmanDCReference(\.0)
        case -5:
            throw JPEG.DecodingError.undefinedScanHuffmanACReference(\.0)
        case -7:
            preconditionFailure("libjpeg_sm100: the reference implementation traps on this input")
        case -8:
            throw JPEG.ParsingError.invalidHuffmanTable
        case -9:
            // the C-ABI reports the first offending interval only; the phases themselves stay on the device
            throw JPEG.DecodingError.invalidRestartPhase(-1, expected: -1)
        case -10:
            // batch lexer only: an image whose RSTn count differs from the batch geometry has no Swift counterpart
            throw JPEG.SM100Error.batch(status)
        case -11:
            throw JPEG.DecodingError.missingRestartIntervalSegment
        case -20, -21, -22:
            throw JPEG.SM100Error.usage(status)
        default: // -100: device missing / launch failure
            throw JPEG.SM100Error.cuda(status)
        }
    }
}

// MARK: tables

extension JPEG.Table.Huffman
{
    /// BITS + HUFFVAL exactly as serialized() writes them (encode.swift:1647-1650)
    var sm100:jpeg_sm100_huff_table
    {
        var table:jpeg_sm100_huff_table = .init()
        table.present = 1
        withUnsafeMutableBytes(of: &table.counts)
        {
            for (l, leaves):(Int, [Symbol]) in self.symbols.enumerated()
            {
                $0[l] = .init(leaves.count)
            }
        }
        withUnsafeMutableBytes(of: &table.values)
        {
            for (i, symbol):(Int, Symbol) in self.symbols.joined().enumerated()
            {
                $0[i] = symbol.value
            }
        }
        return table
    }
}

func sm100Slots<Symbol>(_ slots:JPEG.Table.Huffman<Symbol>.Slots) -> [jpeg_sm100_huff_table]
{
    [slots.0, slots.1, slots.2, slots.3].map{ $0?.sm100 ?? .init() }
}

// MARK: decode seams

extension JPEG.Data.Spectral
{
    /// replaces the body of decode(ecss:interval:scan:tables:extend:) after `layout.push(scan:)` and
    /// `dequantize(components:tables:)` (decode.swift:3484-3498), i.e. the loop at decode.swift:3500-3550.
    mutating
    func sm100Decode(ecss:[[UInt8]], interval:Int, scan:JPEG.Scan,
        tables slots:(dc:JPEG.Table.HuffmanDC.Slots, ac:JPEG.Table.HuffmanAC.Slots), extend:Bool) throws
    {
        var descriptor:jpeg_sm100_scan_desc = .init()
        descriptor.band_lo  = .init(scan.band.lowerBound)
        descriptor.band_hi  = .init(scan.band.upperBound)
        descriptor.bit_lo   = .init(scan.bits.lowerBound)
        descriptor.bit_hi   = scan.bits.upperBound == .max ? -1 : .init(scan.bits.upperBound)
        descriptor.n_comp   = .init(scan.components.count)
        descriptor.blocks_x = .init(self.blocks.x)
        descriptor.blocks_y = .init(self.blocks.y)
        withUnsafeMutablePointer(to: &descriptor.comp)
        {
            $0.withMemoryRebound(to: jpeg_sm100_scan_comp.self, capacity: 4)
            {
                for (i, (c, component)):(Int, (Int, JPEG.Scan.Component)) in scan.components.enumerated()
                {
                    let factor:(x:Int, y:Int) = self.layout.planes[c].component.factor
                    $0[i].plane     = self.indices ~= c ? .init(c) : -1
                    $0[i].factor_x  = .init(factor.x)
                    $0[i].factor_y  = .init(factor.y)
                    $0[i].dc        = .init(JPEG.Table.HuffmanDC.serialize(selector: component.selector.dc))
                    $0[i].ac        = .init(JPEG.Table.HuffmanAC.serialize(selector: component.selector.ac))
                }
            }
        }

        let flat:[UInt8]        = .init(ecss.joined())
        var offsets:[UInt64]    = [0]
        for ecs:[UInt8] in ecss
        {
            offsets.append(offsets.last! + .init(ecs.count))
        }
        let dc:[jpeg_sm100_huff_table] = sm100Slots(slots.dc),
            ac:[jpeg_sm100_huff_table] = sm100Slots(slots.ac)

        // hand the planes' own storage to the library (in/out: progressive scans read-modify-write)
        var buffers:[[Int16]] = self.indices.map{ self[$0].takeBuffer() }
        defer
        {
            for p:Int in self.indices
            {
                self[p].set(values: buffers[p], units: self[p].units)
            }
        }
        var planes:[jpeg_sm100_plane_i16] = []
        for p:Int in self.indices
        {
            planes.append(.init(coef: nil, units_x: .init(self[p].units.x), units_y: .init(self[p].units.y)))
        }
        try JPEG.SM100.withPointers(&buffers, &planes)
        {
            try JPEG.SM100.check(jpeg_sm100_decode_scan(JPEG.SM100.shared.ctx, &descriptor,
                flat, offsets, .init(ecss.count),
                interval == .max ? UInt64.max : .init(interval), extend ? 1 : 0,
                dc, ac, $0, .init(planes.count)))
        }
    }
}

extension JPEG.Data.Spectral.Plane
{
    /// replaces the body of idct(quanta:precision:) (decode.swift:4101-4133)
    func sm100IDCT(quanta table:JPEG.Table.Quantization, precision:Int) throws -> JPEG.Data.Planar<Format>.Plane
    {
        let count:Int       = 64 * self.units.x * self.units.y
        let values:[UInt16] = try .init(unsafeUninitializedCapacity: count)
        {
            (samples:inout UnsafeMutableBufferPointer<UInt16>, initialized:inout Int) in
            try self.withUnsafeCoefficients
            {
                try JPEG.SM100.check(jpeg_sm100_idct(JPEG.SM100.shared.ctx, $0.baseAddress,
                    .init(self.units.x), .init(self.units.y), table.storage, .init(precision), samples.baseAddress))
            }
            initialized = count
        }
        return .init(values, units: self.units, factor: self.factor)
    }
}

extension JPEG.Data.Planar
{
    /// replaces the body of interleaved(cosite:) (decode.swift:4182-4276)
    func sm100Interleaved(cosite cosited:Bool) throws -> JPEG.Data.Rectangular<Format>
    {
        let count:Int = self.size.x * self.size.y * self.count
        let interleaved:[UInt16] = try .init(unsafeUninitializedCapacity: count)
        {
            (out:inout UnsafeMutableBufferPointer<UInt16>, initialized:inout Int) in
            try self.withUnsafePlanes
            {
                (planes:[jpeg_sm100_plane_u16]) in
                try JPEG.SM100.check(jpeg_sm100_interleave(JPEG.SM100.shared.ctx, planes, .init(planes.count),
                    .init(self.size.x), .init(self.size.y), cosited ? 1 : 0, out.baseAddress))
            }
            initialized = count
        }
        return .init(size: self.size, layout: self.layout, metadata: self.metadata, values: interleaved)
    }
}

extension JPEG.RGB
{
    /// replaces the body of unpack(_:of:) (jpeg.swift:551-572)
    static func sm100Unpack(_ interleaved:[UInt16], of format:JPEG.Common) throws -> [Self]
    {
        let arity:Int32 = format.components.count == 1 ? 1 : 3
        let pixels:Int  = interleaved.count / .init(arity)
        return try .init(unsafeUninitializedCapacity: pixels)
        {
            (rgb:inout UnsafeMutableBufferPointer<Self>, initialized:inout Int) in
            // JPEG.RGB is @frozen with three UInt8 stored properties: 3 bytes per element, no padding
            try rgb.withMemoryRebound(to: UInt8.self)
            {
                try JPEG.SM100.check(jpeg_sm100_unpack_rgb8(JPEG.SM100.shared.ctx, interleaved, .init(pixels), arity,
                    $0.baseAddress))
            }
            initialized = pixels
        }
    }
}

// MARK: the image resident on the device
//
// JPEG.Context keeps ONE Spectral for the whole file and pushes every scan into it (decode.swift:3565-3587, 3706-3725).  With the
// per-call seams above every scan uploads and downloads all planes; with a resident image only the scan's bytes travel, the
// stages read the planes where they are, and the host planes are materialised once, on demand.  `Spectral` gains ONE stored
// property, `var device:JPEG.SM100.Resident?` (a final class: value-type copies of a Spectral share it, copy-on-write is
// restored by `isKnownUniquelyReferenced` in the mutating paths), created by `Spectral.init(layout:)` when a context exists.
extension JPEG.SM100
{
    final class Resident
    {
        let handle:OpaquePointer
        var hostIsCurrent:Bool = true     // the Swift planes equal the device planes (true right after creation: all zero)
        init(units:[(x:Int, y:Int)]) throws
        {
            var handle:OpaquePointer?
            let flat:[Int32] = units.flatMap{ [Int32.init($0.x), Int32.init($0.y)] }
            try JPEG.SM100.check(jpeg_sm100_spectral_create(JPEG.SM100.shared.ctx, .init(units.count), flat, &handle))
            self.handle = handle!
        }
        deinit
        {
            jpeg_sm100_spectral_destroy(JPEG.SM100.shared.ctx, self.handle)
        }
    }
}
extension JPEG.Data.Spectral
{
    /// decode(ecss:interval:scan:tables:extend:) into the resident image; `descriptor`, `flat`, `offsets`, `dc`, `ac` built as in
    /// sm100Decode above.  The Swift planes go stale (`hostIsCurrent = false`) instead of being rewritten.
    mutating
    func sm100DecodeResident(_ device:JPEG.SM100.Resident, descriptor:inout jpeg_sm100_scan_desc, flat:[UInt8], offsets:[UInt64],
        interval:Int, extend:Bool, dc:[jpeg_sm100_huff_table], ac:[jpeg_sm100_huff_table]) throws
    {
        device.hostIsCurrent = false
        try JPEG.SM100.check(jpeg_sm100_spectral_decode_scan(JPEG.SM100.shared.ctx, device.handle, &descriptor,
            flat, offsets, .init(offsets.count - 1), interval == .max ? UInt64.max : .init(interval), extend ? 1 : 0, dc, ac))
    }
    /// Spectral.set(width:) / set(height:) (decode.swift:2456-2495), e.g. Context.push(height:) after a DNL segment
    mutating
    func sm100Resize(_ device:JPEG.SM100.Resident, units:[(x:Int, y:Int)]) throws
    {
        let flat:[Int32] = units.flatMap{ [Int32.init($0.x), Int32.init($0.y)] }
        try JPEG.SM100.check(jpeg_sm100_spectral_resize(JPEG.SM100.shared.ctx, device.handle, flat))
    }
    /// called by every accessor of `Plane.buffer` (subscripts, with(ci:), serialisation) before it reads the Swift planes
    mutating
    func sm100Materialize(_ device:JPEG.SM100.Resident) throws
    {
        guard !device.hostIsCurrent
        else
        {
            return
        }
        var buffers:[[Int16]] = self.indices.map{ .init(repeating: 0, count: 64 * self[$0].units.x * self[$0].units.y) }
        var planes:[jpeg_sm100_plane_i16] = self.indices.map
        {
            .init(coef: nil, units_x: .init(self[$0].units.x), units_y: .init(self[$0].units.y))
        }
        try JPEG.SM100.withPointers(&buffers, &planes)
        {
            try JPEG.SM100.check(jpeg_sm100_spectral_download(JPEG.SM100.shared.ctx, device.handle, $0, .init(planes.count)))
        }
        for p:Int in self.indices
        {
            self[p].set(values: buffers[p], units: self[p].units)
        }
        device.hostIsCurrent = true
    }
    /// Spectral.idct() (decode.swift:4154) on the resident image: one call, the sample planes come back, the coefficients stay
    func sm100IDCTResident(_ device:JPEG.SM100.Resident, quanta:[UInt16], precision:Int,
        planes:inout [jpeg_sm100_plane_u16]) throws
    {
        try JPEG.SM100.check(jpeg_sm100_spectral_idct(JPEG.SM100.shared.ctx, device.handle, quanta, .init(precision),
            &planes, .init(planes.count)))
    }
    /// idct().interleaved(cosite:).unpack(as: RGB.self) on the resident image: one download of 3 bytes per pixel
    func sm100RGBResident(_ device:JPEG.SM100.Resident, quanta:[UInt16], factors:[Int32], cosite cosited:Bool) throws -> [JPEG.RGB]
    {
        let pixels:Int = self.size.x * self.size.y
        return try .init(unsafeUninitializedCapacity: pixels)
        {
            (rgb:inout UnsafeMutableBufferPointer<JPEG.RGB>, initialized:inout Int) in
            try rgb.withMemoryRebound(to: UInt8.self)
            {
                try JPEG.SM100.check(jpeg_sm100_spectral_rgb8(JPEG.SM100.shared.ctx, device.handle, quanta, factors,
                    .init(self.size.x), .init(self.size.y), cosited ? 1 : 0, $0.baseAddress))
            }
            initialized = pixels
        }
    }
    /// Spectral.encode(scan:) (encode.swift:1559) from the resident image (after sm100Upload if the Swift planes were edited)
    func sm100EncodeResident(_ device:JPEG.SM100.Resident, descriptor:inout jpeg_sm100_scan_desc,
        dc:inout [jpeg_sm100_huff_table], ac:inout [jpeg_sm100_huff_table], ecs:inout [UInt8]) throws -> Int
    {
        var length:UInt64 = 0
        try JPEG.SM100.check(jpeg_sm100_spectral_encode_scan(JPEG.SM100.shared.ctx, device.handle, &descriptor, 0,
            &dc, &ac, &ecs, .init(ecs.count), &length))
        return .init(length)
    }
    /// Swift planes -> device, after the host edited coefficients (subscript setters mark the resident copy stale)
    mutating
    func sm100Upload(_ device:JPEG.SM100.Resident) throws
    {
        var buffers:[[Int16]] = self.indices.map{ self[$0].takeBuffer() }
        defer
        {
            for p:Int in self.indices
            {
                self[p].set(values: buffers[p], units: self[p].units)
            }
        }
        var planes:[jpeg_sm100_plane_i16] = self.indices.map
        {
            .init(coef: nil, units_x: .init(self[$0].units.x), units_y: .init(self[$0].units.y))
        }
        try JPEG.SM100.withPointers(&buffers, &planes)
        {
            try JPEG.SM100.check(jpeg_sm100_spectral_upload(JPEG.SM100.shared.ctx, device.handle, $0, .init(planes.count)))
        }
    }
}

// MARK: spectral-domain operations (N3)
//
// The reference has no library function for these: examples/recompress/main.swift:35-58 and examples/rotate/main.swift:164-190
// write them as loops over `Plane[x:y:z:]`.  These conveniences run the same arithmetic on the GPU, plane by plane.
extension JPEG.Data.Spectral.Plane
{
    /// examples/recompress/main.swift:46-52: Int16(Double(Int16(old[z]) * c) / Double(new[z]) + 0.3 * sign)
    func sm100Requantized(from old:JPEG.Table.Quantization, to new:JPEG.Table.Quantization) throws -> [Int16]
    {
        let count:Int = 64 * self.units.x * self.units.y
        return try .init(unsafeUninitializedCapacity: count)
        {
            (out:inout UnsafeMutableBufferPointer<Int16>, initialized:inout Int) in
            try self.withUnsafeCoefficients
            {
                try JPEG.SM100.check(jpeg_sm100_requantize(JPEG.SM100.shared.ctx, $0.baseAddress,
                    .init(self.units.x), .init(self.units.y), old.storage, new.storage, out.baseAddress))
            }
            initialized = count
        }
    }
    /// examples/rotate/main.swift:164-190: block s lands at offset + M s, coefficient z = source[mapping[z].z] * multiplier
    func sm100Transformed(matrix:(x:(x:Int, y:Int), y:(x:Int, y:Int)), mapping:[(z:Int, multiplier:Int16)],
        units:(x:Int, y:Int)) throws -> [Int16]
    {
        let count:Int       = 64 * units.x * units.y
        let m:[Int32]       = [.init(matrix.x.x), .init(matrix.x.y), .init(matrix.y.x), .init(matrix.y.y)]
        let zmap:[UInt8]    = mapping.map{ .init($0.z) }
        let mul:[Int8]      = mapping.map{ .init($0.multiplier) }
        return try .init(unsafeUninitializedCapacity: count)
        {
            (out:inout UnsafeMutableBufferPointer<Int16>, initialized:inout Int) in
            try self.withUnsafeCoefficients
            {
                try JPEG.SM100.check(jpeg_sm100_transform_blocks(JPEG.SM100.shared.ctx, $0.baseAddress,
                    .init(self.units.x), .init(self.units.y), m, zmap, mul, out.baseAddress, .init(units.x), .init(units.y)))
            }
            initialized = count
        }
    }
}

// MARK: encode seams (mirror image; same pattern)
//
//   JPEG.RGB.pack(_:as:)                 jpeg.swift:584   -> jpeg_sm100_pack_rgb8
//   Rectangular.decomposed()             encode.swift:389 -> jpeg_sm100_decompose
//   Spectral.Plane.fdct(_:quanta:...)    encode.swift:199 -> jpeg_sm100_fdct
//   Spectral.encode(scan:)               encode.swift:1559 -> jpeg_sm100_encode_scan; the returned
//       jpeg_sm100_huff_table values are turned back into Table.Huffman with init(counts:values:target:)
//       (decode.swift:368) and the bytes are what stream.format(prefix:) writes (encode.swift:1967).

// MARK: helpers the seams above assume (added to the module alongside this file)
//
//   Spectral.Plane.takeBuffer() -> [Int16]                    moves `buffer` out (no copy), leaves it empty
//   Spectral.Plane.withUnsafeCoefficients(_:)                 buffer.withUnsafeBufferPointer
//   Planar.withUnsafePlanes(_:)                               nests buffer.withUnsafeBufferPointer over the planes and
//                                                             builds [jpeg_sm100_plane_u16] (samples, units, factor)
//   JPEG.SM100.withPointers(&buffers, &planes, body)          pins every buffers[p] and stores its base address in
//                                                             planes[p].coef for the duration of `body`

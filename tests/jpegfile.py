"""Minimal JPEG container splitter for the tests (markers + raw entropy-coded bytes; nothing is decoded)."""
from __future__ import annotations

import numpy as np


def split(data: bytes):
    """Returns a list of (marker, body, ecs) where ecs is the raw (still stuffed, incl. RSTn) bytes that follow
    an SOS segment, b'' otherwise."""
    out = []
    i, n = 0, len(data)
    while i < n:
        assert data[i] == 0xFF, (i, data[i])
        while data[i] == 0xFF:
            i += 1
        m = data[i]
        i += 1
        if m in (0xD8, 0xD9):
            out.append((m, b"", b""))
            if m == 0xD9:
                break
            continue
        ln = (data[i] << 8) | data[i + 1]
        body = data[i + 2:i + ln]
        i += ln
        ecs = b""
        if m == 0xDA:
            j = i
            while True:
                k = data.index(b"\xff", j)
                nxt = data[k + 1]
                if nxt == 0x00 or 0xD0 <= nxt <= 0xD7:
                    j = k + 2
                    continue
                if nxt == 0xFF:
                    j = k + 1
                    continue
                break
            ecs = data[i:k]
            i = k
        out.append((m, body, ecs))
    return out


def unstuff_split(ecs: bytes):
    """Raw scan bytes -> list of unstuffed per-interval ECS (what the reference lexer hands to the decoder,
    decode.swift:130-190)."""
    parts, cur = [], bytearray()
    i, n = 0, len(ecs)
    while i < n:
        b = ecs[i]
        if b != 0xFF:
            cur.append(b)
            i += 1
            continue
        nxt = ecs[i + 1]
        if nxt == 0x00:
            cur.append(0xFF)
            i += 2
        elif 0xD0 <= nxt <= 0xD7:
            parts.append(bytes(cur))
            cur = bytearray()
            i += 2
        elif nxt == 0xFF:
            i += 1
        else:
            raise ValueError("marker inside scan")
    parts.append(bytes(cur))
    return parts


def parse_dht(body: bytes):
    """-> list of (class, target, counts(16 bytes), values)."""
    out, i = [], 0
    while i < len(body):
        cls, tgt = body[i] >> 4, body[i] & 15
        counts = body[i + 1:i + 17]
        n = sum(counts)
        out.append((cls, tgt, bytes(counts), bytes(body[i + 17:i + 17 + n])))
        i += 17 + n
    return out


def parse_dqt(body: bytes):
    """-> list of (target, 64 quanta in zig-zag order); 8-bit and 16-bit (Pq = 1, big-endian) tables"""
    out, i = [], 0
    while i < len(body):
        prec, tgt = body[i] >> 4, body[i] & 15
        assert prec in (0, 1)
        if prec == 0:
            out.append((tgt, list(body[i + 1:i + 65])))
            i += 65
        else:
            out.append((tgt, [(body[i + 1 + 2 * k] << 8) | body[i + 2 + 2 * k] for k in range(64)]))
            i += 129
    return out


def with_dnl(data: bytes, frame_height: int, dnl_height: int) -> bytes:
    """The same file with the frame header's height replaced by `frame_height` (0 = "defined by DNL") and a DNL segment
    (T.81 B.2.5; decode.swift:3905-3924) of `dnl_height` lines after the first scan."""
    out, first = bytearray(), True
    for m, body, ecs in split(data):
        if 0xC0 <= m <= 0xCF and m not in (0xC4, 0xC8, 0xCC):
            body = body[:1] + frame_height.to_bytes(2, "big") + body[3:]
        out += bytes([0xFF, m])
        if m not in (0xD8, 0xD9):
            out += (len(body) + 2).to_bytes(2, "big") + body + ecs
        if m == 0xDA and first:
            out += bytes([0xFF, 0xDC, 0, 4]) + dnl_height.to_bytes(2, "big")
            first = False
    return bytes(out)


def oracle_on_virtual_grid(O, src, comps, interval):
    """The oracle extended to ITU-T T.81 interval placement the same way the library is (remap.cu): the scan's MCUs, in order, on
    a grid of gcd(interval, MCUs) columns, where every interval is whole rows.  Returns the virtual Spectral (scan components only)."""
    import math
    inter = len(comps) > 1
    W, Hh = src.blocks if inter else src.units(comps[0])
    fac = [src.factor(c) for c in comps] if inter else [(1, 1)]
    sx, sy = (max(f[0] for f in fac), max(f[1] for f in fac))
    M = W * Hh
    g = math.gcd(interval, M)
    v = O.Spectral.create((g * 8 * sx, (M // g) * 8 * sy), fac, progressive=True)
    m = np.arange(M)
    for i, c in enumerate(comps):
        fx, fy = fac[i]
        real, virt = src.coefficients(c), v.coefficients(i)
        assert virt.shape[:2] == ((M // g) * fy, g * fx)
        for dy in range(fy):
            for dx in range(fx):
                ry, rx = (m // W) * fy + dy, (m % W) * fx + dx
                vy, vx = (m // g) * fy + dy, (m % g) * fx + dx
                ok = (ry < real.shape[0]) & (rx < real.shape[1])
                virt[vy[ok], vx[ok]] = real[ry[ok], rx[ok]]
    return v

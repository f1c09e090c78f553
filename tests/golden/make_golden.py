#!/usr/bin/env python3
"""Generates tests/golden/ from the reference checkout (run in the build container, where /root/reference exists).

Inputs (small JPEG / RGB files) are copied verbatim; the reference's own committed OUTPUTS are recorded as
sha256 digests in manifest.json, so the fixtures stay small while remaining bit-exact acceptance vectors:

  * tests/regression/gold/*.jpg + .ycc/.rgb          (tests/regression/tests.swift:39-138)
  * examples/decode-basic, decode-advanced (+ per-plane IDCT dumps), in-memory (.rgb and re-encoded .jpg.jpg)
  * examples/encode-basic: 32 JPEGs produced from karlie-milan-sp12-2011.rgb   (main.swift:22-57)
  * examples/recompress: original.jpg -> recompressed-requantized.jpg
  * examples/rotate: karlie-kwk-wwdc-2017.jpg -> -ii / -iii / -iv .jpg (lossless rotations in the spectral domain)
  * examples/encode-advanced: 11-scan progressive 4:2:2 (input 1.6 MB, not copied: digest-only, checked when
    /root/reference is present)
  * tests/integration/decode/*-restart.jpg (no reference output exists; inputs only)

Nothing here is produced by our oracle or our kernels.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import jpegfile as J  # noqa: E402

REF = os.environ.get("JPEG_REFERENCE", "/root/reference")


def sha(b):
    return hashlib.sha256(b).hexdigest()


def rd(p):
    with open(os.path.join(REF, p), "rb") as f:
        return f.read()


def copy(p, sub, name=None):
    os.makedirs(os.path.join(HERE, sub), exist_ok=True)
    name = name or os.path.basename(p)
    dst = os.path.join(HERE, sub, name)
    shutil.copyfile(os.path.join(REF, p), dst)
    os.chmod(dst, 0o644)
    return os.path.join(sub, name)


def scans_of(jpeg_bytes):
    """per-scan expectation: SOS body, sha of raw ECS, DHT tables that precede the scan, plus all DQT."""
    out, dht = [], []
    for m, b, e in J.split(jpeg_bytes):
        if m == 0xC4:
            dht += [[c, t, cn.hex(), v.hex()] for c, t, cn, v in J.parse_dht(b)]
        elif m == 0xDA:
            out.append({"sos": b.hex(), "ecs_sha256": sha(e), "ecs_len": len(e), "dht": dht})
            dht = []
    dqt = [[t, q] for m, b, e in J.split(jpeg_bytes) if m == 0xDB for t, q in J.parse_dqt(b)]
    return {"scans": out, "dqt": dqt, "file_sha256": sha(jpeg_bytes)}


def main():
    man = {"reference": "tayloraswift/jpeg @ 8fe8fda1", "decode": [], "encode_basic": {}, "reencode": {}}
    gold = "tests/regression/gold"
    names = sorted(f for f in os.listdir(os.path.join(REF, gold)) if f.endswith(".jpg"))
    for n in names:
        rel = copy(f"{gold}/{n}", "gold")
        man["decode"].append({"jpeg": rel, "rgb_sha256": sha(rd(f"{gold}/{n}.rgb")),
                              "ycc_sha256": sha(rd(f"{gold}/{n}.ycc"))})
    rel = copy("examples/decode-basic/karlie-kwk-2019.jpg", "examples")
    man["decode"].append({"jpeg": rel, "rgb_sha256": sha(rd("examples/decode-basic/karlie-kwk-2019.jpg.rgb"))})
    rel = copy("examples/decode-advanced/karlie-2019.jpg", "examples")
    man["decode"].append({"jpeg": rel, "rgb_sha256": sha(rd("examples/decode-advanced/karlie-2019.jpg.rgb")),
                          "planes_sha256": [sha(rd(f"examples/decode-advanced/karlie-2019.jpg-{s}.gray"))
                                            for s in ("0.640x432", "1.320x216", "2.320x216")]})
    rel = copy("examples/in-memory/karlie-2011.jpg", "examples")
    man["decode"].append({"jpeg": rel, "rgb_sha256": sha(rd("examples/in-memory/karlie-2011.jpg.rgb"))})
    man["reencode"]["in-memory"] = {"source": rel, **scans_of(rd("examples/in-memory/karlie-2011.jpg.jpg"))}
    rel = copy("examples/recompress/original.jpg", "examples")
    man["reencode"]["recompress-requantized"] = {
        "source": rel, **scans_of(rd("examples/recompress/recompressed-requantized.jpg"))}
    # examples/rotate: lossless rotation in the spectral domain (N3); the three committed outputs pin the block mapping,
    # the crop (Spectral.set(width:/height:)) and the quantisation-table permutation
    rel = copy("examples/rotate/karlie-kwk-wwdc-2017.jpg", "examples")
    man["rotate"] = {"source": rel, "outputs": {k: scans_of(rd(f"examples/rotate/karlie-kwk-wwdc-2017-{k}.jpg"))
                                                 for k in ("ii", "iii", "iv")}}
    man["restart"] = [copy(f"tests/integration/decode/{n}", "restart")
                      for n in sorted(os.listdir(os.path.join(REF, "tests/integration/decode")))
                      if n.endswith("-restart.jpg")]
    rel = copy("examples/encode-basic/karlie-milan-sp12-2011.rgb", "examples")
    man["encode_basic"] = {"rgb": rel, "size": [400, 665], "files": {}}
    for name in ("4-4-4", "4-4-0", "4-2-2", "4-2-0"):
        for level in ("0.0", "0.125", "0.25", "0.5", "1.0", "2.0", "4.0", "8.0"):
            man["encode_basic"]["files"][f"{name}-{level}"] = scans_of(
                rd(f"examples/encode-basic/karlie-milan-sp12-2011-{name}-{level}.jpg"))
    man["encode_advanced"] = {"rgb_sha256": sha(rd("examples/encode-advanced/karlie-cfdas-2011.png.rgb")),
                              "size": [600, 900],
                              **scans_of(rd("examples/encode-advanced/karlie-cfdas-2011.png.rgb.jpg"))}
    # examples/custom-color: a user-defined format (components 4-7, 12-bit samples, 16-bit DQT), 10-scan progression with
    # two-component DC scans of unequal sampling; the .rgb is the INPUT of the encode (main.swift:206), 12 bits in 2 bytes
    rel = copy("examples/custom-color/output.jpg", "examples", "custom-color-output.jpg")
    man["custom_color"] = {"jpeg": rel, "size": [1000, 200], "format": [[4, 5, 6, 7], 12],
                           "factors": [[2, 2], [2, 2], [2, 2], [1, 1]],
                           "rgb_sha256": sha(rd("examples/custom-color/output.jpg.rgb")),
                           **scans_of(rd("examples/custom-color/output.jpg"))}
    # examples/decode-online: JPEG.Context driven segment by segment; the example dumps idct().interleaved().unpack(as: RGB)
    # after EVERY scan of a 10-scan progressive file with DRI -- ten per-scan vectors (and the only restart-interval file
    # of the reference with committed output)
    rel = copy("examples/decode-online/karlie-oscars-2017.jpg", "examples")
    man["decode_online"] = {"jpeg": rel, "rgb_sha256": [sha(rd(f"examples/decode-online/karlie-oscars-2017.jpg-{k}.rgb"))
                                                        for k in range(10)]}
    # tests/unit/tests.swift:170-340: 162 (length, codeword) pairs of the T.81 K.3.3.2 AC-luminance table, listed in
    # the order the unit test walks the symbols (run/size 0x00, 0x01..0x0A, 0x11.., 0xF0, ...)
    import re
    src = rd("tests/unit/tests.swift").decode()
    body = src[src.index("for (length, codeword):(Int, UInt16) in"):]
    body = body[:body.index("]\n")]
    man["unit"] = {"annex_k_ac_codewords": [[int(a), int(b, 2)] for a, b in
                                             re.findall(r"\((\d+),\s*0b([01]+)\)", body)]}
    assert len(man["unit"]["annex_k_ac_codewords"]) == 162
    with open(os.path.join(HERE, "manifest.json"), "w") as f:
        json.dump(man, f, indent=1)
    print("wrote manifest with", len(man["decode"]), "decode vectors,", len(man["encode_basic"]["files"]),
          "encode-basic files")


if __name__ == "__main__":
    main()

"""Hand-worked known-answer vectors for the two `decode_data` codecs (tests/golden/codec_kat.json).

No reference fixture exists for them (LegendDataTypes.jl is not vendored under /root/reference and its tests need LH5
files), so the vectors are worked out BY HAND from the published algorithms and written down here with their derivation;
the script only serialises them.  They pin the byte layout both the oracle and the product encoders must produce.

radware-sigcompress v1.0 (`compress_signal`, D. Radford), 16-bit words big-endian, shift = 0 unless noted:
  A  signal 0,1,2,3            first 48 samples: values span 3, differences span 0 -> difference section
     words: [n=4] [nw=4] [nb+32 = 2+32 = 34] [start 0] [min_diff 1] [3 x (1-1)=0 in 2 bits -> 0x0000]        6 words
  B  signal 5,5,5,9            values span 4, differences (0,0,4) span 4: 4 <= 4 -> absolute section, 3 bits (4 > 3)
     words: [4] [4] [3] [min 5] [000 000 000 100 + 0000 -> 0x0040] + 1 padding word (even word count)        6 words
  C  signal 7 (one sample)     one-sample section: difference range is -32000 (initial +-16000), so the difference
     branch is taken with 2 bits and no payload: [1] [1] [34] [7] [min 16000 = 0x3e80] + padding              6 words
  D  B with UInt16 samples 32773,32773,32773,32777 and shift -32768: same stream as B
ULEB128 zig-zag difference codec:
  E  0,1,300 (uint32)          differences 0,1,299 -> zig-zag 0,2,598 -> 00 | 02 | d6 04
  F  5,3 (uint16)              differences 5,-2 -> zig-zag 10,3 -> 0a 03
  G  70000,0 (uint32)          differences 70000,-70000 -> zig-zag 140000,139999 -> e0 c5 08 | df c5 08
"""
import json
import os


def words(*w):
    out = []
    for v in w:
        out += [(v >> 8) & 0xff, v & 0xff]
    return out


KAT = {
    "radware": [
        {"name": "A", "shift": 0, "signal": [0, 1, 2, 3], "bytes": words(4, 4, 34, 0, 1, 0)},
        {"name": "B", "shift": 0, "signal": [5, 5, 5, 9], "bytes": words(4, 4, 3, 5, 0x0040, 0)},
        {"name": "C", "shift": 0, "signal": [7], "bytes": words(1, 1, 34, 7, 16000, 0)},
        {"name": "D", "shift": -32768, "signal": [32773, 32773, 32773, 32777], "bytes": words(4, 4, 3, 5, 0x0040, 0)},
    ],
    "uleb128zzd": [
        {"name": "E", "dtype": "uint32", "signal": [0, 1, 300], "bytes": [0x00, 0x02, 0xd6, 0x04]},
        {"name": "F", "dtype": "uint16", "signal": [5, 3], "bytes": [0x0a, 0x03]},
        {"name": "G", "dtype": "uint32", "signal": [70000, 0], "bytes": [0xe0, 0xc5, 0x08, 0xdf, 0xc5, 0x08]},
    ],
}

if __name__ == "__main__":
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "codec_kat.json"), "w") as f:
        json.dump(KAT, f, indent=1)
    print("wrote codec_kat.json")

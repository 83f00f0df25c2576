"""Generates the committed golden fixtures.  Run once in the build container:

    python tests/golden/make_golden.py

* xxh3_kat.json — XXH3-128 known answers from python-xxhash 3.7.0 (bundles xxHash 0.8.2, the
  published implementation of the algorithm xxhash-rust 0.8.15 restates; output frozen since
  0.8.0), stored the way hash_chunk stores them (reference src/util/chunk.rs:46-49:
  u128::to_le_bytes() == byte-reversed canonical digest).  Inputs: the literals the reference's
  own tests hash (src/util/tests.rs:90,98-99,109,120,149; tests/roundtrip.rs:9) plus a pattern
  input at every length class boundary.
* dummy_archive.hex — the reference's hand-built archive fixture (src/archive/tests.rs:14-58),
  byte for byte, with a fixed timestamp.  Its payload is what zstd::encode_all(b"test", 0) emits:
  a streaming frame WITHOUT content size (FHD 0x00, window descriptor 0x58, one raw last block).
"""
import json
import struct
from pathlib import Path

import xxhash

HERE = Path(__file__).resolve().parent


def stored(b: bytes) -> str:
    return xxhash.xxh3_128_digest(b)[::-1].hex()


def pattern(n: int) -> bytes:
    return bytes(((i * 2654435761) & 0xFFFFFFFF) >> 24 for i in range(n))


def main():
    literals = {
        "some test data": b"some test data", "data 1": b"data 1", "data 2": b"data 2", "hello squish": b"hello squish",
        "[1u8;1024]": bytes([1]) * 1024, "[2u8;1024]": bytes([2]) * 1024, "[3u8;1024]": bytes([3]) * 1024,
        "[42u8;2048]": bytes([42]) * 2048, "2MiB zeros": bytes(2 << 20), "Hello, world!\\n": b"Hello, world!\n",
    }
    kat = {"literals": [{"name": k, "hex_input": v.hex() if len(v) <= 64 else None, "len": len(v),
                          "fill": v[0] if len(v) > 64 else None, "digest": stored(v)} for k, v in literals.items()],
           "pattern": [{"len": n, "digest": stored(pattern(n))} for n in
                       [1, 2, 3, 4, 5, 8, 9, 15, 16, 17, 31, 32, 33, 64, 96, 97, 128, 129, 160, 161, 192, 239, 240, 241, 255, 256, 304,
                        511, 512, 1023, 1024, 1025, 1087, 1088, 1089, 2047, 2048, 2049, 4096, 65535, 65536, 131072, 1000003,
                        2097151, 2097152]]}
    (HERE / "xxh3_kat.json").write_text(json.dumps(kat, indent=1))

    frame = bytes.fromhex("28b52ffd") + bytes([0x00, 0x58]) + bytes([0x21, 0x00, 0x00]) + b"test"
    a = b"squish" + b"1.2.0" + struct.pack("<Q", 1760000000) + struct.pack("<Q", 1)
    a += bytes([1]) * 16 + struct.pack("<Q", 4) + struct.pack("<Q", len(frame)) + frame
    a += struct.pack("<I", 1) + struct.pack("<I", 9) + b"file1.txt" + struct.pack("<Q", 4) + struct.pack("<I", 1) + bytes([1]) * 16
    (HERE / "dummy_archive.hex").write_text(a.hex() + "\n")
    print("wrote", HERE / "xxh3_kat.json", HERE / "dummy_archive.hex")


if __name__ == "__main__":
    main()

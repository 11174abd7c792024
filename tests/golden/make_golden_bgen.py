"""Regenerate tests/golden/bgen_example.npz from the reference's own BGEN / GEN test resources (authoring container only):

    python tests/golden/make_golden_bgen.py

  bgen        the first 64 variants of hail/hail/test/resources/example.8bits.bgen (v1.2 layout 2, zlib, 8 bits, 500 samples,
              sample identifiers in the file), re-headed to say 64 variants
  gen_dosage  float32 [64, 500]: GP[1] + 2 GP[2] of the same variants from example.gen (the text form of the same data at
              full precision), NaN where the triple is 0 0 0 -- the pair the reference compares in
              test_impex.py:1264-1272 (`bgenmt._same(genmt, tolerance=1/255, absolute=True)`)
  samples     the ids of example.sample;  varid / rsid / position of the 64 variants from example.gen
"""
import os
import struct

import numpy as np

RES = "/root/reference/hail/hail/test/resources"
OUT = os.path.dirname(os.path.abspath(__file__))
KEEP = 64


def main():
    b = open(f"{RES}/example.8bits.bgen", "rb").read()
    offset, = struct.unpack_from("<I", b, 0)
    header_len, M, N = struct.unpack_from("<III", b, 4)
    p = 4 + offset
    for _ in range(KEEP):   # walk KEEP variant blocks
        for _ in range(3):
            ln, = struct.unpack_from("<H", b, p)
            p += 2 + ln
        _, k = struct.unpack_from("<IH", b, p)
        p += 6
        for _ in range(k):
            ln, = struct.unpack_from("<I", b, p)
            p += 4 + ln
        size, = struct.unpack_from("<I", b, p)
        p += 4 + size
    cut = bytearray(b[:p])
    struct.pack_into("<I", cut, 8, KEEP)
    dos, varid, rsid, pos = [], [], [], []
    with open(f"{RES}/example.gen") as f:
        for i, line in enumerate(f):
            if i == KEEP:
                break
            t = line.split()
            g = np.array(t[6:], dtype=np.float64).reshape(-1, 3)
            d = g[:, 1] + 2.0 * g[:, 2]
            d[g.sum(axis=1) == 0.0] = np.nan
            dos.append(d)
            varid.append(t[1])
            rsid.append(t[2])
            pos.append(int(t[3]))
    with open(f"{RES}/example.sample") as f:
        samples = [ln.split()[0] for ln in f.read().splitlines()[2:] if ln.strip()]
    assert len(samples) == N == 500
    np.savez_compressed(os.path.join(OUT, "bgen_example.npz"), bgen=np.frombuffer(bytes(cut), dtype=np.uint8),
                        gen_dosage=np.array(dos, dtype=np.float32), samples=np.array(samples), varid=np.array(varid),
                        rsid=np.array(rsid), position=np.array(pos, dtype=np.int64))
    print("bgen bytes", len(cut), "->", os.path.getsize(os.path.join(OUT, "bgen_example.npz")), "bytes")


if __name__ == "__main__":
    main()

"""BGEN ingest on the host (CPU): `read_bgen` against the reference's own pair of test resources (example.8bits.bgen and
example.gen, tests/golden/bgen_example.npz; hail/python/test/hail/methods/test_impex.py:1223-1294), and the fatal conditions
of the reference's decoder (hail/hail/src/is/hail/io/bgen/StagedBGENReader.scala:205-480) on hand-built files."""
import os
import struct
import zlib

import numpy as np
import pytest

from hail_b200.impex import read_bgen
from hail_b200.statgen import FatalError

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "bgen_example.npz")


def write_bgen(variants, n_samples, *, compression=1, layout=2, sample_ids=None, bits=8, phased=0, ploidy=2, n_alleles=2,
               min_ploidy=2, max_ploidy=2):
    """A minimal BGEN v1.2 writer: `variants` = [(varid, rsid, chrom, pos, d0 uint8 [N], d1 uint8 [N], missing bool [N])]."""
    blocks = b""
    for varid, rsid, chrom, pos, d0, d1, miss in variants:
        blk = b""
        for sfield in (varid, rsid, chrom):
            blk += struct.pack("<H", len(sfield)) + sfield.encode()
        blk += struct.pack("<IH", pos, n_alleles)
        for a in ["A", "G", "T"][:n_alleles]:
            blk += struct.pack("<I", len(a)) + a.encode()
        pl = np.full(n_samples, ploidy, dtype=np.uint8) | (np.asarray(miss, dtype=np.uint8) << 7)
        pr = np.stack([d0, d1], axis=1).astype(np.uint8).tobytes()
        if bits != 8:
            pr = pr * (bits // 8)
        data = struct.pack("<IHBB", n_samples, n_alleles, min_ploidy, max_ploidy) + pl.tobytes() + bytes([phased, bits]) + pr
        if compression == 1:
            z = zlib.compress(data)
            blk += struct.pack("<II", len(z) + 4, len(data)) + z
        else:
            blk += struct.pack("<I", len(data)) + data
        blocks += blk
    ids = b""
    if sample_ids is not None:
        body = b"".join(struct.pack("<H", len(s)) + s.encode() for s in sample_ids)
        ids = struct.pack("<II", len(body) + 8, len(sample_ids)) + body
    flags = compression | (layout << 2) | ((1 << 31) if sample_ids is not None else 0)
    header = struct.pack("<III", 20, len(variants), n_samples) + b"bgen" + struct.pack("<I", flags)
    return struct.pack("<I", len(header) + len(ids)) + header + ids + blocks


def test_example_bgen_equals_example_gen():   # test_impex.py:1264-1272, :1223-1227
    z = np.load(GOLDEN)
    d = read_bgen(z["bgen"].tobytes(), want_probabilities=True)
    assert d["q"].shape == (64, 500) and d["samples"] == list(z["samples"])          # the file carries its sample ids
    assert d["varid"] == list(z["varid"]) and d["rsid"] == list(z["rsid"]) and np.array_equal(d["position"], z["position"])
    assert set(d["contig"]) == {"01"} and d["alleles"][0] == ("A", "G")
    miss = d["q"] == 0xFFFF
    assert np.array_equal(miss, np.isnan(z["gen_dosage"]))                           # the same entries are missing
    dosage = np.where(miss, np.nan, d["q"] / 255.0)
    # each stored probability is within 1/255 of the text file's (the reference's tolerance): the dosage within 3/255
    assert np.nanmax(np.abs(dosage - z["gen_dosage"])) <= 3.0 / 255 + 1e-6
    # `dosage` and gp_dosage(GP) agree (test_impex.py:1285-1294): (d1 + 2 d2) / 255 with d0 + d1 + d2 = 255
    d2 = 255 - d["d0"].astype(int) - d["d1"].astype(int)
    assert (d2[~miss] >= 0).all()
    assert np.array_equal(d["q"][~miss], (d["d1"].astype(int) + 2 * d2)[~miss])


def test_round_trip_and_sample_sources(tmp_path):
    rng = np.random.default_rng(5)
    N, M = 37, 9
    vs = []
    for v in range(M):
        d0 = rng.integers(0, 256, N)
        d1 = np.array([rng.integers(0, 256 - a) for a in d0])
        vs.append((f"v{v}", f"rs{v}", "20", 100 + v, d0, d1, rng.random(N) < 0.1))
    for comp, ids in ((1, [f"s{i}" for i in range(N)]), (0, None)):
        raw = write_bgen(vs, N, compression=comp, sample_ids=ids)
        path = tmp_path / f"t{comp}.bgen"
        path.write_bytes(raw)
        d = read_bgen(str(path), want_probabilities=True)
        assert d["samples"] == (ids if ids else [f"sample_{i}" for i in range(N)])
        for v, (_, _, _, pos, d0, d1, miss) in enumerate(vs):
            want = np.where(miss, 0xFFFF, d1 + 2 * (255 - d0 - d1))
            assert np.array_equal(d["q"][v], want) and d["position"][v] == pos
            assert np.array_equal(d["d0"][v], d0) and np.array_equal(d["d1"][v], d1)
    sample = tmp_path / "t.sample"
    sample.write_text("ID_1 ID_2\n0 0\n" + "".join(f"id{i} x\n" for i in range(N)))
    assert read_bgen(raw, str(sample))["samples"] == [f"id{i}" for i in range(N)]
    short = tmp_path / "short.sample"
    short.write_text("ID_1\n0\nonly_one\n")
    with pytest.raises(FatalError, match="different numbers of samples"):
        read_bgen(raw, str(short))


@pytest.mark.parametrize("kw,msg", [
    (dict(bits=16), "Hail only supports 8-bit probabilities, found 16"),
    (dict(phased=1), "Hail does not support phased genotypes"),
    (dict(phased=3), "Phase value must be 0 or 1. Found 3"),
    (dict(ploidy=1), "Ploidy value must equal to 2. Found 1"),
    (dict(min_ploidy=1), "Hail only supports diploid genotypes. Found min ploidy '1' and max ploidy '2'"),
    (dict(n_alleles=3), "Only biallelic variants supported, found variant with 3 alleles: 20:7"),
    (dict(layout=1), "layout 2"),
])
def test_fatal_conditions_of_the_decoder(kw, msg):   # StagedBGENReader.scala:205-215, 424-480
    N = 4
    v = [("v", "rs", "20", 7, np.array([255, 0, 0, 10]), np.array([0, 255, 0, 20]), np.zeros(N, dtype=bool))]
    with pytest.raises(FatalError, match=msg):
        read_bgen(write_bgen(v, N, **kw))


def test_threaded_decode_of_a_large_file():
    """Files of more than 4M entries are decoded by a thread pool; the result is the serial one, and a fatal condition in any
    block still surfaces."""
    rng = np.random.default_rng(2)
    N, M = 70_000, 64
    vs = []
    for v in range(M):
        d0 = rng.integers(0, 256, N)
        d1 = rng.integers(0, 256, N) * (255 - d0) // 255
        vs.append((f"v{v}", f"rs{v}", "20", 100 + v, d0, d1, rng.random(N) < 0.01))
    d = read_bgen(write_bgen(vs, N))
    for v in range(M):
        assert np.array_equal(d["q"][v], np.where(vs[v][6], 0xFFFF, vs[v][5] + 2 * (255 - vs[v][4] - vs[v][5])))
    bad = write_bgen(vs[:63], N)
    one = write_bgen(vs[63:], N, phased=1)
    # splice a phased block behind 63 good ones: same header, one more variant
    off1, = struct.unpack_from("<I", one, 0)
    spliced = bytearray(bad + one[4 + off1:])
    struct.pack_into("<I", spliced, 8, 64)
    with pytest.raises(FatalError, match="phased genotypes"):
        read_bgen(bytes(spliced))

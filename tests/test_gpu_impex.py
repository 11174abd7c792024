"""GPU tests of PLINK import / export end to end.

Mirrors hail/python/test/hail/methods/test_impex.py:836-858 (export -> import gives the same dataset), :874-900
(a2_reference=False swaps alleles and homozygote counts, keeps hets and missing), :948-951 (no reference genome),
:1128-1160 (white space in ids raises).
"""
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import bed as obed
from oracle import linreg_oracle as O
from tests.helpers import assert_fields_close


def _dataset(hb, N=37, M=300, seed=4):
    rng = np.random.default_rng(seed)
    x = rng.integers(-1, 3, size=(M, N)).astype(np.int8)
    contigs = rng.choice(["1", "2", "10", "X", "MT"], size=M)
    pos = rng.permutation(10_000)[:M] + 1          # unique loci: (locus, alleles) is a key
    loc = np.array([(c, int(p)) for c, p in zip(contigs, pos)] + [None], dtype=object)[:-1]
    al = np.array([("A", "C")] * M + [None], dtype=object)[:-1]
    mt = hb.MatrixTable(hb.PackedGenotypes.from_dosage(x), rows={"locus": loc, "alleles": al},
                        cols={"s": np.array([f"s{i}" for i in range(N)], dtype=object)},
                        row_key=("locus", "alleles"), col_key=("s",))
    return mt, x, contigs, pos


def _sorted_order(contigs, pos):
    rank = {c: i for i, c in enumerate([str(i) for i in range(1, 23)] + ["X", "Y", "MT"])}
    return np.array(sorted(range(len(pos)), key=lambda i: (rank[contigs[i]], int(pos[i]), ("A", "C"))))


def test_export_import_plink_same(tmp_path):
    import hail_b200 as hb
    mt, x, contigs, pos = _dataset(hb)
    prefix = str(tmp_path / "rt")
    hb.export_plink(mt, prefix, ind_id=mt.s, cm_position=np.full(x.shape[0], 15.0))
    order = _sorted_order(contigs, pos)
    for resident in (True, False):
        back = hb.import_plink(prefix + ".bed", prefix + ".bim", prefix + ".fam", a2_reference=True,
                               reference_genome="GRCh37", n_partitions=8, resident=resident)
        assert back.count_rows() == x.shape[0] and back.count_cols() == x.shape[1]
        assert np.array_equal(back.genotypes.to_dosage(), x[order])          # rows sorted by (locus, alleles)
        assert [tuple(l) for l in back._rows["locus"]] == [(contigs[i], int(pos[i])) for i in order]
        assert all(tuple(a) == ("A", "C") for a in back._rows["alleles"])
        assert (back._rows["cm_position"] == 15.0).all()
        assert list(back._cols["s"]) == [f"s{i}" for i in range(x.shape[1])]
        assert np.isnan(back._cols["is_female"]).all() and np.isnan(back._cols["is_case"]).all()
        assert all(v is None for v in back._cols["fam_id"])
    # default variant id and the bytes ExportPlink would write (MatrixWriter.scala:2236-2285)
    first = open(prefix + ".bim").readline().rstrip("\n").split("\t")
    assert first == [contigs[0], f"{contigs[0]}:{pos[0]}:A:C", "15.0", str(pos[0]), "C", "A"]
    body = obed.bed_body(np.fromfile(prefix + ".bed", dtype=np.uint8), x.shape[1], x.shape[0])
    want = np.where(x < 0, np.nan, x).astype(np.float64)
    assert np.array_equal(obed.decode_rows(body, x.shape[1]), want, equal_nan=True)


def test_import_plink_a1_major_and_no_reference(tmp_path):
    import hail_b200 as hb
    mt, x, contigs, pos = _dataset(hb, seed=8)
    prefix = str(tmp_path / "a1")
    hb.export_plink(mt, prefix, ind_id=mt.s)
    a2 = hb.import_plink(prefix + ".bed", prefix + ".bim", prefix + ".fam", a2_reference=True)
    a1 = hb.import_plink(prefix + ".bed", prefix + ".bim", prefix + ".fam", a2_reference=False)
    d2, d1 = a2.genotypes.to_dosage(), a1.genotypes.to_dosage()
    key2 = {(tuple(l), tuple(a)): i for i, (l, a) in enumerate(zip(a2._rows["locus"], a2._rows["alleles"]))}
    for j, (l, a) in enumerate(zip(a1._rows["locus"], a1._rows["alleles"])):
        i = key2[(tuple(l), (a[1], a[0]))]                                  # alleles swapped
        assert np.array_equal(d1[j] < 0, d2[i] < 0)                          # n_not_called equal
        assert np.array_equal(d1[j] == 1, d2[i] == 1)                        # hets equal
        assert np.array_equal(d1[j] == 0, d2[i] == 2) and np.array_equal(d1[j] == 2, d2[i] == 0)
    # reference_genome=None: locus is (contig string, position), ordered by the string (test_impex.py:948-951)
    nr = hb.import_plink(prefix + ".bed", prefix + ".bim", prefix + ".fam", reference_genome=None)
    want = sorted((contigs[i], int(pos[i])) for i in range(len(pos)))
    assert [tuple(l) for l in nr._rows["locus"]] == want


def test_export_plink_fam_fields_and_whitespace(tmp_path):
    import hail_b200 as hb
    mt, x, _, _ = _dataset(hb, N=6, M=10)
    prefix = str(tmp_path / "f")
    hb.export_plink(mt, prefix, ind_id=mt.s, fam_id=np.array(["f1", None, "f3", "f4", "f5", "f6"], dtype=object),
                    is_female=np.array([True, False, True, False, True, False]),
                    pheno=np.array([1.5, np.nan, 2.0, -1.0, 0.0, 3.25]))
    lines = [l.rstrip("\n").split("\t") for l in open(prefix + ".fam")]
    assert lines[0] == ["f1", "s0", "0", "0", "2", "1.5"] and lines[1] == ["0", "s1", "0", "0", "1", "NA"]
    back = hb.import_plink(prefix + ".bed", prefix + ".bim", prefix + ".fam", quant_pheno=True)
    assert np.array_equal(back._cols["quant_pheno"], [1.5, np.nan, 2.0, -1.0, 0.0, 3.25], equal_nan=True)
    assert np.array_equal(back._cols["is_female"], [1, 0, 1, 0, 1, 0])
    with pytest.raises(TypeError, match="has spaces in the following values"):     # test_impex.py:1128-1131
        hb.export_plink(mt, prefix, ind_id=mt.s, fam_id=np.array(["a b"] * 6, dtype=object))
    with pytest.raises(TypeError, match="has spaces in the following values"):     # test_impex.py:1133-1136
        hb.export_plink(mt, prefix, ind_id=mt.s, varid=np.array(["v 1"] * 10, dtype=object))


def test_linreg_on_imported_plink_resident_and_streamed(tmp_path):
    """The whole ingest path: files -> import_plink -> linear_regression_rows, resident and streamed, vs the oracle."""
    import hail_b200 as hb
    mt, x, contigs, pos = _dataset(hb, N=500, M=700, seed=12)
    prefix = str(tmp_path / "lr")
    rng = np.random.default_rng(1)
    N = x.shape[1]
    y = rng.normal(size=N)
    y[::29] = np.nan
    cov = rng.normal(size=N)
    hb.export_plink(mt, prefix, ind_id=mt.s, pheno=y)
    order = _sorted_order(contigs, pos)
    xf = np.where(x < 0, np.nan, x).astype(np.float64)[order]
    want = O.linreg_group(xf, y[:, None], np.column_stack([np.ones(N), cov]))
    res = []
    for resident in (True, False):
        pl = hb.import_plink(prefix + ".bed", prefix + ".bim", prefix + ".fam", quant_pheno=True, resident=resident)
        pl = pl.annotate_cols(c=cov)
        ht = hb.linear_regression_rows(y=pl.quant_pheno, x=pl.GT.n_alt_alleles(), covariates=[1.0, pl.c],
                                       pass_through=["rsid"])
        got = {"n": ht.n, "sum_x": np.asarray(ht.sum_x)}
        for f in ("y_transpose_x", "beta", "standard_error", "t_stat", "p_value"):
            got[f] = np.asarray(ht[f])[:, None]
        nondeg = np.isfinite(want["standard_error"]).all(axis=1)
        assert_fields_close({k: v[nondeg] for k, v in got.items()}, {k: v[nondeg] for k, v in want.items() if k != "_d"},
                            t_floor=1e-9, ctx=f"resident={resident}")
        assert list(ht.rsid) == [f"{contigs[i]}:{pos[i]}:A:C" for i in order]
        res.append(ht)
    assert np.array_equal(res[0].beta, res[1].beta, equal_nan=True) and np.array_equal(res[0].p_value, res[1].p_value, equal_nan=True)

"""CPU tests of the PLINK text parsing and the import checks (no GPU): the reference's rules and messages.

Mirrors hail/python/test/hail/methods/test_impex.py:826-872 (import_fam, empty fam / bim) and the parse rules of
hail/hail/src/is/hail/io/plink/LoadPlink.scala:40-82, 102-186, 225-251.  The .fam contents below are the reference's
fixtures `importFamCaseControl.fam`, `importFamQPheno.fam`, `importFamQPheno.space.m9.fam` and
`importFamCaseControlNumericException.fam` (5 lines each).
"""
import numpy as np
import pytest

from hail_b200 import FatalError
from hail_b200 import impex

CASE_CONTROL = "Newton\tA\tC\tD\t1\t1\nTuring\tB\tC\tD\t2\t2\n0\tC\t0\t0\t0\t0\n0\tD\t0\t0\t0\t-9\n0\tE\t0\t0\t0\tnon-numeric\n"
QPHENO = "Newton\tA\tC\tD\t1\t1.0\nTuring\tB\tC\tD\t2\t2.0\n0\tC\t0\t0\t0\t0\n0\tD\t0\t0\t0\t-9\n0\tE\t0\t0\t0\tNA\n"
QPHENO_SPACE_M9 = "Newton A C D 1 1.0\nTuring B C D 2 2.0\n0 C 0 0 0 0\n0 D 0 0 0 -9\n0 E 0 0 0 3.0\n"
NUMERIC_EXC = "Newton\tA\tC\tD\t1\t1\nTuring\tB\tC\tD\t2\t2\n0\tC\t0\t0\t0\t3\n0\tD\t0\t0\t0\t-9\n0\tE\t0\t0\t0\tnon-numeric\n"


def _w(tmp_path, name, text):
    p = tmp_path / name
    p.write_text(text)
    return str(p)


def test_import_fam_case_control(tmp_path):
    f = impex.import_fam(_w(tmp_path, "cc.fam", CASE_CONTROL))
    assert list(f["id"]) == ["A", "B", "C", "D", "E"]
    assert list(f["fam_id"]) == ["Newton", "Turing", None, None, None]
    assert list(f["pat_id"]) == ["C", "C", None, None, None] and list(f["mat_id"]) == ["D", "D", None, None, None]
    assert np.array_equal(f["is_female"], [0.0, 1.0, np.nan, np.nan, np.nan], equal_nan=True)
    assert np.array_equal(f["is_case"], [0.0, 1.0, np.nan, np.nan, np.nan], equal_nan=True)


def test_import_fam_quant_pheno(tmp_path):
    f = impex.import_fam(_w(tmp_path, "q.fam", QPHENO), quant_pheno=True)
    assert np.array_equal(f["quant_pheno"], [1.0, 2.0, 0.0, -9.0, np.nan], equal_nan=True)   # -9 is a value, NA missing
    g = impex.import_fam(_w(tmp_path, "q9.fam", QPHENO_SPACE_M9), quant_pheno=True, delimiter="\\\\s+", missing="-9")
    assert np.array_equal(g["quant_pheno"], [1.0, 2.0, 0.0, np.nan, 3.0], equal_nan=True)


def test_import_fam_errors(tmp_path):
    with pytest.raises(FatalError, match="Invalid case-control phenotype: '3'"):
        impex.import_fam(_w(tmp_path, "e.fam", NUMERIC_EXC))
    with pytest.raises(FatalError, match="Invalid quantitative phenotype: 'non-numeric'"):
        impex.import_fam(_w(tmp_path, "e2.fam", CASE_CONTROL), quant_pheno=True)
    with pytest.raises(FatalError, match="expected 6 fields, but found 5"):
        impex.import_fam(_w(tmp_path, "e3.fam", "a\tb\tc\td\t1\n"))
    with pytest.raises(FatalError, match="Invalid sex: '7'"):
        impex.import_fam(_w(tmp_path, "e4.fam", "a\tb\t0\t0\t7\t1\n"))
    with pytest.raises(FatalError, match="Empty FAM file"):
        impex.import_fam(_w(tmp_path, "e5.fam", ""))


def test_import_plink_checks_before_any_device_work(tmp_path):
    fam = _w(tmp_path, "t.fam", CASE_CONTROL)                       # 5 samples -> 2 bytes per variant
    bim = _w(tmp_path, "t.bim", "1\trs1\t0\t100\tA\tC\n1\trs2\t0\t200\tG\tT\n")
    bed = tmp_path / "t.bed"
    with pytest.raises(FatalError, match="Empty FAM file"):        # test_impex.py:860-865
        impex.import_plink(str(bed), bim, _w(tmp_path, "empty.fam", ""))
    with pytest.raises(FatalError, match="BIM file does not contain any variants"):   # test_impex.py:867-872
        impex.import_plink(str(bed), _w(tmp_path, "empty.bim", ""), fam)
    with pytest.raises(FatalError, match="Invalid .bim line.  Expected 6 fields, found 5 fields"):
        impex.import_plink(str(bed), _w(tmp_path, "bad.bim", "1\trs1\t0\t100\tA\n"), fam)
    bed.write_bytes(bytes([1, 2, 3, 0, 0, 0, 0]))
    with pytest.raises(FatalError, match="do not match PLINK magic numbers 108 & 27"):
        impex.import_plink(str(bed), bim, fam)
    bed.write_bytes(bytes([108, 27, 0, 0, 0, 0, 0]))
    with pytest.raises(FatalError, match="individual major mode"):
        impex.import_plink(str(bed), bim, fam)
    bed.write_bytes(bytes([108, 27, 1, 0, 0, 0]))                   # one byte short
    with pytest.raises(FatalError, match="BED file size does not match"):
        impex.import_plink(str(bed), bim, fam)
    with pytest.raises(FatalError, match="Invalid locus"):          # test_impex.py:968-976
        impex.import_plink(str(bed), _w(tmp_path, "inv.bim", "chr1\trs1\t0\t100\tA\tC\n1\trs2\t0\t200\tG\tT\n"), fam)


def test_swap_table_is_the_a1_major_recoding():
    # LoadPlink.scala:475-481: with a2_reference=False code 0 is hom-ref and 3 hom-alt; 1 (missing) and 2 (het) stay
    for b in range(256):
        out = int(impex._SWAP_HOM[b])
        for f in range(4):
            c, o = (b >> (2 * f)) & 3, (out >> (2 * f)) & 3
            assert o == {0: 3, 3: 0, 1: 1, 2: 2}[c]

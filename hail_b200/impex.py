"""PLINK ingest for the B200 path: .bed/.bim/.fam -> MatrixTable with packed GT in HBM.

Reference behaviour followed: hail/python/hail/methods/impex.py:2505 (`import_plink` signature defaults:
a2_reference=True, missing='NA', quant_pheno=False), hail/hail/src/is/hail/io/plink/LoadPlink.scala:225-251
(header / size checks), :470-530 (entry decode).
"""
from __future__ import annotations

import numpy as np

from .genotypes import PackedGenotypes
from .matrixtable import MatrixTable


def import_plink(bed, bim, fam, device=0, quant_pheno=False, missing="NA") -> MatrixTable:
    contig, rsid, cm, pos, a1, a2 = [], [], [], [], [], []
    with open(bim) as f:
        for line in f:
            r = line.split()
            if not r:
                continue
            contig.append(r[0]); rsid.append(r[1]); cm.append(float(r[2])); pos.append(int(r[3]))
            a1.append(r[4]); a2.append(r[5])
    fam_id, s, pat, mat, sex, pheno = [], [], [], [], [], []
    with open(fam) as f:
        for line in f:
            r = line.split()
            if not r:
                continue
            fam_id.append(r[0]); s.append(r[1]); pat.append(r[2]); mat.append(r[3]); sex.append(r[4]); pheno.append(r[5])
    n_variants, n_samples = len(rsid), len(s)
    gt = PackedGenotypes.from_bed_file(bed, n_samples, n_variants, device)
    is_female = np.array([1.0 if v == "2" else 0.0 if v == "1" else np.nan for v in sex])
    if quant_pheno:
        ph = np.array([np.nan if v in (missing, "-9") else float(v) for v in pheno])
        pheno_field = {"quant_pheno": ph}
    else:
        ph = np.array([1.0 if v == "2" else 0.0 if v == "1" else np.nan for v in pheno])
        pheno_field = {"is_case": ph}
    rows = {
        "locus": np.array(list(zip(contig, pos)), dtype=object),
        # a2_reference=True: A2 is the reference allele (impex.py:2505 docs)
        "alleles": np.array(list(zip(a2, a1)), dtype=object),
        "rsid": np.array(rsid, dtype=object),
        "cm_position": np.array(cm),
    }
    cols = {"s": np.array(s, dtype=object), "fam_id": np.array(fam_id, dtype=object),
            "pat_id": np.array(pat, dtype=object), "mat_id": np.array(mat, dtype=object),
            "is_female": is_female, **pheno_field}
    return MatrixTable(gt, rows=rows, cols=cols, row_key=("locus", "alleles"), col_key=("s",))

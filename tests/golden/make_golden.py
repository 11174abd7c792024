"""Regenerate the committed golden fixtures from the reference's own test resources.

Run in the authoring container only (reads /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

Outputs (committed):
  regression_linear.json   the 8-sample x 10-variant R-lm() golden case
                           (hail/hail/test/resources/regressionLinear.{vcf,gen,sample,pheno,cov,fam};
                           expected values transcribed from
                           hail/python/test/hail/methods/test_statgen.py:223-234, 262-284, 303-316, 334-348, 380-424)
  pt_known_answers.json    hail/python/test/hail/expr/test_expr.py:3564-3568
  fastlmm.npz              PLINK parity data (fastlmmTest.bed/.fam + fastlmmPheno.txt + fastlmmCov.txt)
  bn_4x1024.npz            balding-nichols-1024-variants-4-samples-3-populations.bed (1 % missing)
"""
import json
import os

import numpy as np

RES = "/root/reference/hail/hail/test/resources"
OUT = os.path.dirname(os.path.abspath(__file__))


def read_table(path, missing=None):
    with open(path) as f:
        header = f.readline().split()
        rows = [line.split() for line in f if line.strip()]
    return header, rows


def main():
    # ---- regressionLinear.vcf: GT -> n_alt_alleles, PL -> pl_dosage -------------------
    samples, gt, pl = None, [], []
    with open(f"{RES}/regressionLinear.vcf") as f:
        for line in f:
            if line.startswith("##"):
                continue
            parts = line.rstrip("\n").split("\t")
            if line.startswith("#"):
                samples = parts[9:]
                continue
            fmt = parts[8].split(":")
            g_row, pl_row = [], []
            for cell in parts[9:]:
                fields = dict(zip(fmt, cell.split(":")))
                g = fields["GT"]
                g_row.append(None if "." in g else sum(int(a) for a in g.replace("|", "/").split("/")))
                p = fields.get("PL", ".")
                pl_row.append(None if p == "." else [int(v) for v in p.split(",")])
            gt.append(g_row)
            pl.append(pl_row)

    # ---- regressionLinear.gen: GP triples in .sample order (import_gen, methods/impex.py:1354-1490: the array is
    # missing when |sum - 1| > tolerance = 0.2); gp_dosage = GP[1] + 2 GP[2] (expr/functions.py:1470-1488)
    with open(f"{RES}/regressionLinear.sample") as f:
        gen_samples = [line.split()[0] for line in f.read().splitlines()[2:] if line.strip()]
    gp = []
    with open(f"{RES}/regressionLinear.gen") as f:
        for line in f:
            vals = [float(v) for v in line.split()[6:]]
            row = []
            for i in range(0, len(vals), 3):
                t = vals[i:i + 3]
                row.append(None if abs(sum(t) - 1.0) > 0.2 else t)
            gp.append(row)
    assert gen_samples == samples and len(gp) == len(gt) and all(len(r) == len(samples) for r in gp)

    _, ph_rows = read_table(f"{RES}/regressionLinear.pheno")
    pheno = {r[0]: float(r[1]) for r in ph_rows}
    _, cv_rows = read_table(f"{RES}/regressionLinear.cov")
    cov = {r[0]: [float(r[1]), float(r[2])] for r in cv_rows}
    fam = {}
    with open(f"{RES}/regressionLinear.fam") as f:
        for line in f:
            r = line.split()
            fam[r[1]] = {"is_female_code": r[4], "pheno_code": r[5]}

    doc = {
        "source": "hail/hail/test/resources/regressionLinear.{vcf,pheno,cov,fam}",
        "samples": samples,
        "gt_n_alt_alleles": gt,  # [10 variants][8 samples], None = missing call
        "pl": pl,
        "gp": gp,  # regressionLinear.gen, None = missing (probabilities do not sum to 1 within 0.2)
        "pheno_table": pheno,  # keyed by sample id; '0' means missing under missing='0' (TS:249-251)
        "cov_table": cov,
        "fam_table": fam,
        "expected": {
            # TS:223-234  covariates=[] , pheno with missing='0'
            "no_intercept": {"1": {"beta": 1.5, "standard_error": 1.161895, "t_stat": 1.290994, "p_value": 0.25317}},
            # TS:262-284  covariates=[1, Cov1, Cov2]
            "with_cov": {
                "1": {"beta": -0.28589421, "standard_error": 1.2739153, "t_stat": -0.22442167, "p_value": 0.84327106},
                "2": {"beta": -0.5417647, "standard_error": 0.3350599, "t_stat": -1.616919, "p_value": 0.24728705},
                "3": {"beta": 1.07367185, "standard_error": 0.6764348, "t_stat": 1.5872510, "p_value": 0.2533675},
                "nan_se": [6, 7, 8, 9, 10],
                "nan_t_p": [6],
            },
            # TS:303-316  x = pl_dosage(PL)
            "pl_dosage": {
                "1": {"beta": -0.29166985, "standard_error": 1.2996510, "t_stat": -0.22442167, "p_value": 0.84327106},
                "2": {"beta": -0.5499320, "standard_error": 0.3401110, "t_stat": -1.616919, "p_value": 0.24728705},
                "3": {"beta": 1.09536219, "standard_error": 0.6901002, "t_stat": 1.5872510, "p_value": 0.2533675},
            },
            # TS:334-348  x = gp_dosage(GP) from regressionLinear.gen: the same numbers, places=4 on beta / standard_error
            "gp_dosage": {
                "1": {"beta": -0.29166985, "standard_error": 1.2996510, "t_stat": -0.22442167, "p_value": 0.84327106},
                "2": {"beta": -0.5499320, "standard_error": 0.3401110, "t_stat": -1.616919, "p_value": 0.24728705},
                "3": {"beta": 1.09536219, "standard_error": 0.6901002, "t_stat": 1.5872510, "p_value": 0.2533675},
                "nan_se": [6],
            },
        },
    }
    with open(f"{OUT}/regression_linear.json", "w") as f:
        json.dump(doc, f, indent=1)

    with open(f"{OUT}/pt_known_answers.json", "w") as f:
        json.dump(
            {
                "source": "hail/python/test/hail/expr/test_expr.py:3564-3568",
                "cases": [
                    {"x": 0, "n": 10, "lower_tail": True, "log_p": False, "value": 0.5},
                    {"x": 1, "n": 10, "lower_tail": True, "log_p": False, "value": 0.82955343384897},
                    {"x": 1, "n": 10, "lower_tail": False, "log_p": False, "value": 0.17044656615103004},
                    {"x": 1, "n": 10, "lower_tail": True, "log_p": True, "value": -0.186867754489647},
                ],
            },
            f,
            indent=1,
        )

    # ---- PLINK parity data -----------------------------------------------------------
    def fam_ids(path):
        with open(path) as f:
            return [line.split()[1] for line in f if line.strip()]

    bed = np.fromfile(f"{RES}/fastlmmTest.bed", dtype=np.uint8)
    ids = fam_ids(f"{RES}/fastlmmTest.fam")
    n_var = sum(1 for _ in open(f"{RES}/fastlmmTest.bim"))
    ph = {}
    for line in open(f"{RES}/fastlmmPheno.txt"):
        r = line.split()
        ph[r[1]] = float(r[2])
    cv = {}
    for line in open(f"{RES}/fastlmmCov.txt"):
        r = line.split()
        cv[r[1]] = [float(v) for v in r[2:]]
    np.savez_compressed(
        f"{OUT}/fastlmm.npz",
        bed=bed,
        n_samples=len(ids),
        n_variants=n_var,
        pheno=np.array([ph.get(i, np.nan) for i in ids]),
        cov=np.array([cv.get(i, [np.nan] * len(next(iter(cv.values())))) for i in ids]),
    )

    stem = f"{RES}/balding-nichols-1024-variants-4-samples-3-populations"
    np.savez_compressed(
        f"{OUT}/bn_4x1024.npz",
        bed=np.fromfile(stem + ".bed", dtype=np.uint8),
        n_samples=len(fam_ids(stem + ".fam")),
        n_variants=sum(1 for _ in open(stem + ".bim")),
    )
    print("wrote golden fixtures to", OUT)


if __name__ == "__main__":
    main()

"""Regenerate regression_logistic.json from the reference's test resources (authoring container only).

    python tests/golden/make_golden_logistic.py

Source: hail/hail/test/resources/regressionLogistic.{vcf,cov} + regressionLogisticBoolean.pheno; expected values
transcribed from hail/python/test/hail/methods/test_statgen.py:987-1021 (R: anova(logfitnull, logfit, test="Rao")).
"""
import json
import os

RES = "/root/reference/hail/hail/test/resources"
OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    samples, gt = None, []
    with open(f"{RES}/regressionLogistic.vcf") as f:
        for line in f:
            if line.startswith("##"):
                continue
            parts = line.rstrip("\n").split("\t")
            if line.startswith("#"):
                samples = parts[9:]
                continue
            row = []
            for cell in parts[9:]:
                g = cell.split(":")[0]
                row.append(None if "." in g else sum(int(a) for a in g.replace("|", "/").split("/")))
            gt.append(row)
    cov = {}
    with open(f"{RES}/regressionLogistic.cov") as f:
        f.readline()
        for line in f:
            r = line.split()
            if r:
                cov[r[0]] = [float(r[1]), float(r[2])]
    pheno = {}
    with open(f"{RES}/regressionLogisticBoolean.pheno") as f:
        f.readline()
        for line in f:
            r = line.split()
            if r:
                pheno[r[0]] = None if r[1] == "0" else (r[1] == "true")     # missing='0' (TS:992-994)
    doc = {
        "source": "hail/hail/test/resources/regressionLogistic.{vcf,cov}, regressionLogisticBoolean.pheno",
        "samples": samples,
        "gt_n_alt_alleles": gt,
        "cov_table": cov,
        "pheno_table": pheno,
        "expected_score": {   # test_statgen.py:1007-1021, places=6
            "1": {"chi_sq_stat": 0.1502364955, "p_value": 0.6983094571},
            "2": {"chi_sq_stat": 0.1823600965, "p_value": 0.6693528073},
            "3": {"chi_sq_stat": 7.047367694, "p_value": 0.007938182229},
            "constant": [6, 7, 8, 9, 10],   # chi_sq_stat missing or < 1e-6
        },
    }
    with open(os.path.join(OUT, "regression_logistic.json"), "w") as f:
        json.dump(doc, f, indent=1)
    print("wrote regression_logistic.json:", len(gt), "variants x", len(samples), "samples")


if __name__ == "__main__":
    main()

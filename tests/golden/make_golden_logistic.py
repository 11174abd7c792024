"""Regenerate regression_logistic.json from the reference's test resources (authoring container only).

    python tests/golden/make_golden_logistic.py

Source: hail/hail/test/resources/regressionLogistic.{vcf,cov} + regressionLogisticBoolean.pheno; expected values
transcribed from hail/python/test/hail/methods/test_statgen.py:987-1021 (R: anova(logfitnull, logfit, test="Rao")),
:737-756 (wald), :958-985 (lrt); logistic_epacts.npz from regressionLogisticEpacts.{vcf,cov,fam} with the expected
values of :1722-1862.
"""
import json
import os

RES = "/root/reference/hail/hail/test/resources"
OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    samples, gt, pl = None, [], []
    with open(f"{RES}/regressionLogistic.vcf") as f:
        for line in f:
            if line.startswith("##"):
                continue
            parts = line.rstrip("\n").split("\t")
            if line.startswith("#"):
                samples = parts[9:]
                continue
            row, pl_row = [], []
            fmt = parts[8].split(":")
            for cell in parts[9:]:
                fields = dict(zip(fmt, cell.split(":")))
                g = fields["GT"]
                row.append(None if "." in g else sum(int(a) for a in g.replace("|", "/").split("/")))
                p = fields.get("PL", ".")
                pl_row.append(None if p == "." else [int(v) for v in p.split(",")])
            gt.append(row)
            pl.append(pl_row)
    # regressionLogistic.gen: GP triples in .sample order; missing when |sum - 1| > 0.2 (methods/impex.py import_gen)
    with open(f"{RES}/regressionLogistic.sample") as f:
        gen_samples = [line.split()[0] for line in f.read().splitlines()[2:] if line.strip()]
    gp = []
    with open(f"{RES}/regressionLogistic.gen") as f:
        for line in f:
            vals = [float(v) for v in line.split()[6:]]
            gp.append([None if abs(sum(vals[i:i + 3]) - 1.0) > 0.2 else vals[i:i + 3] for i in range(0, len(vals), 3)])
    assert gen_samples == samples and len(gp) == len(gt), (gen_samples, samples)
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
        "pl": pl,
        "gp": gp,
        "cov_table": cov,
        "pheno_table": pheno,
        "expected_score": {   # test_statgen.py:1007-1021, places=6
            "1": {"chi_sq_stat": 0.1502364955, "p_value": 0.6983094571},
            "2": {"chi_sq_stat": 0.1823600965, "p_value": 0.6693528073},
            "3": {"chi_sq_stat": 7.047367694, "p_value": 0.007938182229},
            "constant": [6, 7, 8, 9, 10],   # chi_sq_stat missing or < 1e-6
        },
        "expected_wald": {    # test_statgen.py:737-756, places=6; variant 3 is separable (fit.converged false)
            "1": {"beta": -0.81226793796, "standard_error": 2.1085483421, "z_stat": -0.3852261396, "p_value": 0.7000698784},
            "2": {"beta": -0.43659460858, "standard_error": 1.0296902941, "z_stat": -0.4240057531, "p_value": 0.6715616176},
            "not_converged": [3],
            "constant": [6, 7, 8, 9, 10],   # not converged, p NaN or |p - 1| < 1e-4
        },
        "expected_wald_dosage": {   # test_statgen.py:851-938: x = pl_dosage(PL) (places=6) and gp_dosage(GP) (places=4)
            "1": {"beta": -0.8286774, "standard_error": 2.151145, "z_stat": -0.3852261, "p_value": 0.7000699},
            "2": {"beta": -0.4431764, "standard_error": 1.045213, "z_stat": -0.4240058, "p_value": 0.6715616},
            "not_converged": [3],
            "constant": [6, 7, 8, 9, 10],
        },
        "expected_lrt": {     # test_statgen.py:958-985
            "1": {"beta": -0.81226793796, "chi_sq_stat": 0.1503349167, "p_value": 0.6982155052},
            "2": {"beta": -0.43659460858, "chi_sq_stat": 0.1813968574, "p_value": 0.6701755415},
            "not_converged": [3],
            "constant": [6, 7, 8, 9, 10],
        },
    }
    with open(os.path.join(OUT, "regression_logistic.json"), "w") as f:
        json.dump(doc, f, indent=1)
    print("wrote regression_logistic.json:", len(gt), "variants x", len(samples), "samples")
    epacts()


def epacts():
    """regressionLogisticEpacts.{vcf,cov,fam}: 2535 samples x 5 variants (test_statgen.py:1694-1862).  Columns follow the
    VCF sample order; is_case / is_female as import_fam reads them (methods/impex.py import_fam: sex '1' male, '2'
    female, else missing; phenotype '1' control, '2' case, '0' / '-9' / 'NA' missing)."""
    import numpy as np

    samples, gt, pos = None, [], []
    with open(f"{RES}/regressionLogisticEpacts.vcf") as f:
        for line in f:
            if line.startswith("##"):
                continue
            parts = line.rstrip("\n").split("\t")
            if line.startswith("#"):
                samples = parts[9:]
                continue
            pos.append(int(parts[1]))
            row = []
            for cell in parts[9:]:
                g = cell.split(":")[0]
                row.append(-1 if "." in g else sum(int(a) for a in g.replace("|", "/").split("/")))
            gt.append(row)
    cov = {}
    with open(f"{RES}/regressionLogisticEpacts.cov") as f:
        f.readline()
        for line in f:
            r = line.split()
            if r:
                cov[r[0]] = (float(r[1]), float(r[2]))
    fam = {}
    with open(f"{RES}/regressionLogisticEpacts.fam") as f:
        for line in f:
            r = line.split()
            if r:
                fam[r[1]] = (r[4], r[5])
    nan = float("nan")
    pc = np.array([cov.get(s, (nan, nan)) for s in samples])
    is_female = np.array([{"1": 0.0, "2": 1.0}.get(fam.get(s, ("0", "0"))[0], nan) for s in samples])
    is_case = np.array([{"1": 0.0, "2": 1.0}.get(fam.get(s, ("0", "0"))[1], nan) for s in samples])
    np.savez_compressed(
        os.path.join(OUT, "logistic_epacts.npz"), gt=np.array(gt, dtype=np.int8), position=np.array(pos),
        pc1=pc[:, 0], pc2=pc[:, 1], is_female=is_female, is_case=is_case,
        # test_statgen.py:1722-1757 (wald), :1759-1778 (lrt), :1826-1834 (score), :1837-1862 (firth); columns per variant
        wald=np.array([[-0.097476, 0.087478, -1.1143, 0.26516], [-0.052632, 0.11272, -0.46691, 0.64056],
                       [-0.15598, 0.079508, -1.9619, 0.049779], [-0.88059, 0.83769, -1.0512, 0.29316],
                       [0.54921, 0.4517, 1.2159, 0.22403]]),
        wald_rel=np.array([[1e-4] * 4, [1e-4] * 4, [1e-4] * 4, [1e-4, 1e-2, 1e-2, 1e-2], [1e-4, 1e-3, 1e-3, 1e-3]]),
        lrt_p=np.array([0.26475, 0.64046, 0.049675, 0.26984, 0.21692]),
        score=np.array([[1.242482, 0.2649933], [0.218038, 0.6405389], [3.850985, 0.04971679], [1.175474, 0.2782793],
                        [1.514245, 0.2184924]]),
        firth=np.array([[-0.097079, 0.26593], [-0.052301, 0.64197], [-0.15567, 0.04991], [-0.7524, 0.30731],
                        [0.5258, 0.22562]]))
    print("wrote logistic_epacts.npz:", len(gt), "variants x", len(samples), "samples")


if __name__ == "__main__":
    main()

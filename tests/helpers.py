"""Shared helpers for the parity tests (golden loaders, comparator)."""
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_regression_linear(pheno_missing_zero=True, fam_pheno=None):
    """The reference's 8x10 golden case, aligned by sample id the way the reference tests do
    (test_statgen.py:245-255: pheno[mt.s], covariates[mt.s]; missing='0' for the pheno).

    Returns x [10, 8] float64 (NaN missing), y [8] (NaN missing), cov [8, 2] (NaN missing), doc.
    """
    with open(os.path.join(GOLDEN, "regression_linear.json")) as f:
        doc = json.load(f)
    samples = doc["samples"]
    x = np.array([[np.nan if g is None else float(g) for g in row] for row in doc["gt_n_alt_alleles"]])
    if fam_pheno is None:
        y = []
        for s in samples:
            v = doc["pheno_table"].get(s)
            if v is None or (pheno_missing_zero and v == 0.0):
                y.append(np.nan)
            else:
                y.append(v)
    elif fam_pheno == "is_case":  # import_fam default: '1' control -> False, '2' case -> True, '0'/-9 missing
        code = {"1": 0.0, "2": 1.0}
        y = [code.get(doc["fam_table"].get(s, {}).get("pheno_code"), np.nan) for s in samples]
    elif fam_pheno == "quant":  # quant_pheno=True, missing='0'
        y = []
        for s in samples:
            c = doc["fam_table"].get(s, {}).get("pheno_code")
            y.append(np.nan if c in (None, "0") else float(c))
    cov = np.array([doc["cov_table"].get(s, [np.nan, np.nan]) for s in samples])
    return x, np.array(y), cov, doc


def pl_dosage(doc):
    """hl.pl_dosage (expr/functions.py:1490-1530): (b' + 2c') / (a' + b' + c'), x' = 10^(-x/10)."""
    out = []
    for row in doc["pl"]:
        r = []
        for pl in row:
            if pl is None:
                r.append(np.nan)
            else:
                a, b, c = (10.0 ** (-v / 10.0) for v in pl)
                r.append((b + 2 * c) / (a + b + c))
        out.append(r)
    return np.array(out)


def gp_dosage(doc):
    """hl.gp_dosage (expr/functions.py:1470-1488): GP[1] + 2 GP[2]; missing where import_gen drops the triple."""
    return np.array([[np.nan if t is None else t[1] + 2.0 * t[2] for t in row] for row in doc["gp"]])


def assert_fields_close(got, want, rel=1e-6, rel_p=1e-5, t_floor=0.0, ctx="", ytx_abs=None):
    """Parity gate: n exact; sum_x / y_transpose_x / beta / standard_error / t_stat within `rel`
    (the reference's own `_same` comparator, oracle.d_eq); p_value within `rel_p`.

    `t_floor` > 0 additionally accepts |delta t| <= t_floor (FP64 roundoff floor on a statistic whose
    true value is ~0: beta/se relative error is unbounded there in ANY float64 implementation,
    the reference included).  `ytx_abs` (array, optional) additionally accepts |delta y_transpose_x| <= ytx_abs: the
    many-phenotype precision profile's stated floor for dot products that cancel to ~0, 5e-7 standard errors of x.y
    (`ytx_floor_of` below computes it).
    """
    from oracle.linreg_oracle import d_eq

    assert np.array_equal(np.asarray(got["n"]), np.asarray(want["n"])), ctx + " n differs"
    for f in ("sum_x", "y_transpose_x", "standard_error"):
        ok = d_eq(got[f], want[f], rel)
        if f == "y_transpose_x" and ytx_abs is not None:
            with np.errstate(invalid="ignore"):
                ok |= np.abs(np.asarray(got[f]) - np.asarray(want[f])) <= ytx_abs
        assert ok.all(), f"{ctx} {f}: {np.count_nonzero(~ok)} mismatches, first at {np.argwhere(~ok)[:3].tolist()}"
    se = np.asarray(want["standard_error"], dtype=np.float64)
    for f, scale in (("beta", se), ("t_stat", np.ones_like(se))):
        ok = d_eq(got[f], want[f], rel)
        if t_floor > 0:
            with np.errstate(invalid="ignore"):
                ok |= np.abs(np.asarray(got[f]) - np.asarray(want[f])) <= t_floor * scale
        assert ok.all(), f"{ctx} {f}: {np.count_nonzero(~ok)} mismatches, first at {np.argwhere(~ok)[:3].tolist()}"
    ok = d_eq(got["p_value"], want["p_value"], rel_p)
    if t_floor > 0:
        with np.errstate(invalid="ignore"):
            ok |= np.abs(np.asarray(got["t_stat"]) - np.asarray(want["t_stat"])) <= t_floor
    assert ok.all(), f"{ctx} p_value: {np.count_nonzero(~ok)} mismatches, first at {np.argwhere(~ok)[:3].tolist()}"


def ytx_floor_of(x, cov, se, floor=5e-7):
    """`floor` standard errors of x.y per (variant, phenotype): floor * se * xxp with xxp = x.x - |Q'x|^2 of the
    mean-imputed rows `x` [M, n] (NaN = missing) over the kept samples' covariates `cov` [n, K]."""
    x = np.array(x, dtype=np.float64)
    miss = np.isnan(x)
    with np.errstate(invalid="ignore", divide="ignore"):
        mean = np.where(miss, 0.0, x).sum(axis=1) / (~miss).sum(axis=1)
    x[miss] = np.broadcast_to(mean[:, None], x.shape)[miss]
    q = np.linalg.qr(cov)[0] if cov.shape[1] else np.zeros((cov.shape[0], 0))
    qtx = x @ q
    xxp = (x * x).sum(axis=1) - (qtx * qtx).sum(axis=1)
    return floor * np.asarray(se) * xxp[:, None]

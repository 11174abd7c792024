"""BASELINE config 1 end to end (1,000 samples x 10,000 variants, 1 phenotype, 2 covariates): the public call on
host-resident .bed rows and on a resident store, next to the CPU port on all host threads."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import hail_b200 as hb
from hail_b200 import _lib
from oracle import bed as obed
from oracle import c_oracle

N, M = 1000, 10000
mt = hb.balding_nichols_model(3, N, M, seed=1)
rng = np.random.default_rng(0)
y, c1 = rng.standard_normal(N), rng.standard_normal(N)
dos = mt.genotypes.to_dosage().astype(np.float64)
rows = obed.encode_rows(dos)
host = hb.MatrixTable(hb.HostBedGenotypes(rows, N), cols={"y": y, "c1": c1})
res = mt.annotate_cols(y=y, c1=c1)
out = {}
for name, src in (("host_bed", host), ("resident", res)):
    ts = []
    for i in range(12):
        t0 = time.perf_counter()
        ht = hb.linear_regression_rows(y=src.y, x=src.GT.n_alt_alleles(), covariates=[1.0, src.c1])
        ts.append(time.perf_counter() - t0)
    out[name] = {"ms_median": 1e3 * float(np.median(ts[2:])), "ms_min": 1e3 * float(np.min(ts[2:])), "kernel": _lib.context(0).last_kernel,
                 "genotypes_per_s": N * M / float(np.median(ts[2:]))}
cov = np.column_stack([np.ones(N), c1])
prep = c_oracle.prepare(y[:, None], cov)
threads = os.cpu_count()
ts = []
for i in range(8):
    t0 = time.perf_counter()
    prep = c_oracle.prepare(y[:, None], cov)
    c_oracle.run_prepared(rows, prep, n_threads=threads)
    ts.append(time.perf_counter() - t0)
out["cpu_port"] = {"ms_median": 1e3 * float(np.median(ts[1:])), "threads": threads, "genotypes_per_s": N * M / float(np.median(ts[1:]))}
print(json.dumps(out))

"""Seeded Balding-Nichols style synthetic genotypes, generated directly in the device store.

Follows the model of `hl.balding_nichols_model` (hail/python/hail/methods/statgen.py:3984-4291): population of
each sample ~ Cat(pi) (SG:4254); ancestral allele frequency ~ U(0.1, 0.9) (SG:4189); per-population frequency
~ Beta(p(1-F)/F, (1-p)(1-F)/F) (SG:4186, 4269-4271); genotype ~ Cat(q^2, 2pq, p^2) (SG:4290-4291).  Hail's
threefry stream cannot be reproduced outside Hail, so draws use our own counter-based generator: the
per-variant frequencies come from numpy's Philox (seeded), the per-call draw from the device kernel
`bn_fill_kernel` (csrc/pack.cu), whose integer thresholds are computed here.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from .genotypes import PackedGenotypes
from .matrixtable import MatrixTable


def bn_parameters(n_populations, n_samples, n_variants, fst=None, pop_dist=None, af_range=(0.1, 0.9),
                  missing_rate=0.0, seed=0, first_variant=0):
    """Host-side draws: per-sample population (uint8 [N]) and per-variant integer thresholds (uint32 [M, pops, 3])."""
    K = int(n_populations)
    fst = np.full(K, 0.1) if fst is None else np.asarray(fst, dtype=np.float64)
    pop_dist = np.full(K, 1.0 / K) if pop_dist is None else np.asarray(pop_dist, dtype=np.float64) / np.sum(pop_dist)
    rng_pop = np.random.Generator(np.random.Philox(key=[seed, 0xB0B]))
    pop = rng_pop.choice(K, size=n_samples, p=pop_dist).astype(np.uint8)
    # per-variant streams are keyed by the global variant block so any shard regenerates the same values
    af = np.empty((n_variants, K))
    block = 1 << 16
    v = first_variant
    done = 0
    while done < n_variants:
        b = v // block
        lo_in_block = v - b * block
        take = min(n_variants - done, block - lo_in_block)
        rng = np.random.Generator(np.random.Philox(key=[seed, 0xAF0000 + b]))
        anc = rng.uniform(af_range[0], af_range[1], size=block)
        a = anc[:, None] * (1 - fst)[None, :] / fst[None, :]
        bb = (1 - anc)[:, None] * (1 - fst)[None, :] / fst[None, :]
        pk = rng.beta(a, bb)
        af[done:done + take] = pk[lo_in_block:lo_in_block + take]
        done += take
        v += take
    q = 1.0 - af
    keep = 1.0 - missing_rate
    t0 = np.full_like(af, missing_rate)
    t1 = t0 + keep * q * q
    t2 = t1 + keep * 2.0 * af * q
    th = np.stack([t0, t1, t2], axis=2)
    thresholds = np.minimum(np.rint(th * 65536.0), 65536.0).astype(np.uint32)
    return pop, thresholds, af


def bn_fill(out: PackedGenotypes, pop, thresholds, seed=0, first_variant=0, chunk_variants=1 << 16):
    """Fill `out` rows [0, M) with calls for global variants [first_variant, first_variant + M)."""
    dev = out.device
    ctx = _lib.context(dev.index)
    M, N = out.n_variants, out.n_samples
    assert thresholds.shape[0] == M and len(pop) == N
    n_pops = thresholds.shape[1]
    with torch.cuda.device(dev):
        d_pop = torch.from_numpy(np.ascontiguousarray(pop, dtype=np.uint8)).to(dev)
        for lo in range(0, M, chunk_variants):
            hi = min(M, lo + chunk_variants)
            d_th = torch.from_numpy(np.ascontiguousarray(thresholds[lo:hi]).view(np.int32)).to(dev)
            ctx.check(ctx.lib.lrr_bn_fill(ctx.handle, d_th.data_ptr(), n_pops, d_pop.data_ptr(), hi - lo,
                                          first_variant + lo, N, seed, out.data[lo:hi].data_ptr(), out.stride,
                                          out.row_flags[lo:hi].data_ptr() if out.row_flags is not None else None,
                                          torch.cuda.current_stream(dev).cuda_stream))
    return out


def balding_nichols_model(n_populations, n_samples, n_variants, fst=None, pop_dist=None, af_range=(0.1, 0.9),
                          missing_rate=0.0, seed=0, device=0) -> MatrixTable:
    """`hl.balding_nichols_model(n_populations, n_samples, n_variants)` with the calls resident in HBM."""
    pop, th, af = bn_parameters(n_populations, n_samples, n_variants, fst, pop_dist, af_range, missing_rate, seed)
    gt = bn_fill(PackedGenotypes.empty(n_variants, n_samples, device), pop, th, seed=seed)
    rows = {
        "locus": np.array([("1", i + 1) for i in range(n_variants)], dtype=object),
        "alleles": np.array([("A", "C")] * n_variants, dtype=object),
        "ancestral_af_by_pop": af,
    }
    cols = {"sample_idx": np.arange(n_samples, dtype=np.float64), "pop": pop.astype(np.float64)}
    return MatrixTable(gt, rows=rows, cols=cols, row_key=("locus", "alleles"), col_key=("sample_idx",))

"""Kernel-only time of lrr_run_logit (Wald) for 12 - 19 covariates: register form against the tiled form (tuning build:
LRR_LOGIT_TILED_FROM).  The model of the last public call stays set in the context, so the kernel is re-launched directly."""
import ctypes, os, sys, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hail_b200 as hb
from hail_b200 import _lib

N, M = 400_000, 592
ctx = _lib.context(0)
dev = torch.device("cuda", 0)
for K in (10, 12, 13, 15, 17, 19):
    mt = hb.balding_nichols_model(3, N, M, missing_rate=0.01, seed=5)
    rng = np.random.default_rng(K)
    cov = np.column_stack([np.ones(N)] + [rng.standard_normal(N) for _ in range(K - 1)])
    y = (rng.random(N) < 1 / (1 + np.exp(-(0.3 * cov[:, 1] - 0.2)))).astype(np.float64)
    mt = mt.annotate_cols(y=y, **{f"c{k}": cov[:, k] for k in range(1, K)})
    ht = hb.logistic_regression_rows("wald", mt.y, mt.GT.n_alt_alleles(), [1.0] + [mt[f"c{k}"] for k in range(1, K)])
    g = mt.genotypes
    outs = {f: torch.empty(M, dtype=torch.float64, device=dev) for f in ("beta", "standard_error", "z_stat", "p_value")}
    outs["n_iterations"] = torch.empty(M, dtype=torch.int32, device=dev)
    outs["converged"] = torch.empty(M, dtype=torch.uint8, device=dev)
    outs["exploded"] = torch.empty(M, dtype=torch.uint8, device=dev)
    out = _lib.LogitOut()
    for f, t in outs.items():
        setattr(out, f, t.data_ptr())
    st = torch.cuda.current_stream(dev).cuda_stream
    res = {}
    for tiled_from in (21, 11):
        os.environ["LRR_LOGIT_TILED_FROM"] = str(tiled_from)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ms = []
        for rep in range(3):
            e0.record()
            ctx.check(ctx.lib.lrr_run_logit(ctx.handle, g.data.data_ptr(), M, g.stride, N, 1, 25, 1e-6, ctypes.byref(out), st))
            e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        res["register" if tiled_from == 21 else "tiled"] = round(min(ms), 2)
        res["beta_equal_public_" + ("register" if tiled_from == 21 else "tiled")] = bool(np.allclose(outs["beta"].cpu().numpy(), ht.beta, rtol=1e-9, atol=1e-12, equal_nan=True))
    print(json.dumps({"K": K, "kernel_ms_per_592_variants": res, "mean_iterations": float(ht.fit["n_iterations"].mean())}), flush=True)

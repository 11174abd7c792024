"""hail_b200 -- B200-native `linear_regression_rows` (one hot path of Hail, behind the same call signature).

Importing the package does not need a GPU; any compute call loads `liblrr_b200.so` (built in-tree for sm_100a)
and fails loudly if it is missing -- there is no CPU fallback.
"""
from .matrixtable import (ColumnExpression, EntryExpression, ExpressionException, MatrixTable, RowExpression, Struct,
                          Table)
from .statgen import FatalError, lambda_gc, linear_regression_rows, _get_regression_row_fields, _warn_if_no_intercept


def __getattr__(name):  # lazy: these import torch-side helpers
    if name in ("PackedGenotypes", "HostBedGenotypes", "DenseDosage", "CompactDosage", "packed_stride"):
        from . import genotypes
        return getattr(genotypes, name)
    if name in ("import_plink", "export_plink", "import_fam", "import_bgen", "read_bgen"):
        from . import impex
        return getattr(impex, name)
    if name == "logistic_regression_rows":
        from .logreg import logistic_regression_rows
        return logistic_regression_rows
    if name == "hwe_normalized_pca":
        from .pca import hwe_normalized_pca
        return hwe_normalized_pca
    if name in ("balding_nichols_model", "bn_parameters", "bn_fill"):
        from . import bn
        return getattr(bn, name)
    raise AttributeError(name)


__all__ = ["linear_regression_rows", "lambda_gc", "hwe_normalized_pca", "logistic_regression_rows", "MatrixTable", "Table", "FatalError", "ExpressionException", "PackedGenotypes",
           "HostBedGenotypes", "CompactDosage", "import_plink", "export_plink", "import_fam", "import_bgen", "balding_nichols_model"]

"""The reference's plugin entry for this path: a relational function looked up BY NAME from a config object.

In the reference, `hl.linear_regression_rows` does not call the regression directly: it builds a config dict
(`hail/python/hail/methods/statgen.py:394-401`), wraps it in `ir.MatrixToTableApply(child, config)` (`ir/table_ir.py:948-997`,
which also derives the result schema from the config) and the JVM extracts a `MatrixToTableFunction` from the JSON by its
`"name"` type hint (`expr/ir/functions/RelationalFunctions.scala:112-138`, classes `LinearRegressionRowsSingle` /
`LinearRegressionRowsChained`, `methods/LinearRegression.scala:18-44, 198-224`).  There is no JVM here, so the same
contract is kept in Python: `lookup_matrix_to_table(config)` returns the function object, `typ()` gives the result
schema, `execute(mt)` runs it (host prologue + CUDA library).  A Hail-side adapter would forward exactly
this config.
"""
from __future__ import annotations

import json

import numpy as np

from .matrixtable import EntryExpression, MatrixTable, Table

STAT_SCHEMA = ["n", "sum_x", "y_transpose_x", "beta", "standard_error", "t_stat", "p_value"]


class MatrixToTableFunction:
    """abstract class MatrixToTableFunction { typ; execute } (RelationalFunctions.scala:24-32)"""

    name = None

    def typ(self, child: MatrixTable):
        raise NotImplementedError

    def execute(self, mt: MatrixTable, **opts) -> Table:
        raise NotImplementedError

    def preserves_partition_counts(self) -> bool:
        return True   # LR:44 / :224


class _LinearRegressionRows(MatrixToTableFunction):
    chained = False

    def __init__(self, yFields, xField, covFields, rowBlockSize, passThrough):
        self.yFields = yFields
        self.xField = xField
        self.covFields = list(covFields)
        self.rowBlockSize = int(rowBlockSize)
        self.passThrough = list(passThrough)
        groups = self.yFields if self.chained else [self.yFields]
        if not isinstance(groups, (list, tuple)) or any(not isinstance(g, (list, tuple)) for g in groups):
            raise ValueError(f"{self.name}: 'yFields' must be a list of " + ("lists of field names" if self.chained else "field names"))
        if any(not isinstance(f, str) for g in groups for f in g):
            raise ValueError(f"{self.name}: 'yFields' must hold field names")

    def typ(self, child: MatrixTable):
        """Row key, pass-through fields, then the statistics (LR:26-42 / :206-222), in this order."""
        return list(child.row_key) + self.passThrough + STAT_SCHEMA

    def execute(self, mt: MatrixTable, **opts) -> Table:
        from . import statgen
        groups = self.yFields if self.chained else [self.yFields]
        for f in [f for g in groups for f in g] + self.covFields:
            if f not in mt.col:
                raise KeyError(f"{self.name}: MatrixTable has no column field {f!r}")
        for f in self.passThrough:
            if f not in mt.row:
                raise KeyError(f"{self.name}: MatrixTable has no row field {f!r}")
        x = mt[self.xField]
        if not isinstance(x, EntryExpression):
            raise KeyError(f"{self.name}: {self.xField!r} is not an entry field")
        col = lambda f: np.asarray(mt.col[f], dtype=np.float64)
        y_vals = [[col(f) for f in g] for g in groups]
        cov_vals = [col(f) for f in self.covFields]
        return statgen._execute(mt, x, y_vals, cov_vals, self.chained, self.passThrough, **opts)


class LinearRegressionRowsSingle(_LinearRegressionRows):
    """case class LinearRegressionRowsSingle(yFields: Seq[String], xField, covFields, rowBlockSize, passThrough) (LR:18-24)"""
    name = "LinearRegressionRowsSingle"
    chained = False


class LinearRegressionRowsChained(_LinearRegressionRows):
    """case class LinearRegressionRowsChained(yFields: Seq[Seq[String]], ...) (LR:198-204)"""
    name = "LinearRegressionRowsChained"
    chained = True


_REGISTRY = {c.name: c for c in (LinearRegressionRowsSingle, LinearRegressionRowsChained)}
_KEYS = ("yFields", "xField", "covFields", "rowBlockSize", "passThrough")


def lookup_matrix_to_table(config) -> MatrixToTableFunction:
    """RelationalFunctions.lookupMatrixToTable (RelationalFunctions.scala:112-138): `config` is the dict of
    statgen.py:394-401 or its JSON string; the class is chosen by the "name" type hint."""
    if isinstance(config, (str, bytes)):
        config = json.loads(config)
    if not isinstance(config, dict) or "name" not in config:
        raise ValueError("relational function config needs a 'name'")
    cls = _REGISTRY.get(config["name"])
    if cls is None:
        raise ValueError(f"no MatrixToTableFunction registered under {config['name']!r} (known: {sorted(_REGISTRY)})")
    missing = [k for k in _KEYS if k not in config]
    unknown = [k for k in config if k not in _KEYS + ("name",)]
    if missing or unknown:
        raise ValueError(f"{config['name']}: bad config (missing {missing}, unknown {unknown})")
    return cls(**{k: config[k] for k in _KEYS})


def matrix_to_table_apply(mt: MatrixTable, config, **opts) -> Table:
    """`Table(ir.MatrixToTableApply(mt._mir, config))` (statgen.py:402): look the function up and run it on `mt`."""
    return lookup_matrix_to_table(config).execute(mt, **opts)

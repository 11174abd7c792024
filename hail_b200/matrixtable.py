"""A deliberately small stand-in for the parts of Hail's data model that `linear_regression_rows` touches.

It is NOT a re-implementation of Hail's expression language or MatrixTable (out of scope).
It carries exactly what the call needs so that user code and the parity tests read like the reference's:

    mt = hb.import_plink(bed, bim, fam)                       # methods/impex.py:2505
    mt = mt.annotate_cols(pheno=..., cov1=...)
    ht = hb.linear_regression_rows(y=mt.pheno, x=mt.GT.n_alt_alleles(), covariates=[1.0, mt.cov1])

* column fields are float64 numpy arrays with NaN = missing (hl.missing);
* row fields are arbitrary numpy arrays / lists (only copied through: key + pass_through, LR:128,168);
* the one entry field is the packed call matrix `GT` (genotypes.PackedGenotypes), and the one entry
  expression understood is `GT.n_alt_alleles()` (typed_expressions.py:3587-3607).
"""
from __future__ import annotations

from collections import OrderedDict

import numpy as np


class ExpressionException(Exception):
    """hail.expr.expressions.ExpressionException"""


class Expression:
    axes: frozenset = frozenset()
    source = None
    is_nested_field = False
    name = None


class ColumnExpression(Expression):
    """Column-indexed float64 expression (NaN = missing).  Arithmetic with scalars / other columns is eager."""

    axes = frozenset({"col"})

    def __init__(self, source, values, name=None, is_field=False):
        self.source = source
        self.values = np.asarray(values, dtype=np.float64)
        self.name = name
        self.is_nested_field = is_field

    def _bin(self, other, op):
        if isinstance(other, ColumnExpression):
            if other.source is not self.source:
                raise ExpressionException("column expressions come from different MatrixTables")
            other = other.values
        elif isinstance(other, Expression):
            raise ExpressionException("cannot combine a column expression with a non-column expression")
        return ColumnExpression(self.source, op(self.values, other))

    def __add__(self, o): return self._bin(o, np.add)
    def __radd__(self, o): return self._bin(o, lambda a, b: b + a)
    def __sub__(self, o): return self._bin(o, np.subtract)
    def __rsub__(self, o): return self._bin(o, lambda a, b: b - a)
    def __mul__(self, o): return self._bin(o, np.multiply)
    def __rmul__(self, o): return self._bin(o, lambda a, b: b * a)
    def __truediv__(self, o): return self._bin(o, np.divide)
    def __neg__(self): return ColumnExpression(self.source, -self.values)
    # comparisons give boolean columns (as 0/1 with NaN where undefined), usable in filter_cols / or_missing
    def __ge__(self, o): return self._bin(o, lambda a, b: np.where(np.isnan(a + b), np.nan, (a >= b).astype(float)))
    def __le__(self, o): return self._bin(o, lambda a, b: np.where(np.isnan(a + b), np.nan, (a <= b).astype(float)))
    def __gt__(self, o): return self._bin(o, lambda a, b: np.where(np.isnan(a + b), np.nan, (a > b).astype(float)))
    def __lt__(self, o): return self._bin(o, lambda a, b: np.where(np.isnan(a + b), np.nan, (a < b).astype(float)))

    def or_missing_unless(self, cond):
        """`hl.case().when(cond, self).or_missing()`: keep the value where cond is true, else missing."""
        c = cond.values if isinstance(cond, ColumnExpression) else np.asarray(cond, dtype=np.float64)
        return ColumnExpression(self.source, np.where(c == 1.0, self.values, np.nan))


class RowExpression(Expression):
    axes = frozenset({"row"})

    def __init__(self, source, values, name, is_nested_field=True, path=()):
        self.source = source
        self.values = values
        self.name = name
        self.is_nested_field = is_nested_field
        self.path = path

    def __getattr__(self, item):
        vals = object.__getattribute__(self, "values")
        if isinstance(vals, dict) and item in vals:
            return RowExpression(self.source, vals[item], item, True, self.path + (self.name,))
        raise AttributeError(item)

    def length(self):
        """A complex (non-field) row expression, e.g. `mt.filters.length()`."""
        return RowExpression(self.source, np.array([len(v) for v in self.values]), None, False)


class EntryExpression(Expression):
    axes = frozenset({"row", "col"})

    def __init__(self, source, kind):
        self.source = source
        self.kind = kind  # 'n_alt_alleles' (packed calls) or 'dosage' (dense float64 entries)


class CallExpression(Expression):
    axes = frozenset({"row", "col"})

    def __init__(self, source):
        self.source = source

    def n_alt_alleles(self):
        return EntryExpression(self.source, "n_alt_alleles")


class MatrixTable:
    def __init__(self, genotypes, rows=None, cols=None, row_key=(), col_key=(), col_index=None, entry_aliases=None):
        self.genotypes = genotypes  # PackedGenotypes (all stored samples)
        self._rows = OrderedDict(rows or {})
        self._cols = OrderedDict(cols or {})
        self.row_key = tuple(row_key)
        self.col_key = tuple(col_key)
        # kept columns as indices into the packed store (filter_cols is zero-copy)
        self.col_index = np.arange(genotypes.n_samples) if col_index is None else np.asarray(col_index)
        self._entry_aliases = dict(entry_aliases or {})
        for k, v in self._rows.items():
            if not isinstance(v, dict) and len(v) != genotypes.n_variants:
                raise ValueError(f"row field {k!r} has {len(v)} values for {genotypes.n_variants} rows")
        for k, v in self._cols.items():
            for kk, vv in (v.items() if isinstance(v, dict) else [(k, v)]):
                if len(vv) != len(self.col_index):
                    raise ValueError(f"col field {kk!r} has {len(vv)} values for {len(self.col_index)} columns")

    # ---- shape -------------------------------------------------------------------------------
    def count_rows(self): return self.genotypes.n_variants
    def count_cols(self): return len(self.col_index)
    def count(self): return self.count_rows(), self.count_cols()

    @property
    def row(self): return self._rows
    @property
    def col(self): return self._cols

    # ---- field access ------------------------------------------------------------------------
    def __getattr__(self, item):
        if item.startswith("_"):
            raise AttributeError(item)
        return self[item]

    def __getitem__(self, item):
        if item == "GT":
            if type(self.genotypes).__name__ in ("DenseDosage", "CompactDosage"):
                raise ExpressionException("this MatrixTable holds a dense dosage entry field `x`, not calls")
            return CallExpression(self)
        if item in ("x", "dosage") and type(self.genotypes).__name__ in ("DenseDosage", "CompactDosage"):
            return EntryExpression(self, "dosage")   # a float64 entry field (statgen.py:229: any float64 x)
        if item in self._entry_aliases:
            return EntryExpression(self, self._entry_aliases[item])
        if item in self._cols:
            v = self._cols[item]
            if isinstance(v, dict):
                return _ColStruct(self, v)
            if np.asarray(v).dtype.kind in "fiub":
                return ColumnExpression(self, _as_float_col(v), item, True)
            return v
        if item in self._rows:
            return RowExpression(self, self._rows[item], item, True)
        raise AttributeError(f"MatrixTable has no field {item!r}")

    # ---- annotate / filter -------------------------------------------------------------------
    def _copy(self, **kw):
        args = dict(genotypes=self.genotypes, rows=self._rows, cols=self._cols, row_key=self.row_key,
                    col_key=self.col_key, col_index=self.col_index, entry_aliases=self._entry_aliases)
        args.update(kw)
        return MatrixTable(**args)

    def annotate_cols(self, **named):
        cols = OrderedDict(self._cols)
        for k, v in named.items():
            if isinstance(v, ColumnExpression):
                v = v.values
            elif isinstance(v, dict):
                v = {kk: (vv.values if isinstance(vv, ColumnExpression) else np.asarray(vv)) for kk, vv in v.items()}
            elif np.isscalar(v):
                v = np.full(len(self.col_index), v)
            cols[k] = v
        return self._copy(cols=cols)

    def annotate_rows(self, **named):
        rows = OrderedDict(self._rows)
        for k, v in named.items():
            rows[k] = v.values if isinstance(v, RowExpression) else v
        return self._copy(rows=rows)

    def annotate_entries(self, **named):
        al = dict(self._entry_aliases)
        for k, v in named.items():
            if not isinstance(v, EntryExpression):
                raise ExpressionException("only GT.n_alt_alleles() is supported as an entry expression")
            al[k] = v.kind
        return self._copy(entry_aliases=al)

    def filter_cols(self, cond):
        c = cond.values if isinstance(cond, ColumnExpression) else np.asarray(cond, dtype=np.float64)
        keep = np.nan_to_num(c, nan=0.0) == 1.0  # missing predicate drops the column, as in Hail
        cols = OrderedDict()
        for k, v in self._cols.items():
            cols[k] = {kk: np.asarray(vv)[keep] for kk, vv in v.items()} if isinstance(v, dict) else np.asarray(v)[keep]
        return self._copy(cols=cols, col_index=self.col_index[keep])

    def cache(self): return self


class _ColStruct:
    def __init__(self, mt, d):
        self._mt, self._d = mt, d

    def __getattr__(self, item):
        d = object.__getattribute__(self, "_d")
        if item in d:
            return ColumnExpression(self._mt, _as_float_col(d[item]), item, True)
        raise AttributeError(item)

    def values(self):
        return [ColumnExpression(self._mt, _as_float_col(v), k, True) for k, v in self._d.items()]


def _as_float_col(v):
    a = np.asarray(v)
    if a.dtype == object:  # None -> missing
        return np.array([np.nan if e is None else float(e) for e in a], dtype=np.float64)
    return a.astype(np.float64, copy=False)   # float64 columns are used as they are (expressions never write into them)


class Struct(dict):
    __getattr__ = dict.__getitem__


class Table:
    """Result rows keyed by the MatrixTable's row key (LR:26-42 / LR:206-222 schema, same field order)."""

    def __init__(self, fields: "OrderedDict[str, object]", key=(), n_rows=0):
        self._fields = OrderedDict(fields)
        self.key = tuple(key)
        self.n_rows = int(n_rows)

    @property
    def row(self): return list(self._fields)

    def __contains__(self, item): return item in self._fields
    def __getitem__(self, item): return self._fields[item]

    def __getattr__(self, item):
        if item.startswith("_"):
            raise AttributeError(item)
        try:
            return self._fields[item]
        except KeyError:
            raise AttributeError(item)

    def count(self): return self.n_rows

    def select(self, *names, **named):
        f = OrderedDict((k, self._fields[k]) for k in self.key)
        for nme in names:
            f[nme] = self._fields[nme]
        f.update(named)
        return Table(f, self.key, self.n_rows)

    def annotate(self, **named):
        """Replace / append fields (statgen.py:404-406 uses it to unwrap the single-phenotype arrays)."""
        f = OrderedDict(self._fields)
        f.update(named)
        t = Table(f, self.key, self.n_rows)
        for extra in ("n_missing", "sharded"):
            if extra in self.__dict__:
                setattr(t, extra, self.__dict__[extra])
        return t

    def collect(self):
        rows = []
        for i in range(self.n_rows):
            s = Struct()
            for k, v in self._fields.items():
                s[k] = _row_value(v, i)
            rows.append(s)
        return rows

    def to_pandas(self):
        import pandas as pd

        cols = {}
        for k, v in self._fields.items():
            if isinstance(v, np.ndarray) and v.ndim == 1:
                cols[k] = v
            else:
                cols[k] = [_row_value(v, i) for i in range(self.n_rows)]
        return pd.DataFrame(cols)

    def gather(self):
        """For the Table of a `_sharded=True` call: every rank's rows concatenated in rank order (= the row order of
        the unsharded dataset) on every rank.  Collective over the default torch.distributed group; numeric fields
        travel as tensors (NCCL on GPUs), row keys / pass-through fields as pickled objects."""
        import torch
        import torch.distributed as dist

        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            return self
        backend = dist.get_backend()
        dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
        counts = [None] * dist.get_world_size()
        dist.all_gather_object(counts, int(self.n_rows))

        def cat_array(v):
            a = np.asarray(v)
            if a.dtype.kind in "fiub":
                t = torch.from_numpy(np.ascontiguousarray(a)).to(dev)
                flat = t.reshape(a.shape[0], -1)
                from .dist import gather_rows
                out = gather_rows(flat, counts=counts)
                return out.cpu().numpy().reshape((sum(counts),) + a.shape[1:])
            parts = [None] * len(counts)
            dist.all_gather_object(parts, a)
            return np.concatenate(parts, axis=0)

        def cat(v):
            if isinstance(v, ChainedField):
                return ChainedField(cat(g) for g in v)
            if isinstance(v, dict):
                return {k: cat(x) for k, x in v.items()}
            return cat_array(v)

        t = Table(OrderedDict((k, cat(v)) for k, v in self._fields.items()), self.key, sum(counts))
        if "n_missing" in self.__dict__:
            t.n_missing = [cat_array(g) for g in self.n_missing] if isinstance(self.n_missing, list) else cat_array(self.n_missing)
        return t

    def _same(self, other, tolerance=1e-6):
        """Table._same (hail/python/hail/table.py:4384): same fields, floats equal within the D_== comparator."""
        if list(self._fields) != list(other._fields) or self.n_rows != other.n_rows:
            return False
        for k in self._fields:
            if not _values_similar(self._fields[k], other._fields[k], tolerance):
                return False
        return True


class ChainedField(list):
    """array<...> over groups: a list (one entry per group) of per-row arrays; row i is [g[i] for g in self]."""


def _row_value(v, i):
    if isinstance(v, ChainedField):
        return [_row_value(g, i) for g in v]
    if isinstance(v, dict):
        return Struct({k: _row_value(x, i) for k, x in v.items()})
    e = v[i]
    if isinstance(e, np.ndarray):
        return e.tolist()
    if isinstance(e, np.generic):
        return e.item()
    return e


def _values_similar(a, b, tol):
    if isinstance(a, ChainedField) or isinstance(b, ChainedField):
        return len(a) == len(b) and all(_values_similar(x, y, tol) for x, y in zip(a, b))
    if isinstance(a, dict):
        return a.keys() == b.keys() and all(_values_similar(a[k], b[k], tol) for k in a)
    a, b = np.asarray(a), np.asarray(b)
    if a.shape != b.shape:
        return False
    if a.dtype.kind == "f" or b.dtype.kind == "f":
        a, b = a.astype(np.float64), b.astype(np.float64)
        with np.errstate(invalid="ignore"):
            ok = (a == b) | (np.isnan(a) & np.isnan(b)) | (
                np.abs(a - b) <= 2.2250738585072014e-308 + tol * np.maximum(np.abs(a), np.abs(b)))
        return bool(ok.all())
    return bool(np.array_equal(a, b))

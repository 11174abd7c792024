"""CPU oracle for the `hl.linear_regression_rows` hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is shipped or measured as
the product: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import, link or execute it.

The oracle is a restatement (not a copy) of the reference's algorithm:

* ``hail/hail/src/is/hail/methods/LinearRegression.scala:46-195``  (Single)
* ``hail/hail/src/is/hail/methods/LinearRegression.scala:226-407`` (Chained)
* ``hail/hail/src/is/hail/stats/RegressionUtils.scala:16-58``     (mean imputation)
* ``hail/hail/src/is/hail/stats/RegressionUtils.scala:88-128``    (complete samples)

Parity pin: the reference cannot run in this container (no JVM / Spark / hail
wheel), so the oracle is pinned against the reference's own R-``lm()``-derived
golden values (``hail/python/test/hail/methods/test_statgen.py:223-284,366-424``)
and ``pT`` known answers (``hail/python/test/hail/expr/test_expr.py:3564-3568``);
see ``tests/test_oracle_golden.py``.
"""

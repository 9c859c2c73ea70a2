"""CPU oracle: a restatement of the reference's hot path in NumPy / pure Python.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is product code: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and there only as the checker (or as the
timed CPU baseline), never as the thing shipped.  The product path
(``spark-data-repair-plugin_b200/repair``) never imports this package and fails
loudly when the CUDA library is missing.

What it restates (reference = maropu/spark-data-repair-plugin @ 7701550d):

* ``detect.py``  -- ``ErrorDetectorApi.scala:128-300`` (Null / RegEx / Constraint / IQR
  detectors), ``DenialConstraints.scala:82-225`` (parser), ``errors.py:85-190,389-461``
* ``stats.py``   -- ``RepairApi.scala:34-67`` (input check), ``:106-169`` (stats + discretize),
  ``:231-273`` (freq stats), ``:280-394`` (conditional entropy), ``:396-477`` (pair selection)
* ``domain.py``  -- ``RepairApi.scala:479-675`` (cell-domain analysis), ``errors.py:507-530``
  (weak-label pruning)
* ``repair.py``  -- ``RepairApi.scala:171-211`` (NULL masking), ``model.py:533-555,677-729,
  955-1143,1398-1401`` (split, encoders, model chain, output shaping), ``errors.py:545-582``
* ``forest.py``  -- the flat-forest evaluator.  The arithmetic of the
  reference's model lives in LightGBM 3.3.1 (``bin/requirements.txt:6``), which is NOT vendored
  under /root/reference and not installed here.

Pinning status
--------------
* detectors, DC parser, freq stats, conditional entropy, attr stats, cell-domain analysis,
  NULL masking, current values, discretisation: PINNED by the reference's own known-answer
  tests transcribed in ``tests/test_oracle_kat.py`` (``RepairSuite.scala``,
  ``ErrorDetectorSuite.scala``, ``DenialConstraintsSuite.scala``, ``tests/test_errors.py``,
  ``tests/test_model.py:510-675``).
* distinct counts: the reference uses Spark's HLL++ (approximate, F5 in SURVEY.md); the oracle
  uses exact counts.  The hospital KAT (``RepairSuite.scala:144-177``) pins only the kept-column
  set, which exact counts reproduce.
* model training (LightGBM + hyperopt), category_encoders 2.2.2: PARITY UNPINNED -- third-party,
  absent.  The forest evaluator is pinned instead against scikit-learn's own
  ``HistGradientBoosting*.predict`` on the same fitted model (an independent implementation),
  and end-to-end against ``bin/testdata/adult_repair.csv`` on the error-cell set.
"""

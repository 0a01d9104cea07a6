"""CPU oracle for the ppcseq hot path.  TEST INFRASTRUCTURE ONLY.

Everything under ``oracle/`` is a checker: a CPU restatement of the arithmetic in
``/root/reference/inst/stan/negBinomial_MPI.stan`` and of the R summarisation code in
``/root/reference/R/utilities.R``.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.  The product
(``ppcseq_b200``) never does: it fails loudly when the CUDA library is missing.

PARITY STATUS: **parity unpinned** for log_prob/grad and for quantiles/flags.  The reference
ships no golden vectors for this path (its only tests are two unseeded end-to-end VB runs,
``tests/testthat/test-ppcSeq.R:26-30,51-55``) and neither R nor Stan exists in this image, so
the reference cannot be executed to produce any ("parity unpinned" by the reference).  What pins the oracles
instead, from outside their own algebra: ``oracle/stan_literal.py`` -- a literal transcription of the Stan program
through the reference's own map_rect packing (R/utilities.R:125-174, :321-359, :1455-1466), exclusion as the
subtraction the Stan code performs (:105-115), gradients by torch.autograd (Stan's mechanism), densities
cross-checked against scipy.stats (nbinom, skewnorm, norm, laplace) -- agrees with the mpmath golden values, the
NumPy and the C oracle to 1e-9 (lp) / 1e-7 (gradient) on moderate counts (``tests/test_oracle_pin.py``).  Still
unpinned: Stan Math's own rounding (and its phi > 1e5 Poisson branch in StanHeaders <= 2.21), edgeR's TMM
(``oracle/prep_np.py`` restates the published algorithm with explicit average ranks and checks the native
``ppcseq_tmm_factors`` / ``ppcseq_prep_table``; only self-generated fixtures), R's ``quantile`` itself (the type-7 definition restated in ``oracle/quantile.py``
agrees with NumPy's ``method="linear"``, SciPy's ``mquantiles(alphap=1, betap=1)`` and pandas' ``quantile`` --
three independent implementations of Hyndman-Fan definition 7 -- ``tests/test_oracle_pin.py``).
Truth for log_prob/grad is the
40-digit mpmath evaluation in ``oracle/model_mp.py`` of the Stan program's semantics; truth for
quantiles is R's documented type-7 definition restated in ``oracle/quantile.py``.  The only
reference-pinned facts are the discrete end-to-end outcomes (``tot_deleterious_outliers ==
c(0,1,0)`` for SLC16A12/CYP1A1/ART3; README.md:75-92 table), which ``tests/`` checks.
"""

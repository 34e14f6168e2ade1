"""CPU oracle for the ITensorCPD.jl CP-ALS hot path.  TEST INFRASTRUCTURE ONLY.

This package is a plain numpy/scipy (LAPACK) restatement of the reference's
algorithm for the dense and the sampled CP-ALS paths.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs
of ``bench.py`` may import it.  The product package (``itensorcpd.jl_b200``)
never imports it and has no CPU fallback.

PARITY STATUS (see DESIGN.md "Oracle pinning"):
  * ``sparse_sign`` / ``sparsestack`` (oracle/sparse_sign_port.c) are PINNED
    bit-exactly against the reference's own C sources compiled in place into
    ``oracle/_ref/libsparse_sign_ref.so`` (oracle/Makefile).
  * the index maps are PINNED against the literal pivot lists of
    ``test/pivot_mapping.jl:17-53``.
  * everything that the reference delegates to un-vendored Julia packages
    (ITensors/NDTensors contraction, LinearAlgebra LAPACK wrappers,
    StatsBase.sample, MersenneTwister) is **parity unpinned**: Julia is not
    installed here and the reference ships no golden vectors.
    The restatement calls the same LAPACK routines Julia calls (dpstrf, dpotrs,
    dgeqp3/dgelsy) and is checked against the property tests of the
    reference's own test-suite re-stated in tests/test_oracle_*.py.
"""
from . import cpals, sampled  # noqa: F401

"""CPU oracle of the NumpyVector hot path — TEST INFRASTRUCTURE ONLY.

Nothing in the product package (eigensolvers_b200/) imports this directory.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use it, and only
as the checker or as the timed CPU baseline.

Parity status: PINNED.  oracle/numpy_vector.py is checked (tests/test_oracle.py) against
  * golden vectors generated HERE by running the unmodified reference from /root/reference
    through oracle/ref_harness/make_golden.py (committed under tests/golden/), and
  * the reference's own Fortran-FEAST golden file (unittests/data_fortranCode.out), whose numbers
    are embedded in those fixtures.
oracle/krylov.py restates SciPy's gcrotmk / minres (the third-party code the reference calls,
README.md:22 pins SciPy 1.10.1; this image has 1.18.1) and is checked against SciPy itself.
"""

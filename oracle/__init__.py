"""CPU oracle for the groth16::prove() hot path of republicprotocol/zksnark-rs.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the shipped
product: only ``tests/`` (incl. its helper scripts ``gen_golden.py``,
``check_full_size.py``, ``sanitize_case.py``), ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``cpu_best_effort`` / ``--impl reference`` legs of ``bench.py``
may import or execute it, and only as
the checker / the timed CPU baseline, never as the thing that produces a result
the product returns.

Two restatements live here:

* Oracle A (this Python package): a slow, literal, big-integer restatement of
  the reference's generic code (``CoefficientPoly``, ``polynomial_division``,
  ``dft``/``idft``, the ``.zk`` parser, ``setup``/``prove``/``verify``) over both
  back-ends the reference has: the toy field Z251 (its own test fixture) and
  BN254 (crate ``bn`` 0.4.3, restated from the published curve definition).
* Oracle B (``oracle_b.c``): the same ``prove`` algorithm in C with 256-bit
  Montgomery arithmetic, fast enough to be the timed CPU baseline.
* Oracle F (``oracle_fast.c``): NOT the reference's algorithm -- the same proof
  computed with an NTT and Pippenger on all host threads, pinned bit for bit
  against A and B (``tests/test_oracle_fast.py``).  ``bench.py`` times it as the
  ``cpu_best_effort`` context number; it is never the reference arm.

Pinning status
--------------
* Generic polynomial / protocol code: PINNED against every golden vector the
  reference's own tests hold (all over Z251) -- see ``tests/test_oracle_kats.py``.
* BN254 values (Fr products, G1/G2 coordinates, proofs): **parity unpinned**.
  The reference has no fixed-value test at the ``bn`` boundary (only
  ``verify(..) == true`` under fresh randomness, src/groth16/fr.rs:240-416) and
  the crate cannot be built here (no cargo/rustc, ``bn`` not vendored).  The BN254
  layer is pinned by us instead: published curve constants and known answers,
  Python-vs-C++ cross-check, group-law identities, and end-to-end
  ``verify == true`` through an independent optimal-ate pairing.
"""

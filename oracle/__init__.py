"""CPU restatement (oracle) of the Log-ODE hot path of log-neural-cdes.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is on the product path:
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it, and there only as the
checker (or as the timed CPU baseline), never as the thing shipped.

Pinning status
--------------
* ``oracle.hall``: PINNED.  Checked for exact equality against the reference's
  own ``data_dir/hall_set.py`` (the one file of the path that runs offline) via
  the fixtures in ``tests/golden/hall_*.npz`` (made by
  ``tests/golden/make_golden.py`` importing the reference in this container).
* ``oracle.tensor_algebra`` / ``oracle.paths`` (signax 0.1.1 semantics, which is
  an un-vendored third-party dependency; the reference holds no test vectors
  for it): PARITY UNPINNED by reference outputs.  Anchored instead on exact
  algebraic identities (Chen, log∘exp, Lévy areas, ``t2l @ l2t = I``) and on
  exact integer known-answer tests (toy-style paths).
* ``oracle.log_ncde`` / ``oracle.linear_ncde`` / ``oracle.train`` (diffrax
  0.7.0, equinox 0.12.2, optax 0.2.4, roughpy 0.2.0 semantics – all absent
  offline): PARITY UNPINNED by reference outputs; restated from the call sites
  cited in each function.

Because those oracles cannot be pinned on reference outputs, each of them is
cross-checked against a second, structurally different derivation of the same
mathematics in ``tests/test_oracle_independent_checks.py`` (brute-force iterated
integrals vs Chen/Horner, exp(log S) = S, BCH, Lie-element round trip, Lie
brackets from explicit Jacobians vs the JVP formulation, hand-rolled Heun
recurrence, bracket definition / commuting-field / chunked-scan identities).
"""

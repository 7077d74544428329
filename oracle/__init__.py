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
* ``oracle.paths`` / ``oracle.log_ncde`` / ``oracle.linear_ncde`` and the loss
  functions: PINNED TO THE REFERENCE'S OWN CODE.  ``tests/golden/make_golden_ref.py``
  imports the reference's source files from /root/reference (generate_paths.py,
  datasets.py, generate_model.py, LogNeuralCDEs.py, NeuralCDEs.py,
  LinearNeuralCDEs.py, train.py) and executes them UNMODIFIED under stand-ins for
  the libraries that cannot be installed offline (``tests/golden/refshim``: jax,
  equinox, diffrax, signax, roughpy on torch float64); the outputs are the
  committed fixtures ``tests/golden/ref_*.npz`` checked by
  ``tests/test_oracle_reference_pinned.py`` (1e-12, or 5e-7 where the oracle folds
  two fp32 factors into one).  This pins everything the reference itself wrote:
  windowing / basepoint / remainder quirks, 128-sample chunking, ``pairs``, the
  searchsorted row and interval indexing, the JVP reshape, bracket pairing and
  sign, flow construction, bracket recursion, chunk / remainder scanner,
  read-outs, both losses with their Lip(2) terms, create_model's plumbing.
* What remains RESTATED, NOT PINNED (third-party packages absent offline, no
  vectors for them in the reference): signax 0.1.1's signature / log / flatten
  (``oracle.tensor_algebra``; the shim holds a second, independent
  implementation), diffrax 0.7.0's Heun step grid, end clipping and SaveAt
  interpolation, equinox 0.12.2's Linear / MLP conventions, optax 0.2.4's adam
  (``oracle.train.adam_update``), roughpy 0.2.0's basis ORDER (assumed = the
  reference's HallSet order; equal by inspection at depth <= 2).

The third-party restatements are additionally cross-checked against a second,
structurally different derivation of the same mathematics in ``tests/test_oracle_independent_checks.py`` (brute-force iterated
integrals vs Chen/Horner, exp(log S) = S, BCH, Lie-element round trip, Lie
brackets from explicit Jacobians vs the JVP formulation, hand-rolled Heun
recurrence, bracket definition / commuting-field / chunked-scan identities).
"""

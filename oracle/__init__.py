"""CPU oracle for the SCAE likelihood hot paths.  TEST INFRASTRUCTURE ONLY.

This package restates, on the CPU, the arithmetic of the two hot paths of bdsaglam/torch-scae that
``torch_scae_b200`` implements as sm_100a CUDA kernels, plus a functional restatement of the rest of the SCAE
forward/loss so that whole-model parity and the CPU baseline can be measured without the reference being present.

Who may import it: ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs.  Nothing under ``torch_scae_b200/`` imports it; the product path has no CPU fallback and raises if the CUDA
library is missing.

Parity pinning: the reference ships no golden vectors (its tests are shape-only), so the
oracle is pinned against outputs of the reference itself: ``tests/golden/make_golden.py`` imports the unmodified
reference from ``/root/reference`` in the build container and writes inputs, injected noise, outputs and gradients
to ``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` replays them through this package.

Every function cites the reference file:line it follows (paths relative to ``/root/reference/torch_scae``).
"""

"""CPU oracle for the P-Companion hot path.  TEST INFRASTRUCTURE ONLY.

This package restates, in numpy (float64 by default), the arithmetic the reference
performs on the path BASELINE.json names.  It is imported only by ``tests/``, by
``__graft_entry__.smoke()`` and by ``bench.py``'s CPU-baseline / ``--impl reference`` legs.
Nothing under ``pcompanion_b200/`` imports it: the product path is CUDA-only and fails
loudly when the extension is missing.

Pinning: the reference ships no golden vectors and its own tests do not run
(SURVEY.md section 4), so the oracle is pinned against outputs of the *real* reference
modules imported in the build container (``tests/golden/make_golden.py`` -> the ``.npz``
fixtures in ``tests/golden/``; checked by ``tests/test_oracle_golden.py``).

Modules
-------
p2v        Product2Vec: FFN, BatchNorm, multi-head attention (dense and CSR form),
           analytic attention backward, triplet hinge.
bpg        Behaviour product graph: CSR/CSC construction, neighbour lookup, set algebra.
pcomp      Type transition, item prediction, PCompanion forward and joint hinge loss.
retrieval  Masked top-K retrieval (fp64 scores, stable sort, ties -> lowest index).
torch_port torch-CPU restatement of the reference modules; the CPU baseline that
           ``bench.py`` times (same ATen ops the reference issues).
"""

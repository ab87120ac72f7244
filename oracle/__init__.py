"""CPU oracle for the MetaBBO rollout hot path -- TEST INFRASTRUCTURE ONLY.

This package is a numpy restatement of the reference algorithms (GMC-DRL/MetaBox,
`src/problem`, `src/optimizer`, the rollout `act` of `src/agent`).  Each function cites
the reference file:line it follows.  It is the *checker* for the CUDA path:

* only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline / `--impl reference`
  legs may import it;
* nothing under `metabox_b200/` imports it, and the product has no CPU fallback.

Parity pin: the reference ships no tests / golden vectors, so the oracle is
pinned against outputs of the reference itself, imported unmodified in the build container by
`tests/golden/make_golden.py`; the resulting fixtures live in `tests/golden/*.npz` and are
checked by `tests/test_oracle_golden.py` (CPU, `-m "not gpu"`).
"""

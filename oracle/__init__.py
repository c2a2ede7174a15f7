"""CPU oracle of the voice-detector batch path.  TEST INFRASTRUCTURE ONLY.

Everything under `oracle/` restates the reference's algorithm for the path
(numpy for integer/byte work, torch-CPU for the float32 network) and exists to
CHECK the CUDA path.  Only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import it.  The
product (`softspoken_b200/`) never does, and fails loudly when its CUDA
library is missing.

Parity pinning: the reference ships no tests, golden vectors or fixtures
(SURVEY.md §8c), so the oracle is pinned against outputs of the reference's own
code imported in the build container (`oracle/ref_shim.py`), frozen by
`oracle/make_golden.py` into `tests/golden/` and re-checked by
`tests/test_oracle_*.py` on every run.
"""

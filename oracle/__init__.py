"""ORACLE — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement (plain PyTorch / numpy, fp32 or fp64) of the reference algorithm for the Multi-StyleGAN
training-step hot path.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` legs may import it, and only as the checker or the timed CPU baseline; nothing under
`multi_stylegan_b200/` imports it.

Pinning: the reference ships no tests, golden vectors or fixtures.  The oracle is
therefore pinned against *outputs of the reference itself*: `oracle/ref_loader.py` imports the
unmodified reference modules from /root/reference (CPU, with stand-ins only for the two pre-built CUDA
extension modules), `oracle/make_golden.py` dumps its outputs for seeded inputs into `tests/golden/`,
and `tests/test_oracle_golden.py` checks this restatement against those fixtures (and live against the
reference whenever /root/reference is present).  Two dependencies remain "parity unpinned" because
nothing in the reference pins them: ATen conv/linear/bmm numerics (torch 1.8.1 pinned by the reference,
2.11 here) and kornia 0.4.1's warp (not vendored, not installed) — see DESIGN.md.
"""

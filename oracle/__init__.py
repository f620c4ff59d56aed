"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatement of the semi-tts vector-quantisation bottleneck
(reference: src/embed.py:57-147, :150-205, :208-213 and the autograd backward
derived in SURVEY.md section 3.4).  Nothing in the product package
(`semi-tts_b200/`) may import this package.  The only permitted users are
`tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference`
legs of `bench.py`, and there only as the checker / the timed CPU arm.

Parity status: PINNED against the reference's own implementation.  The reference
ships no tests or golden vectors, so the pin is
`oracle/gen_golden.py`: it imports the unmodified reference modules from
/root/reference in the build container, runs them on seeded inputs and commits
inputs+outputs under tests/golden/.  `tests/test_oracle_golden.py` checks the
restatement against those vectors on every CPU run.

Exception (stated here and in DESIGN.md): the commitment / codebook loss terms
(`vq_loss`, `commit_loss` with non-zero weights) have NO reference arithmetic --
the reference asserts both weights are zero (src/embed.py:65-66) -- so for those
two scalars parity is UNPINNED; the restatement follows van den Oord et al. 2017.
"""

"""CPU oracle for the vispeech `SynthesizerTrn.infer` hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in `vispeech_b200/` may import this package;
only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` /
`--impl reference` legs use it, and there only as the checker / CPU baseline.

Parity status: PINNED against the unmodified reference.  `tests/golden/make_golden.py`
imports `/root/reference/models.py`, loads the state dict produced by
`oracle.weights.make_state_dict`, runs `SynthesizerTrn.infer` and commits the outputs
under `tests/golden/*.npz`; `tests/test_oracle_golden.py` checks this restatement
against those vectors (the reference itself ships no tests or golden vectors,
SURVEY.md section 4 / 8c).
"""

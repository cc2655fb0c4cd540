"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the Valle2 hot path.

Nothing in the product package (``valle2_b200``) may import this package. Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs use it, and only as the checker or the CPU baseline.

Contents
--------
``valle_oracle``  plain fp32 CPU restatement (torch CPU ops as the array library) of the
                  reference algorithm, each function citing the reference file:line.
``ref_shims``     import recipe that makes the real reference (``/root/reference``) importable
                  offline in the authoring container.
``make_golden``   runs the real reference through ``ref_shims`` and freezes golden vectors
                  under ``tests/golden/``; the oracle is pinned against those vectors by
                  ``tests/test_oracle_golden.py``.
``synth``         deterministic synthetic weights / inputs shared by the goldens and the tests.

Parity status: PINNED for a1-a13 (module stack, AR forward / generate, masks, sampling filter)
against outputs of the executed reference (tests/golden/*.npz). The NAR generate path is pinned
against the reference's own Transformer/embedding modules driven by the repaired stage loop
(SURVEY Appendix A-5..9), because upstream ``ValleNAR.generate`` raises.
"""

"""CPU oracle for AGRL's test-time hot path -- TEST INFRASTRUCTURE ONLY.

Nothing under ``agrl/`` imports this package.  Allowed importers: ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs.

Contents
    head.py       torch-fp32 (and fp64) restatement of the VMGN graph head,
                  reference torchreid/models/vmgn.py:104-172, :270-321
    distance.py   torch-fp32 restatement of torchreid/metrics/distance.py:59-89
    rank.py       ctypes front-end of rank_oracle.c (plain-C restatement of
                  rank_cy.pyx:154-249 and rank.py:160-212) and loader of oracle/_ref/rank_cy
                  (the reference's own Cython evaluator compiled by oracle/Makefile)
    synth.py      seeded synthetic inputs shared by tests, smoke and bench

Pinning: tests/test_oracle_*.py check every function here against tests/golden/*.npz, which
tests/golden/make_golden.py produced by importing the reference itself from /root/reference.
"""

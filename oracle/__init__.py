"""TEST INFRASTRUCTURE - NOT PRODUCT CODE.

CPU oracles (restatements of the reference's algorithms) used only by tests/,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs.
Nothing under ``rlipv2_b200/`` may import this package.
"""

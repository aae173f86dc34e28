"""TEST INFRASTRUCTURE ONLY — CPU/fp32 restatement ("oracle") of the TeReDiff hot path.

Nothing in ``tair_b200/`` imports this package.  It is the checker used by ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs.
See DESIGN.md ("Oracle") for how each module is pinned against the reference.
"""

"""TEST INFRASTRUCTURE ONLY — CPU restatements of the reference's hot path.

Nothing under `circuitvision_b200/` may import this package.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs use it,
and only as the checker or the reported CPU baseline.
"""

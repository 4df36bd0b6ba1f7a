"""ORACLE -- test infrastructure only.

CPU restatement of AD4SM.jl's `makeϕrKt` path (see the module headers for the reference
file:line each function follows).  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s
`cpu_baseline` / `--impl reference` legs may import or execute anything under `oracle/`; the
product (`ad4sm.jl_b200/`) never does.

PARITY UNPINNED BY REFERENCE TESTS: the reference has no tests or golden vectors and Julia is
not installed in the build container.  The oracle is pinned by analytic known-answer tests,
sympy Hessians and mpmath finite differences instead (tests/test_oracle_*.py).
"""

"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatements of the reference PointNet grasp-quality path
(/root/reference/PointNetGPD/model/pointnet.py:8-45,123-154,177-194).

Nothing under this directory is part of the product path.  Only `tests/`,
`__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of
`bench.py` may import it -- and there only as the checker or as the CPU arm
being timed, never as the thing shipped.  The product (`pointnetgpd_b200`)
never imports `oracle`.

Parity pinning: the reference ships NO tests or golden vectors for this path
(SURVEY.md section 4 / section 8c), so the oracle is pinned against outputs of the
reference module itself, imported from /root/reference in the build container
by `oracle/make_golden.py`; the resulting vectors live in `tests/golden/`.
"""

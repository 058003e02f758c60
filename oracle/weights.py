"""Deterministic generators for model state and inputs -- re-exported from the package.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

The generators themselves live in pointnetgpd_b200/synth.py (they are the product's synthetic-data source for
bench.py and smoke() as well); the oracle and the golden-vector script (oracle/make_golden.py) import them from
here so that fixtures, tests and benchmarks share ONE definition of "seed -> arrays".
"""
from pointnetgpd_b200.synth import (  # noqa: F401
    GRIPPER_W, _splitmix64, _stable_hash, _layout, uniform01, uniform, normal, state_keys, make_state, make_clouds,
    make_labels)

"""TEST INFRASTRUCTURE ONLY — the deterministic synthetic parameter / input generator lives in `maed_b200.synth` (bench.py and
the scripts need synthetic weights without importing anything under oracle/); the oracle and the tests keep this name."""
from maed_b200.synth import *  # noqa: F401,F403
from maed_b200.synth import _gen  # noqa: F401

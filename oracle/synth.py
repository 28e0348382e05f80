"""
ORACLE -- TEST INFRASTRUCTURE ONLY.  The synthetic model / signal generators live in `bonito_b200/synth.py` (bench.py's product arm
needs them and must not import from `oracle/`); this module re-exports them for the oracle, the golden-vector script and the tests.
"""
from bonito_b200.synth import *  # noqa: F401,F403
from bonito_b200.synth import _orthogonal_blocks  # noqa: F401

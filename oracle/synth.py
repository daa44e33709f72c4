"""The synthetic weight / input generators live in speech2lip_b200/synth.py (bench.py's product arm needs them and must
not import anything from oracle/); this shim keeps `from oracle import synth` working for the tests and the oracle."""
from speech2lip_b200.synth import *          # noqa: F401,F403
from speech2lip_b200.synth import _fan_in    # noqa: F401

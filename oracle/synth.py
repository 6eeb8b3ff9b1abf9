"""Deterministic synthetic weights / mixtures: re-export of speech_enhancement_mi_b200.synth (data generation only, no
algorithm of the path lives there) so that oracle/make_golden.py and the tests keep one import site."""
from speech_enhancement_mi_b200.synth import *  # noqa: F401,F403
from speech_enhancement_mi_b200.synth import (crn_param_shapes, make_crn_weights, make_mixture, normal,  # noqa: F401
                                               uniform01, with_alias_keys)

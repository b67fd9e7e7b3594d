"""maed_b200 — B200-native (sm_100a) implementation of MAED's per-clip forward hot path.

Public surface mirrors the reference's ``lib.models``:  ``from maed_b200.models import MAED``.
"""
__version__ = "0.1.0"

"""Drop-in for the reference's `lib/models/__init__.py` (two lines there: `from .maed import MAED`, `from .ops import *`).

Copy this file over `lib/models/__init__.py` of a ziniuwan/maed checkout that has `maed_b200` importable (the repo root on
PYTHONPATH): `train.py:27` / `eval.py:7` (`from lib.models import MAED`) then build the B200 module — same constructor,
`forward`, `extract_feature`, state_dict keys — and every other line of `train.py`, `eval.py`, `lib/core/trainer.py`,
`lib/core/evaluate.py` runs unchanged.  `tests/test_integration_shim.py` does exactly that with a copy of the reference tree
and walks the reference's own training iteration (its `get_optimizer`, its `Loss`, the video + image double forward).
"""
from maed_b200.models import MAED        # noqa: F401  (replaces `from .maed import MAED`)
from .ops import *                       # noqa: F401,F403  (unchanged re-export, reference lib/models/__init__.py:2)

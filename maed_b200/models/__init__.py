"""Drop-in for the reference's ``lib/models/__init__.py`` (``from .maed import MAED``)."""
from .maed import MAED  # noqa: F401

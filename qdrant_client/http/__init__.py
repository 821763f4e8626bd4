"""``qdrant_client.http`` namespace of the drop-in (models + exceptions)."""
from . import models  # noqa: F401
from . import exceptions  # noqa: F401

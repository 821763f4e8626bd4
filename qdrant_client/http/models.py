"""Alias of ``qdrant_client.models`` (the third-party package exposes both import paths)."""
from ..models import *  # noqa: F401,F403
from ..models import _Model  # noqa: F401

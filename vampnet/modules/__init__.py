"""vampnet.modules (reference vampnet/modules/__init__.py:6)."""
from . import activations, layers, transformer  # noqa: F401
from .transformer import VampNet  # noqa: F401

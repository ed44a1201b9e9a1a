"""``vampnet`` — the reference's import surface (reference vampnet/__init__.py:1-76), served by vampnet_b200.

Unmodified reference callers keep their import lines:

    from vampnet.interface import Interface, signal_concat      # app.py:16
    from vampnet import mask as pmask                           # app.py:17, experiment.py:12, train.py:23
    import vampnet; vampnet.interface.Interface.default()       # hello.py:2-6
    from vampnet.modules.transformer import VampNet             # train.py:20
    from vampnet.util import codebook_unflatten, codebook_flatten   # train.py:22

Put this repository on PYTHONPATH ahead of (or instead of) the reference checkout.  Everything here re-exports
vampnet_b200; nothing computes.  The hub download helpers resolve against the local cache only (there is no network
in this build): a hit returns the reference's paths, a miss raises.
"""
from pathlib import Path

from . import modules  # noqa: F401
from . import interface, mask, util  # noqa: F401
from .interface import Interface  # noqa: F401
from .modules.transformer import VampNet  # noqa: F401

__version__ = "0.0.1"

ROOT = Path(__file__).parent.parent
MODELS_DIR = Interface.models_dir()
DEFAULT_HF_MODEL_REPO = "hugggof/vampnet"  # reference DEFAULT_HF_MODEL_REPO:1 (informational: nothing is downloaded)


def _cached(*parts):
    return str(Interface._cached(*parts))


def download_codec():
    """vampnet/__init__.py:19-30, from the local cache ($VAMPNET_MODELS_DIR or ./models/vampnet)."""
    return _cached("codec.pth")


def download_default():
    """vampnet/__init__.py:33-46 -> (coarse path, c2f path)."""
    return _cached("coarse.pth"), _cached("c2f.pth")


def download_finetuned(name, repo_id=DEFAULT_HF_MODEL_REPO):
    """vampnet/__init__.py:49-59."""
    return _cached("loras", name, "coarse.pth"), _cached("loras", name, "c2f.pth")


def list_finetuned(repo_id=DEFAULT_HF_MODEL_REPO):
    """vampnet/__init__.py:61-76: names that have both coarse.pth and c2f.pth."""
    return [n for n in Interface.available_models() if n != "default"]

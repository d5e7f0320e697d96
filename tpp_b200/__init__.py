"""Import alias for the product package.

The package directory is named ``train-procgen-pytorch_b200`` (not a valid Python identifier), so this thin
package extends its ``__path__`` to that directory: ``import tpp_b200.common.storage`` resolves to
``train-procgen-pytorch_b200/common/storage.py``.
"""
import os as _os

_PKG_DIR = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "train-procgen-pytorch_b200")
__path__.append(_PKG_DIR)

from . import _lib  # noqa: E402,F401

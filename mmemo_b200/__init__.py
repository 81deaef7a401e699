"""Importable alias for the package directory ``multimodal-emotion-processing_b200/``.

The repository layout names the package after the reference
(``multimodal-emotion-processing_b200``); a hyphen is not a legal Python identifier, so this shim
points ``mmemo_b200``'s ``__path__`` at that directory and runs its ``__init__`` in this namespace.
``import mmemo_b200.realformer`` therefore loads
``multimodal-emotion-processing_b200/realformer.py``.
"""
import os as _os

_PKG_DIR = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                         "multimodal-emotion-processing_b200")
__path__ = [_PKG_DIR]
with open(_os.path.join(_PKG_DIR, "__init__.py"), "r", encoding="utf-8") as _fh:
    exec(compile(_fh.read(), _os.path.join(_PKG_DIR, "__init__.py"), "exec"))
del _fh

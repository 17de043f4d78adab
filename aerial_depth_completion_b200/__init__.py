"""Importable alias of the package directory ``aerial-depth-completion_b200/`` (a hyphen cannot be imported).

``import aerial_depth_completion_b200`` executes ``aerial-depth-completion_b200/__init__.py`` with this module's
``__path__`` pointing there, so sub-modules (``.model_zoo``, ``.metrics`` ...) resolve to the real package.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "aerial-depth-completion_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _f

"""Import shim: the package directory is ``nextsearch-api_b200/`` (hyphen, as the repo layout
prescribes), which Python cannot import by name.  ``import nsb200`` registers it as
``nextsearch_api_b200`` and re-exports its public names."""
import importlib.util
import os
import sys

_ROOT = os.path.dirname(os.path.abspath(__file__))
_PKG_DIR = os.path.join(_ROOT, "nextsearch-api_b200")
_NAME = "nextsearch_api_b200"

if _NAME not in sys.modules:
    _spec = importlib.util.spec_from_file_location(
        _NAME, os.path.join(_PKG_DIR, "__init__.py"), submodule_search_locations=[_PKG_DIR])
    _mod = importlib.util.module_from_spec(_spec)
    sys.modules[_NAME] = _mod
    _spec.loader.exec_module(_mod)

pkg = sys.modules[_NAME]
from nextsearch_api_b200 import *  # noqa: E402,F401,F403
from nextsearch_api_b200 import _lib  # noqa: E402,F401

"""Import shim: the product package directory is named after the reference
(`toypathtracer-games101-assignment7_b200/`), which is not a valid Python identifier.
`import tpt_b200` loads that directory as the module `tpt_b200`."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "toypathtracer-games101-assignment7_b200")
_spec = importlib.util.spec_from_file_location("tpt_b200", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["tpt_b200"] = _mod
_spec.loader.exec_module(_mod)

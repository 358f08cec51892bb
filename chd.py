"""Import alias: `import chd` loads the package that lives in `contact-human-dynamics_b200/`
(a hyphenated directory name cannot be imported directly)."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "contact-human-dynamics_b200")
_spec = importlib.util.spec_from_file_location("chd", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["chd"] = _mod
_spec.loader.exec_module(_mod)

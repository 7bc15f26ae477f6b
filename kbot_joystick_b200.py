"""Import shim: the package directory is `kbot-joystick_b200/` (hyphen, per the repo layout contract), which
Python cannot import by name.  `import kbot_joystick_b200` loads that directory as a regular package."""
import importlib.util
import sys
from pathlib import Path

_dir = Path(__file__).resolve().parent / "kbot-joystick_b200"
_spec = importlib.util.spec_from_file_location(
    "kbot_joystick_b200", _dir / "__init__.py", submodule_search_locations=[str(_dir)])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["kbot_joystick_b200"] = _mod
_spec.loader.exec_module(_mod)

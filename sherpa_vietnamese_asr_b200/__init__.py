"""Importable alias for the package directory `sherpa-vietnamese-asr_b200/` (a hyphen is not a valid
Python identifier, so this shim points the import system at it)."""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                          "sherpa-vietnamese-asr_b200")]
with open(_os.path.join(__path__[0], "__init__.py"), encoding="utf-8") as _f:
    exec(compile(_f.read(), _os.path.join(__path__[0], "__init__.py"), "exec"))

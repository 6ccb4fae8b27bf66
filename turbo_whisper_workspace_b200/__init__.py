"""Import shim: the product package lives in the directory ``turbo-whisper-workspace_b200/`` (the
name the build contract fixes), which is not a valid Python identifier.  Importing
``turbo_whisper_workspace_b200`` exposes that directory as this package."""
import os as _os

_impl = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "turbo-whisper-workspace_b200")
__path__ = [_impl]
with open(_os.path.join(_impl, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_impl, "__init__.py"), "exec"))

"""Import shim: `import ggufb200` loads the package that lives in `llama-gguf-inference_b200/`
(a directory name python cannot import directly)."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "llama-gguf-inference_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _f, _os

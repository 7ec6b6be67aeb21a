"""Import alias: `eccq_b200` is the package stored in ``quantized-kv-cache-ecc-protection_b200/`` (a directory name the
task fixes but Python cannot import).  Sub-packages resolve inside that directory."""
from pathlib import Path as _Path

_real = _Path(__file__).resolve().parent.parent / "quantized-kv-cache-ecc-protection_b200"
__path__ = [str(_real)]
exec(compile((_real / "__init__.py").read_text(), str(_real / "__init__.py"), "exec"))

"""Importable alias of the `lpd-net-pytorch_b200/` package directory (a hyphen cannot be imported)."""
from pathlib import Path as _Path

_real = _Path(__file__).resolve().parent.parent / "lpd-net-pytorch_b200"
__path__ = [str(_real)]
exec(compile((_real / "__init__.py").read_text(), str(_real / "__init__.py"), "exec"))

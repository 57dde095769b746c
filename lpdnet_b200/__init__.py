"""lpdnet_b200 — importable name of the `lpd-net-pytorch_b200/` package directory (a hyphen cannot be imported): the
sub-modules are found there through __path__; see lpd-net-pytorch_b200/__init__.py for the public surface."""
from pathlib import Path as _Path

__path__ = [str(_Path(__file__).resolve().parent.parent / "lpd-net-pytorch_b200")]
__version__ = "0.2.0"

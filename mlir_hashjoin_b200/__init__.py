"""Importable alias for the package directory ``mlir-hashjoin_b200/`` (a hyphen is not a legal module name).

``import mlir_hashjoin_b200`` executes ``mlir-hashjoin_b200/__init__.py`` with this module's ``__path__`` pointing
at that directory, so ``mlir_hashjoin_b200.join`` etc. resolve to the files that live there."""
from pathlib import Path as _Path

_real = _Path(__file__).resolve().parent.parent / "mlir-hashjoin_b200"
__path__ = [str(_real)]
exec(compile((_real / "__init__.py").read_text(), str(_real / "__init__.py"), "exec"))

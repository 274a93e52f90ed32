"""Import shim: the package directory is named ``voitta-rag_b200`` (repo contract), which is
not a valid Python identifier.  ``import voitta_rag_b200`` resolves to it."""
from pathlib import Path as _Path

_real = _Path(__file__).resolve().parent.parent / "voitta-rag_b200"
__path__ = [str(_real)]
exec(compile((_real / "__init__.py").read_text(), str(_real / "__init__.py"), "exec"))

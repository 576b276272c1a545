"""Importable alias of the product package, whose directory name `mammo-clip_b200/` is not a Python identifier."""
import os as _os

__path__.insert(0, _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "mammo-clip_b200"))

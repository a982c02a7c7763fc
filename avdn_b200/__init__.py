"""Importable alias of the product package.

The product lives in ``aerial-vision-and-dialog-navigation_b200/`` (the directory
name the project layout prescribes).  A hyphenated directory cannot be imported
by name, so this module makes ``import avdn_b200`` resolve to that directory:
``avdn_b200.env``, ``avdn_b200.models.ET_haa`` ... are the files in there.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "aerial-vision-and-dialog-navigation_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _os, _f

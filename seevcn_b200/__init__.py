"""Import alias: the product package lives in ``see-vcn_b200/`` (a directory name Python
cannot import because of the hyphen); ``import seevcn_b200`` resolves to it."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "see-vcn_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))

"""Import alias: the package lives in the directory `slam-localization_b200/` (the name the
project layout prescribes), which is not a valid Python identifier.  This module turns itself
into that package so `import slam_localization_b200.synth` etc. work."""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "slam-localization_b200")]
__file__ = _os.path.join(__path__[0], "__init__.py")
with open(__file__) as _f:
    exec(compile(_f.read(), __file__, "exec"))

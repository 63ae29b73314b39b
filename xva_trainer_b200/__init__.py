"""Import alias: the package sources live in ``xva-trainer_b200/`` (the directory name the project layout asks for,
which is not a valid Python identifier). This stub makes ``import xva_trainer_b200.<module>`` resolve there."""
import os as _os

__path__.append(_os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "xva-trainer_b200"))

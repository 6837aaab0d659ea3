from .base import LocoEnv, ValidTaskConf, ObservationType
from .unitree_a1 import UnitreeA1

UnitreeA1.register()

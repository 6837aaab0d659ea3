from .base import LocoEnv, ValidTaskConf, ObservationType
from .unitree_a1 import UnitreeA1

UnitreeA1.register()
from .robot_humanoids import Atlas, Talos

Atlas.register()
Talos.register()
from .humanoids import HumanoidTorque

HumanoidTorque.register()

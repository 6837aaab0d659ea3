from .base import LocoEnv, ValidTaskConf, ObservationType
from .unitree_a1 import UnitreeA1

UnitreeA1.register()
from .robot_humanoids import Atlas, Talos, UnitreeH1, UnitreeG1

Atlas.register()
Talos.register()
UnitreeH1.register()
UnitreeG1.register()
from .humanoids import HumanoidTorque, HumanoidTorque4Ages

HumanoidTorque.register()
HumanoidTorque4Ages.register()

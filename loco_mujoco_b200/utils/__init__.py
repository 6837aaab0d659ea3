from .reward import (RewardInterface, NoReward, PosReward, CustomReward, TargetVelocityReward,
                     MultiTargetVelocityReward, VelocityVectorReward)
from .math import (rotate_obs, mat2angle_xy, angle2mat_xy, transform_angle_2pi, euler_to_mat, mat_to_euler)
from .goals import GoalDirectionVelocity
from .checks import check_validity_task_mode_dataset
from ..trajectory import Trajectory

"""Goal container of the goal-conditioned envs (UnitreeA1: desired direction + speed).

Same public surface as the reference's helper (/root/reference/loco_mujoco/utils/goals.py): `set_goal`, `get_goal`,
`get_direction`, `get_velocity`, call = `get_goal`; values are handed out as copies."""
import copy


class GoalDirectionVelocity:
    __slots__ = ("_goal",)

    def __init__(self):
        self._goal = {}

    def set_goal(self, direction, velocity):
        self._goal = {"direction": direction, "velocity": velocity}

    def _get(self, key):
        if key not in self._goal or self._goal[key] is None:
            raise AssertionError("goal %s has not been set" % key)
        return copy.deepcopy(self._goal[key])

    def get_direction(self):
        return self._get("direction")

    def get_velocity(self):
        return self._get("velocity")

    def get_goal(self):
        return self.get_direction(), self.get_velocity()

    __call__ = get_goal

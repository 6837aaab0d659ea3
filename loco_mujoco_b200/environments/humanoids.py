"""
HumanoidTorque (tasks walk / run, real dataset) on the batched CUDA engine.
Mirrors /root/reference/loco_mujoco/environments/humanoids/base_humanoid.py:12-496 (XML surgery: joint / motor /
equality removal :86-127, box feet :435-472, arm re-orientation :474-496; observation/action specification :292-433;
_has_fallen :129-180; generate :211-290) and humanoids.py:7-317 (HumanoidTorque).
"""
import os
import warnings

import numpy as np

from .. import mjcf
from ..utils.checks import check_validity_task_mode_dataset
from .base import LocoEnv, ObservationType, ValidTaskConf, reference_data_root, ASSET_DIR

_ROOT = ["pelvis_tx", "pelvis_tz", "pelvis_ty", "pelvis_tilt", "pelvis_list", "pelvis_rotation"]
_LEG = ["hip_flexion", "hip_adduction", "hip_rotation", "knee_angle", "ankle_angle", "subtalar_angle", "mtp_angle"]
_LUMBAR = ["lumbar_extension", "lumbar_bending", "lumbar_rotation"]
_ARM = ["arm_flex", "arm_add", "arm_rot", "elbow_flex", "pro_sup", "wrist_flex", "wrist_dev"]
_ARM_MOT = ["shoulder_flex", "shoulder_add", "shoulder_rot", "elbow_flex", "pro_sup", "wrist_flex", "wrist_dev"]


class BaseHumanoid(LocoEnv):
    def _collision_groups_spec(self):
        if self._use_box_feet:      # base_humanoid.py:106-114
            return [("floor", ["floor"]), ("foot_r", ["foot_box_r"]), ("foot_l", ["foot_box_l"])]
        return [("floor", ["floor"]), ("foot_r", ["r_foot"]), ("front_foot_r", ["r_bofoot"]), ("foot_l", ["l_foot"]),
                ("front_foot_l", ["l_bofoot"])]

    def _get_grf_size(self):        # base_humanoid.py:182-191
        return 6 if self._use_box_feet else 12

    def _grf_group_names(self):     # base_humanoid.py:193-209
        return ["foot_r", "foot_l"] if self._use_box_feet else ["foot_r", "front_foot_r", "foot_l", "front_foot_l"]

    def __init__(self, use_muscles=False, use_box_feet=True, disable_arms=True, alpha_box_feet=0.5, **kwargs):
        if use_muscles:
            raise NotImplementedError("muscle actuation (tendons / <muscle>) is out of scope")
        self._use_muscles, self._use_box_feet, self._disable_arms = use_muscles, use_box_feet, disable_arms
        action_spec = self._get_action_specification(use_muscles)
        observation_spec = self._get_observation_specification()
        joints_to_remove, motors_to_remove, equ = self._get_xml_modifications()
        hide = ["q_" + j for j in joints_to_remove] + ["dq_" + j for j in joints_to_remove]
        observation_spec = [e for e in observation_spec if e[0] not in hide]
        action_spec = [a for a in action_spec if a not in motors_to_remove]
        if kwargs.get("compiled_model") is None:
            root = reference_data_root()
            if root is None:
                raise FileNotFoundError("loco_mujoco model data not found (set LOCO_MUJOCO_PATH)")
            h = mjcf.XmlHandle(os.path.join(root, "environments", "data", "humanoid", "humanoid_torque.xml"))
            if use_box_feet or disable_arms:
                h = self._delete_from_xml_handle(h, joints_to_remove, motors_to_remove, equ)
                if use_box_feet:
                    h = self._add_box_feet_to_xml_handle(h, alpha_box_feet)
                if disable_arms:
                    h = self._reorient_arms(h)
        else:
            h = None
        super().__init__(h, action_spec, observation_spec, self._collision_groups_spec(), **kwargs)

    def create_dataset(self, ignore_keys=None):
        if ignore_keys is None:
            ignore_keys = ["q_pelvis_tx", "q_pelvis_tz"]
        return super().create_dataset(ignore_keys)

    def _get_xml_modifications(self):
        joints, motors, equ = [], [], []
        if self._use_box_feet:
            joints += ["subtalar_angle_l", "mtp_angle_l", "subtalar_angle_r", "mtp_angle_r"]
            motors += ["mot_" + j for j in joints]
            equ += [j + "_constraint" for j in joints]
        if self._disable_arms:
            arms = ["%s_%s" % (j, s) for s in ("r", "l") for j in _ARM]
            joints += arms
            motors += ["mot_%s_%s" % (j, s) for s in ("r", "l") for j in _ARM_MOT]
            equ += ["wrist_flex_r_constraint", "wrist_dev_r_constraint", "wrist_flex_l_constraint",
                    "wrist_dev_l_constraint"]
        return joints, motors, equ

    def _has_fallen_terms(self):
        return [("q_pelvis_ty", -0.46, 0.1), ("q_pelvis_tilt", -np.pi / 4.5, np.pi / 12),
                ("q_pelvis_list", -np.pi / 12, np.pi / 8), ("q_pelvis_rotation", -np.pi / 9, np.pi / 9),
                ("q_lumbar_extension", -np.pi / 4, np.pi / 10), ("q_lumbar_bending", -np.pi / 10, np.pi / 10),
                ("q_lumbar_rotation", -np.pi / 4.5, np.pi / 4.5)]

    @staticmethod
    def _add_box_feet_to_xml_handle(h, alpha_box_feet, scaling=1.0):
        size = (np.array([0.112, 0.03, 0.05]) * scaling).tolist()
        pos = (np.array([-0.09, 0.019, 0.0]) * scaling).tolist()
        h.add(h.find("body", "toes_l"), "geom", name="foot_box_l", type="box", size=size, pos=pos, euler=[0.0, 0.15, 0.0])
        h.add(h.find("body", "toes_r"), "geom", name="foot_box_r", type="box", size=size, pos=pos, euler=[0.0, -0.15, 0.0])
        for g in ("r_foot", "r_bofoot", "l_foot", "l_bofoot"):
            e = h.find("geom", g)
            e.set("contype", "0")
            e.set("conaffinity", "0")
        return h

    @staticmethod
    def _reorient_arms(h):
        for name, quat in (("humerus_l", [1.0, -0.1, -1.0, -0.1]), ("ulna_l", [1.0, 0.6, 0.0, 0.0]),
                           ("humerus_r", [1.0, 0.1, 1.0, -0.1]), ("ulna_r", [1.0, -0.6, 0.0, 0.0])):
            h.find("body", name).set("quat", " ".join(repr(x) for x in quat))
        return h

    @staticmethod
    def _get_observation_specification():
        joints = _ROOT + ["%s_r" % j for j in _LEG] + ["%s_l" % j for j in _LEG] + _LUMBAR + \
                 ["%s_r" % j for j in _ARM] + ["%s_l" % j for j in _ARM]
        return [("q_" + j, j, ObservationType.JOINT_POS) for j in joints] + \
               [("dq_" + j, j, ObservationType.JOINT_VEL) for j in joints]

    @staticmethod
    def _get_action_specification(use_muscles=False):
        return ["mot_lumbar_ext", "mot_lumbar_bend", "mot_lumbar_rot"] + \
               ["mot_%s_%s" % (j, s) for s in ("r", "l") for j in _ARM_MOT] + \
               ["mot_%s_%s" % (j, s) for s in ("r", "l") for j in _LEG]

    @classmethod
    def _generate(cls, stubs, task="walk", dataset_type="real", debug=False, **kwargs):
        check_validity_task_mode_dataset(cls.__name__, task, None, dataset_type, *cls.valid_task_confs.get_all())
        if dataset_type != "real":
            raise NotImplementedError("perfect datasets are not shipped (network download in the reference)")
        reward_type = kwargs.pop("reward_type", "target_velocity")
        reward_params = kwargs.pop("reward_params", dict(target_velocity=1.25 if task == "walk" else 2.5))
        root = reference_data_root()
        if root is not None:
            mdp = cls(reward_type=reward_type, reward_params=reward_params, **kwargs)
            path = os.path.join(root, "datasets", "humanoids", "real", stubs[task])
            if debug or not os.path.exists(path):
                if not os.path.exists(path) and not debug:
                    warnings.warn("Datasets not found, falling back to test datasets. Please download and install "
                                  "the datasets to use this environment for imitation learning!")
                path = os.path.join(root, "datasets", "humanoids", "real", "mini_datasets", stubs[task])
            mdp.load_trajectory(dict(traj_path=path, traj_dt=1 / 500.0, control_dt=mdp.dt))
        else:
            from .. import modelpack
            asset = np.load(os.path.join(ASSET_DIR, "%s.%s.npz" % (cls.__name__, task)), allow_pickle=False)
            model = modelpack.from_npz_dict({k[6:]: asset[k] for k in asset.files if k.startswith("model_")})
            mdp = cls(reward_type=reward_type, reward_params=reward_params, compiled_model=model, **kwargs)
            mdp.load_trajectory(dict(processed={k[5:]: asset[k] for k in asset.files if k.startswith("traj_")}))
        return mdp


class HumanoidTorque(BaseHumanoid):
    valid_task_confs = ValidTaskConf(tasks=["walk", "run"], data_types=["real", "perfect"])

    def __init__(self, **kwargs):
        if "use_muscles" in kwargs:
            assert kwargs.pop("use_muscles") is False, "Activating muscles in this environment not allowed. "
        super().__init__(use_muscles=False, **kwargs)

    @staticmethod
    def generate(task="walk", dataset_type="real", **kwargs):
        return HumanoidTorque._generate({"walk": "02-constspeed_reduced_humanoid.npz",
                                         "run": "05-run_reduced_humanoid.npz"}, task, dataset_type, **kwargs)

"""
Generate the bundled task assets (compiled ModelPack + interpolated reset trajectories) from a loco_mujoco
checkout (default: /root/reference/loco_mujoco). The GPU box has no reference tree, so tests / bench there use
these files. Re-run after changing mjcf.py or the task definitions:   python tools/build_assets.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from loco_mujoco_b200 import LocoEnv, modelpack                                   # noqa: E402
from loco_mujoco_b200.environments.base import processed_trajectory_dict, ASSET_DIR  # noqa: E402

TASKS = sys.argv[1:] or ["UnitreeA1.simple", "UnitreeA1.hard"]
for task in TASKS:
    env = LocoEnv.make(task + ".real", debug=True)
    if len(env._models) > 1:        # multi-model env (carry): every model, prefix model<i>_
        d = {"n_models": np.int32(len(env._models))}
        for i, m in enumerate(env._models):
            d.update({"model%d_%s" % (i, k): v for k, v in modelpack.to_npz_dict(m).items()})
    else:
        d = {"model_" + k: v for k, v in modelpack.to_npz_dict(env._model).items()}
    d.update({"traj_" + k: v for k, v in processed_trajectory_dict(env.trajectories).items()})
    path = os.path.join(ASSET_DIR, task + ".npz")
    np.savez_compressed(path, **d)
    print(task, "->", path, os.path.getsize(path), "bytes")

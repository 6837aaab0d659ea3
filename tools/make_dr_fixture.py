"""
Golden fixture for the pooled domain-randomisation test (tests/test_domain_randomization.py): K randomised
recompilations of Atlas (walk config) from the reference MJCF with the config dict below, seeded.  Stores the pool
rows and, per row, the packed ModelPack reals so that the GPU box (no reference checkout) can rebuild the oracle.
Run here:  python tools/make_dr_fixture.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_domain_randomization import CONF          # noqa: E402
from loco_mujoco_b200 import LocoEnv, modelpack      # noqa: E402

np.random.seed(3)
env = LocoEnv.make("Atlas.walk.real", debug=True, domain_randomization_config=CONF, domain_randomization_pool_size=6)
pool = env.domain_randomization_pool()
packs = [modelpack.pack(m) for m in env._domain_rand.models]
base = modelpack.pack(env._model)
assert all(np.array_equal(p[0], base[0]) for p in packs)
out = os.path.join(ROOT, "tests", "golden", "dr_atlas_pool.npz")
np.savez_compressed(out, pool=pool, model_ints=base[0], model_reals=np.stack([p[1] for p in packs]))
print(out, pool.shape, os.path.getsize(out))

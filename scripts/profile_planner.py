import cProfile, pstats, io, math, random, sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from torque_constrained_motion_planning_b200 import collision, ikfast_panda_arm as ik, ik_utils, panda_primitives as pp, utils
Q_HOME = [0, -math.pi / 4, 0.0, -3 * math.pi / 4, 0, math.pi / 2, math.pi / 4]
pos8, rot8 = ik.get_fk([0.7, 0.3, 0.2, -1.9, 0.1, 2.2, 1.0])
c, s = math.cos(-math.pi / 4), math.sin(-math.pi / 4)
Rt = np.array(rot8) @ np.array([[c, -s, 0], [s, c, 0], [0, 0, 1.0]])
pose = (tuple(np.array(pos8) + Rt @ np.array([0, 0, 0.105])), tuple(ik_utils.quat_from_matrix(Rt)))
scene = collision.cluttered_scene()
def run(batch=0):
    random.seed(3); np.random.seed(3)
    p = utils.Problem(None, scene, "coke", 5.0, 5, "rne")
    return pp.planner_fn_force_aware(tuple(Q_HOME), pose, p, batch=batch)
run(); run(32)
for b in (0, 32):
    pr = cProfile.Profile(); pr.enable(); run(b); pr.disable()
    st = io.StringIO(); pstats.Stats(pr, stream=st).sort_stats("cumulative").print_stats(18); print("batch", b); print(st.getvalue()[:3500])

"""Drop-in for the reference's ``franka_ik_fast.py`` front-end (franka_ik_fast.py:19-79): ``PANDA_INFO``,
``get_ik_generator``, ``sample_tool_ik``, ``bi_panda_inverse_kinematics``.  Poses are those of the
``panda_grasptarget`` tool frame in the robot base frame; the tool -> link8 offset the reference reads from
PyBullet (get_tool_from_ik, :30-34) is the URDF constant (panda_mod.urdf:7-11,87-91)."""
from __future__ import annotations

from .ik_utils import PANDA_INFO, IKFastInfo, USE_ALL, USE_CURRENT  # noqa: F401
from .ikfast import ikfast_inverse_kinematics, is_ik_compiled
from .panda_model import Q_LOWER, Q_UPPER, TOP_HOLDING_LEFT_ARM
from .panda_primitives import bi_panda_inverse_kinematics, tool_pose_to_link8  # noqa: F401

PANDA_LEFT_INFO = IKFastInfo(module_name="ikfast_panda_arm", base_link="l_panda_link0", ee_link="l_panda_link8",
                             free_joints=["l_panda_joint7"])
PANDA_RIGHT_INFO = PANDA_INFO
info = {"left": PANDA_LEFT_INFO, "right": PANDA_RIGHT_INFO}
FRANKA_URDF = "models/panda_mod.urdf"


def get_tool_from_ik(robot=None, arm="right"):
    """franka_ik_fast.py:30-34: pose of the IK frame (``panda_grasptarget``) in the tool frame; both names denote
    the same link in this robot (utils.py PANDA_TOOL_FRAME, IK_FRAME['right']), so it is the identity."""
    return (0.0, 0.0, 0.0), (0.0, 0.0, 0.0, 1.0)


def get_joint_distances(current_config, new_config):
    """Mean squared joint difference (franka_ik_fast.py:39-44)."""
    n = len(new_config)
    return sum((new_config[i] - current_config[i]) ** 2 / n for i in range(n))


def get_ik_generator(robot, arm, gripper_link, gripper_pose, max_attempts=25, max_time=1.3, current_conf=None):
    """franka_ik_fast.py:36-37 (the reference hard-codes 25 attempts / 1.3 s regardless of the arguments)."""
    return ikfast_inverse_kinematics(robot, info[arm], gripper_link, tool_pose_to_link8(gripper_pose),
                                     max_attempts=25, max_time=1.3, current_conf=current_conf)


def sample_tool_ik(robot, arm, tool_pose, nearby_conf=USE_CURRENT, max_attempts=25, custom_limits={},
                   current_conf=None, **kwargs):
    """First configuration of the sweep that lies inside the (custom) joint limits, else None (:46-62)."""
    lower, upper = Q_LOWER.copy(), Q_UPPER.copy()
    for j, (lo, hi) in custom_limits.items():
        lower[j], upper[j] = lo, hi
    generator = get_ik_generator(robot, arm, None, tool_pose, current_conf=current_conf, **kwargs)
    for _ in range(max_attempts):
        try:
            conf = next(generator)
        except StopIteration:
            break
        if conf and all(lo <= v <= hi for lo, v, hi in zip(lower, conf, upper)):
            return conf
    return None

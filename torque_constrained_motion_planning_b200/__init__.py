"""B200-native torque-feasibility engine for the inner loop of
HIRO-group/torque_constrained_motion_planning.

Modules mirror the reference's names for the hot path -- ``rne``, ``panda_model``, ``ik_utils``,
``ikfast_panda_arm``, ``min_jerk_v2``, ``rrt_star``, ``panda_primitives``, ``utils`` (subset) -- and all
torque / IK arithmetic runs in ``libtcmp.so`` (hand-written CUDA for sm_100a behind the C-ABI of
``include/tcmp.h``).  There is no CPU fallback.
"""
__version__ = "0.1.0"

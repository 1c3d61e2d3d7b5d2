"""Drop-in for the reference's ``rne.py`` public surface, backed by libtcmp.so.

Reference contract kept (rne.py:173-198): ``rne(q, qd, qdd) -> np.ndarray(7)``,
``add_payload(r, m)``, ``remove_payload()``, ``get_has_payload()``, with the payload held in module
state between calls exactly as the reference's closures expect (panda_primitives.py:178-190).
``add_payload`` ignores ``r`` and attaches the payload iff ``m > 0`` like the reference (rne.py:181-188).

New: ``rne_batch`` evaluates many states per call ([7][n] arrays, CUDA tensors or NumPy) and takes
the payload as an argument instead of module state, so it is re-entrant.

The reference keeps its inertial tables in the module-level lists ``ms`` / ``cs`` / ``inertia_matrices``
(rne.py:102,119,138), which a user can edit in place.  The equivalent here is ``set_inertial_model(model)`` with an
``engine.InertialModel`` (tcmp_model); ``get_ms_global()`` reports the masses in the reference's 10-entry layout.

Every call runs on the GPU; a scalar ``rne()`` is a batch of one.
"""
from __future__ import annotations

import numpy as np

from . import engine
from ._lib import PAYLOAD_THRESHOLD_RAW

_has_payload = False
_payload_mass = 0.0
_model = None   # engine.InertialModel, or None = the compiled-in Panda


def set_inertial_model(model) -> None:
    """Use another inertial set (engine.InertialModel) for rne() / rne_batch(); None restores the stock Panda."""
    global _model
    _model = model


def get_inertial_model():
    return _model


def get_ms_global() -> list:
    """rne.py:154-156: link1..7, link8, hand, payload link (0.0 while no payload is attached)."""
    mass = (_model or engine.InertialModel.default()).mass
    return [float(m) for m in mass] + [_payload_mass if _has_payload else 0.0]


def get_cs_global() -> list:
    """rne.py:162-164: centres of mass, 10 entries (the payload link's stays at the origin, rne.py:181-188)."""
    com = (_model or engine.InertialModel.default()).com
    return [np.array(c) for c in com] + [np.zeros(3)]


def get_inertia_matricies() -> list:
    """rne.py:142-144 (sic): 3x3 inertias about the centres of mass; the payload's is appended while one is
    attached, as add_payload does (rne.py:186-187)."""
    mdl = _model or engine.InertialModel.default()
    out = [np.array([[i[0], i[1], i[2]], [i[1], i[3], i[4]], [i[2], i[4], i[5]]]) for i in mdl.inertia]
    if _has_payload:
        r2 = mdl.payload_radius ** 2
        out.append(np.diag([_payload_mass * r2, _payload_mass * r2, 0.0]))
    return out


def get_has_payload() -> bool:
    return _has_payload


def set_has_payload(val: bool) -> None:
    global _has_payload
    _has_payload = bool(val)


def add_payload(r, m) -> None:
    """rne.py:181-188: replaces any previous payload; only ``m > 0`` attaches one; ``r`` is unused."""
    global _has_payload, _payload_mass
    remove_payload()
    if m > 0:
        _has_payload = True
        _payload_mass = float(m)


def remove_payload() -> None:
    global _has_payload, _payload_mass
    _has_payload = False
    _payload_mass = 0.0


def rne(q, qd, qdd) -> np.ndarray:
    """Joint torques (7,) of the Panda + hand (+ current payload) -- rne.py:198-254."""
    col = lambda v: np.asarray(v, dtype=np.float64).reshape(-1)[:7].reshape(7, 1)
    tau, _ = engine.torque_test_batch(col(q), col(qd), col(qdd), _payload_mass if _has_payload else 0.0,
                                      mode="rne", payload_threshold=PAYLOAD_THRESHOLD_RAW, want_mask=False,
                                      model=_model)
    return tau[:, 0].copy()


def rne_batch(q, qd=None, qdd=None, payload_mass=0.0, dtype="f64", model=None):
    """Torques [7][n] for states [7][n]; payload attached iff mass > 0 (the raw rne.py rule)."""
    tau, _ = engine.torque_test_batch(q, qd, qdd, payload_mass, mode="rne", dtype=dtype,
                                      payload_threshold=PAYLOAD_THRESHOLD_RAW, want_mask=False,
                                      model=model if model is not None else _model)
    return tau

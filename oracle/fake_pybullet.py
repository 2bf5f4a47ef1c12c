"""Stand-in for the `pybullet` / `pybullet_data` / `gymnasium` modules -- TEST INFRASTRUCTURE.

PyBullet and Gymnasium are not installable in this image.  This module lets the
reference's own, unmodified class `/root/reference/env/enhanced_rocket_tvc_env.py:
EnhancedRocketTVCEnv` run: every PyBullet call it makes (ref:327-352, 392, 417-458, 477,
522-527, 546-556, 563-585, 589-590, 610-614, 725-728, 752) is forwarded to the oracle's physics
layer (oracle/tvc_oracle.c, rows B1-B9), so everything *outside* Bullet -- reward, phases,
success detector, termination, info -- is the reference's code, not a restatement.
tests/golden/make_golden.py uses it to freeze the golden trajectories.

Use:  install() registers the fake modules in sys.modules; load_reference_env() imports the
reference env module from /root/reference (only available in the build container).
"""
from __future__ import annotations

import ctypes as C
import importlib.util
import sys
import types

import numpy as np

from . import oracle as O

DIRECT, GUI = 2, 1
WORLD_FRAME, LINK_FRAME = 2, 1
GEOM_CYLINDER = 4


class _World:
    def __init__(self):
        self.reset()

    def reset(self):
        self.params = O.BodyParams()
        O.lib().orc_body_params_default(C.byref(self.params))
        self.params.gravity[2] = 0.0       # pybullet default: no gravity until setGravity
        self.params.substeps = 1
        self.params.dt_step = 1.0 / 240.0
        self.params.lin_damp = 0.04        # pybullet defaults before changeDynamics
        self.params.ang_damp = 0.04
        self.params.ground = 0
        self.bodies = {}
        self.next_id = 0
        self.plane_friction = 1.0
        self.body_friction = 0.5
        self.spin = [0.0, 0.0]
        self.roll = [0.0, 0.0]
        self.body_restitution = 0.0
        self.n_calls = 0


_W = _World()
_connected = [False]


def _tick():
    _W.n_calls += 1


def connect(mode, *a, **k):
    _connected[0] = True
    _W.reset()          # a new DIRECT connection is a new, empty physics server
    return 0


def disconnect(*a, **k):
    _connected[0] = False


def setAdditionalSearchPath(path):
    _tick()


def setGravity(x, y, z, *a, **k):
    _tick()
    _W.params.gravity[0], _W.params.gravity[1], _W.params.gravity[2] = x, y, z


def setPhysicsEngineParameter(fixedTimeStep=None, numSubSteps=None, enableConeFriction=None,
                              contactBreakingThreshold=None, enableFileCaching=None, **k):
    _tick()
    if fixedTimeStep is not None:
        _W.params.dt_step = fixedTimeStep
    if numSubSteps is not None:
        _W.params.substeps = max(1, int(numSubSteps))


def resetSimulation(*a, **k):
    _tick()
    _W.reset()


def loadURDF(name, *a, **k):
    _tick()
    assert name == "plane.urdf", name
    _W.params.ground = 1
    i = _W.next_id
    _W.next_id += 1
    _W.bodies[i] = "plane"
    return i


def createCollisionShape(shapeType, radius=0.5, height=1.0, **k):
    _tick()
    assert shapeType == GEOM_CYLINDER
    _W.params.radius = radius
    _W.params.half_len = 0.5 * height
    return 0


def createVisualShape(*a, **k):
    _tick()
    return 0


def createMultiBody(baseMass, baseCollisionShapeIndex, baseVisualShapeIndex, basePosition, baseOrientation,
                    baseInertialFramePosition=None, baseInertialFrameOrientation=None, **k):
    _tick()
    _W.params.mass = baseMass
    b = O.Body()
    O.lib().orc_body_init(C.byref(b), O._d(basePosition), O._d(baseOrientation))
    i = _W.next_id
    _W.next_id += 1
    _W.bodies[i] = b
    return i


def _combine():
    p = _W.params
    p.mu = _W.body_friction * _W.plane_friction
    p.mu_spin = _W.spin[0] * _W.plane_friction + _W.spin[1] * _W.body_friction
    p.mu_roll = _W.roll[0] * _W.plane_friction + _W.roll[1] * _W.body_friction
    p.restitution = _W.body_restitution


def changeDynamics(bodyId, linkIndex, localInertiaDiagonal=None, linearDamping=None, angularDamping=None,
                   restitution=None, lateralFriction=None, spinningFriction=None, rollingFriction=None, **k):
    _tick()
    tgt = _W.bodies[bodyId]
    is_plane = isinstance(tgt, str)
    if localInertiaDiagonal is not None:
        for i in range(3):
            _W.params.inertia[i] = localInertiaDiagonal[i]
    if linearDamping is not None:
        _W.params.lin_damp = linearDamping
    if angularDamping is not None:
        _W.params.ang_damp = angularDamping
    if restitution is not None and not is_plane:
        _W.body_restitution = restitution
    if lateralFriction is not None:
        if is_plane:
            _W.plane_friction = lateralFriction
        else:
            _W.body_friction = lateralFriction
    if spinningFriction is not None:
        _W.spin[1 if is_plane else 0] = spinningFriction
    if rollingFriction is not None:
        _W.roll[1 if is_plane else 0] = rollingFriction
    _combine()


def getDynamicsInfo(bodyId, linkIndex):
    _tick()
    p = _W.params
    return (p.mass, _W.body_friction, tuple(p.inertia))


def getBasePositionAndOrientation(bodyId):
    _tick()
    b = _W.bodies[bodyId]
    return tuple(b.pos), O.reported_quat(tuple(b.quat))


def getBaseVelocity(bodyId):
    _tick()
    b = _W.bodies[bodyId]
    return tuple(b.vel), tuple(b.omega)


def getMatrixFromQuaternion(q):
    _tick()
    return O.matrix_from_quat(q)


def getEulerFromQuaternion(q):
    _tick()
    return O.euler_from_quat(q)


def applyExternalForce(objectUniqueId, linkIndex, forceObj, posObj, flags):
    _tick()
    assert flags == WORLD_FRAME and linkIndex == -1
    O.lib().orc_apply_external_force(C.byref(_W.bodies[objectUniqueId]), O._d(forceObj), O._d(posObj))


def applyExternalTorque(objectUniqueId, linkIndex, torqueObj, flags):
    _tick()
    assert flags == WORLD_FRAME and linkIndex == -1
    O.lib().orc_apply_external_torque(C.byref(_W.bodies[objectUniqueId]), O._d(torqueObj))


def stepSimulation(*a, **k):
    _tick()
    for b in _W.bodies.values():
        if not isinstance(b, str):
            O.lib().orc_step_simulation(C.byref(_W.params), C.byref(b), None)


def world():
    return _W


# ---------------------------------------------------------------------------------------------
# minimal gymnasium stand-in (reset(seed) seeds np_random exactly like gymnasium.Env.reset)
# ---------------------------------------------------------------------------------------------
class _Box:
    def __init__(self, low, high, shape=None, dtype=np.float32):
        if shape is None:
            shape = np.shape(low)
        self.low = np.broadcast_to(np.asarray(low, dtype), shape).copy()
        self.high = np.broadcast_to(np.asarray(high, dtype), shape).copy()
        self.shape = tuple(shape)
        self.dtype = np.dtype(dtype)
        self._rng = np.random.default_rng()

    def seed(self, seed=None):
        self._rng = np.random.default_rng(seed)

    def sample(self):
        return self._rng.uniform(self.low, self.high).astype(self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

    def __repr__(self):
        return f"Box({self.low.min()}, {self.high.max()}, {self.shape}, {self.dtype})"


class _Env:
    metadata = {}
    np_random = None

    def reset(self, *, seed=None, options=None):
        if seed is not None or self.np_random is None:
            self.np_random = np.random.default_rng(seed)

    def close(self):
        pass


def install():
    """Register fake `pybullet`, `pybullet_data`, `gymnasium` in sys.modules (idempotent)."""
    me = sys.modules[__name__]
    sys.modules.setdefault("pybullet", me)
    pd = types.ModuleType("pybullet_data")
    pd.getDataPath = lambda: "/nonexistent/pybullet_data"
    sys.modules.setdefault("pybullet_data", pd)
    if "gymnasium" not in sys.modules:
        g = types.ModuleType("gymnasium")
        g.Env = _Env
        sp = types.ModuleType("gymnasium.spaces")
        sp.Box = _Box
        g.spaces = sp
        reg = types.ModuleType("gymnasium.envs.registration")
        reg.register = lambda **k: None
        envs = types.ModuleType("gymnasium.envs")
        envs.registration = reg
        g.envs = envs
        g.register = reg.register
        sys.modules["gymnasium"] = g
        sys.modules["gymnasium.spaces"] = sp
        sys.modules["gymnasium.envs"] = envs
        sys.modules["gymnasium.envs.registration"] = reg


def load_reference_env(path="/root/reference/env/enhanced_rocket_tvc_env.py"):
    """Import the reference env module (unmodified source) with the fake modules installed."""
    install()
    spec = importlib.util.spec_from_file_location("ref_enhanced_rocket_tvc_env", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod

"""Drop-in boundary on the CPU box (no GPU, no engine): `install_as_reference_env()` makes the import paths of the reference
trainer (scripts/train.py:44, env/__init__.py:105-111) resolve to the facade, and the facade's public API equals the
reference's, as frozen in tests/golden/reference_api.json by tests/golden/make_reference_api_snapshot.py (ast-parsed from
/root/reference/env/enhanced_rocket_tvc_env.py:21-29, :279-288, :381, :466 and env/__init__.py:28-111)."""
import importlib
import inspect
import json
import os
import sys

import pytest


@pytest.fixture()
def ref_api(golden_dir):
    return json.load(open(os.path.join(golden_dir, "reference_api.json")))


@pytest.fixture()
def installed():
    saved = {k: sys.modules.get(k) for k in ("env", "env.enhanced_rocket_tvc_env")}
    from tvc_ai_b200 import env as facade
    pkg = facade.install_as_reference_env()
    yield facade, pkg
    for k, v in saved.items():
        if v is None:
            sys.modules.pop(k, None)
        else:
            sys.modules[k] = v


def _sig(fn):
    out = []
    for name, p in inspect.signature(fn).parameters.items():
        if p.kind in (p.KEYWORD_ONLY, p.VAR_KEYWORD, p.VAR_POSITIONAL):
            continue                                  # the facade's extra knobs (device=, contract=, **engine overrides) are keyword-only
        out.append([name, None if p.default is inspect.Parameter.empty else repr(p.default)])
    return out


def test_reference_import_paths_resolve_to_the_facade(installed, ref_api):
    facade, pkg = installed
    # scripts/train.py:44  `from env.enhanced_rocket_tvc_env import EnhancedRocketTVCEnv, MissionPhase`
    mod = importlib.import_module("env.enhanced_rocket_tvc_env")
    assert mod.EnhancedRocketTVCEnv is facade.EnhancedRocketTVCEnv and mod.MissionPhase is facade.MissionPhase
    # env/__init__.py:105-111 `__all__`
    top = importlib.import_module("env")
    for name in ref_api["__all__"]:
        assert getattr(top, name) is getattr(facade, name), name
    from env import make_training_env, make_evaluation_env, make_debug_env  # noqa: F401


def test_constructor_and_method_signatures_equal_the_reference(installed, ref_api):
    facade, _ = installed
    cls = facade.EnhancedRocketTVCEnv
    for meth, want in ref_api["EnhancedRocketTVCEnv"].items():
        assert _sig(getattr(cls, meth)) == want, (meth, _sig(getattr(cls, meth)), want)
    assert cls.metadata == ref_api["metadata"]
    assert [[m.name, m.value] for m in facade.MissionPhase] == ref_api["MissionPhase"]


def test_factories_forward_the_reference_defaults(installed, ref_api, monkeypatch):
    facade, _ = installed
    seen = {}

    class Recorder:
        def __init__(self, **kw):
            seen.update(kw)

    monkeypatch.setattr(facade, "EnhancedRocketTVCEnv", Recorder)
    for name, spec in ref_api["factories"].items():
        fn = getattr(facade, name)
        assert _sig(fn) == spec["signature"] and spec["vararg_kwargs"] == "kwargs"
        seen.clear()
        cfg = {"reward_function": {"gradient_penalty": 0.2}}
        fn(config=cfg)
        assert seen.pop("config") is cfg and seen == spec["default_kwargs"], (name, seen)
        seen.clear()
        fn(max_episode_steps=7, debug=True)                    # caller kwargs override the defaults (dict.update)
        assert seen["max_episode_steps"] == 7 and seen["debug"] is True and seen["config"] is None


def test_registered_ids_with_a_stand_in_registry(installed, ref_api, monkeypatch):
    """env/__init__.py:28-64: three ids, max_episode_steps=1000 each (quirk Q22), the reference's kwargs.  Gymnasium is not
    installed in this image, so a stand-in `gymnasium.envs.registration` receives the calls."""
    facade, _ = installed
    import types
    calls = {}
    reg = types.ModuleType("gymnasium.envs.registration")
    reg.registry = {}

    def register(id, entry_point, max_episode_steps=None, kwargs=None, **rest):  # noqa: A002
        calls[id] = dict(entry_point=entry_point, max_episode_steps=max_episode_steps, kwargs=kwargs)
        reg.registry[id] = calls[id]

    reg.register = register
    g, ge = types.ModuleType("gymnasium"), types.ModuleType("gymnasium.envs")
    g.envs, ge.registration = ge, reg
    for k, v in (("gymnasium", g), ("gymnasium.envs", ge), ("gymnasium.envs.registration", reg)):
        monkeypatch.setitem(sys.modules, k, v)
    monkeypatch.setattr(facade.spaces, "HAVE_GYMNASIUM", True)
    assert facade.register_gym_ids() is True
    assert set(calls) == set(ref_api["registered"])
    for env_id, want in ref_api["registered"].items():
        got = calls[env_id]
        assert got["max_episode_steps"] == want["max_episode_steps"] == 1000
        assert got["kwargs"] == want["kwargs"], env_id
        modname, clsname = got["entry_point"].split(":")
        assert clsname == want["entry_point"].split(":")[1]
        assert getattr(importlib.import_module(modname), clsname) is facade.EnhancedRocketTVCEnv
    calls.clear()
    assert facade.register_gym_ids() is True and calls == {}      # idempotent: ids already in the registry are left alone


def test_spaces_match_the_reference_declaration():
    """observation_space Box(10,) float32 / action_space Box(2,) in [-1, 1] (enhanced_rocket_tvc_env.py:354-379)."""
    import numpy as np
    from tvc_ai_b200 import spaces
    o, a = spaces.observation_space(), spaces.action_space()
    assert o.shape == (10,) and o.dtype == np.float32 and a.shape == (2,) and a.dtype == np.float32
    assert np.all(a.low == -1.0) and np.all(a.high == 1.0)
    s = a.sample()
    assert s.shape == (2,) and np.all(np.abs(s) <= 1.0)

"""Quirk switches (SURVEY.md section 8(a) quirk index, include/tvc_b200.h TVC_Q_*), CPU side: clearing one bit in
the oracle changes exactly the output the quirk is about and leaves the rest of the scenario alone.  The GPU side
(tests/test_gpu_parity.py::test_quirk_switches_device_equals_oracle) checks that the device follows the oracle with
every single bit cleared.  No GPU here."""
import ctypes as C
import math

import numpy as np
import pytest


def _sim(O, clear=0, **over):
    return O.OracleSim(O.default_config(O.CONTRACT_R, quirks=O.Q_ALL_REFERENCE & ~clear, diversity_mode=O.DIV_FAST, **over), 1)


def _run(sim, actions):
    outs = []
    for a in actions:
        _, r, term, trunc, o = sim.step(np.asarray([a], np.float32))
        e = sim.env(0)
        outs.append(dict(r=float(r[0]), term=bool(term[0]), trunc=bool(trunc[0]), comp=list(o[0].comp), tilt=o[0].tilt,
                         alt=o[0].altitude, phase=o[0].phase, success=o[0].success, obs=np.frombuffer(o[0].obs, np.float32).copy(),
                         pos=list(e.body.pos), vel=list(e.body.vel), omega=list(e.body.omega), quat=list(e.body.quat)))
    return outs


ZERO = [(0.0, 0.0)]


def test_bits_match_the_c_header():
    import re, os
    from oracle import oracle as O
    hdr = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "tvc_b200.h")).read()
    bits = dict(re.findall(r"#define TVC_Q_(\w+)\s+\(1u << (\d+)\)", hdr))
    assert len(bits) == 14
    for name, sh in bits.items():
        assert getattr(O, "Q_" + name) == 1 << int(sh), name
    assert O.Q_ALL_REFERENCE == 0x3FFF and "#define TVC_Q_ALL_REFERENCE    0x3FFFu" in hdr


def test_q1_double_gravity(oracle_mod):
    O = oracle_mod
    a, b = _run(_sim(O), ZERO)[0], _run(_sim(O, O.Q_DOUBLE_GRAVITY), ZERO)[0]
    assert a["vel"][2] < 0 < b["vel"][2]                      # 17.5 - 19.62 < 0 < 17.5 - 9.81 (m/s^2)
    assert abs((b["vel"][2] - a["vel"][2]) - 9.81 * 0.02) < 1e-3
    assert a["quat"] == b["quat"] and a["omega"] == b["omega"]


def test_q2_thrust_vector_norm(oracle_mod):
    O = oracle_mod
    # zero action: [0, 0, 1] is already unit norm -> identical trajectories, bit for bit
    assert _run(_sim(O), ZERO * 5)[-1]["pos"] == _run(_sim(O, O.Q_THRUST_VECTOR), ZERO * 5)[-1]["pos"]
    # full deflection, gravity off the books: |dv| / dt of the first step is |F| / m
    g = np.array([0, 0, -2 * 9.81])
    acc = lambda o: np.linalg.norm(np.array(o["vel"]) / 0.02 - g)  # noqa: E731
    a, b = _run(_sim(O), [(1.0, 1.0)])[0], _run(_sim(O, O.Q_THRUST_VECTOR), [(1.0, 1.0)])[0]
    s, c = math.sin(math.radians(18)), math.cos(math.radians(18))
    assert abs(acc(b) - 35.0 / 2.0) < 2e-2                                        # cleared: |F| = thrust
    assert abs(acc(a) - 35.0 / 2.0 * math.sqrt(2 * s * s + c ** 4)) < 2e-2        # reference: 35 |[s, s, c c]|


def test_q3_frozen_forces(oracle_mod):
    O = oracle_mod
    assert _run(_sim(O), ZERO * 3)[-1]["vel"] == _run(_sim(O, O.Q_FROZEN_FORCES), ZERO * 3)[-1]["vel"]   # upright, no deflection: same
    a, b = _run(_sim(O), [(1.0, 0.0)] * 3)[-1], _run(_sim(O, O.Q_FROZEN_FORCES), [(1.0, 0.0)] * 3)[-1]
    d = np.abs(np.array(a["vel"]) - np.array(b["vel"])).max()
    assert 1e-7 < d < 1e-2                                    # the thrust direction follows the turning body within the step


def test_q5_drag_cutoff(oracle_mod):
    O = oracle_mod
    a, b = _run(_sim(O), ZERO * 2), _run(_sim(O, O.Q_DRAG_CUTOFF), ZERO * 2)
    assert a[0]["vel"] == b[0]["vel"]                         # at rest there is no drag either way
    assert abs(a[0]["vel"][2]) < 0.1                          # second step starts below the cut-off speed ...
    assert abs(b[1]["vel"][2]) < abs(a[1]["vel"][2])          # ... so only the cleared variant is slowed by drag
    assert abs(b[1]["vel"][2] - a[1]["vel"][2]) < 1e-5


def test_q6_stacked_damping(oracle_mod):
    O = oracle_mod
    a, b = _run(_sim(O), [(1.0, 0.0)] * 3)[-1], _run(_sim(O, O.Q_STACKED_DAMPING), [(1.0, 0.0)] * 3)[-1]
    wa, wb = np.linalg.norm(a["omega"]), np.linalg.norm(b["omega"])
    assert wb > wa > 0 and (wb - wa) / wa < 1e-2              # less damping -> slightly faster rotation
    assert a["pos"][2] == b["pos"][2] or abs(a["pos"][2] - b["pos"][2]) < 1e-6


def test_q7_euler_tilt(oracle_mod):
    O = oracle_mod
    res = []
    for clear in (0, O.Q_EULER_TILT):
        sim = _sim(O, clear)
        q = sim.env(0).body.quat
        q[0], q[1], q[2], q[3] = 0.0, 0.0, math.sin(0.15), math.cos(0.15)     # heading 0.3 rad about the vertical
        res.append(_run(sim, ZERO)[0])
    assert abs(res[0]["tilt"] - 0.3) < 1e-6                   # reference: heading counts as tilt
    assert res[1]["tilt"] < 1e-6                              # cleared: the axis is vertical
    assert res[0]["pos"] == res[1]["pos"] and res[0]["term"] == res[1]["term"] is False


def test_q12_diversity_bonus(oracle_mod):
    O = oracle_mod
    a, b = _run(_sim(O), ZERO * 6), _run(_sim(O, O.Q_DIVERSITY_BONUS), ZERO * 6)
    assert [x["comp"][11] for x in a[1:]] == [1.0] * 5 and all(x["comp"][11] == 0.0 for x in b)
    for x, y in zip(a[1:], b[1:]):
        assert abs((x["r"] - y["r"]) - 0.05) < 1e-12 and x["pos"] == y["pos"]
    assert a[0]["r"] == b[0]["r"]                              # empty history: no bonus on the first call either way


def test_q13_variance_penalty(oracle_mod):
    O = oracle_mod
    hard = [(1.0, -1.0)] * 60
    a, b = _run(_sim(O), hard), _run(_sim(O, O.Q_VARIANCE_PENALTY), hard)
    crash = next(i for i, x in enumerate(a) if x["comp"][6] == -1000.0)
    assert all(x["r"] == -1000.0 for x in a[crash:crash + 10])                  # reference: ten steps at the clip
    assert b[crash]["r"] == -1000.0 and all(x["r"] > -1000.0 for x in b[crash + 1:crash + 10] if x["comp"][6] == 0.0)
    assert [x["pos"] for x in a] == [x["pos"] for x in b]                       # physics untouched


def test_q16_success_masks_truncation(oracle_mod):
    O = oracle_mod
    a = _run(_sim(O, max_episode_steps=100), ZERO * 100)
    b = _run(_sim(O, O.Q_SUCCESS_MASKS_TRUNCATION, max_episode_steps=100), ZERO * 100)
    assert a[99]["success"] and a[99]["term"] and not a[99]["trunc"]            # reference: success hides the step limit
    assert b[99]["success"] and b[99]["term"] and b[99]["trunc"]
    assert [x["r"] for x in a] == [x["r"] for x in b]


def test_q17_crash_is_com_height(oracle_mod):
    O = oracle_mod
    hard = [(1.0, -1.0)] * 60
    a, b = _run(_sim(O), hard), _run(_sim(O, O.Q_CRASH_IS_COM_HEIGHT), hard)
    ia = next(i for i, x in enumerate(a) if x["comp"][6] == -1000.0)
    ib = next(i for i, x in enumerate(b) if x["comp"][6] == -1000.0)
    assert a[ia]["alt"] < 0.1 <= a[ia - 1]["alt"]                               # reference: centre of mass below 0.1 m
    assert ib < ia and b[ib]["alt"] > 0.1                                       # cleared: the tilted touchdown itself
    assert [x["pos"] for x in a[:ib]] == [x["pos"] for x in b[:ib]]


def test_q8_q9_lagged_phase(oracle_mod):
    O = oracle_mod
    a, b = _run(_sim(O), ZERO * 100), _run(_sim(O, O.Q_LAGGED_PHASE), ZERO * 100)
    # success fires on step 100 (index 99); the reference's reward sees the pre-update flag, the cleared variant pays R1 at once
    assert a[99]["success"] and a[99]["comp"][0] == 0.0 and b[99]["comp"][0] == 100.0
    assert [x["pos"] for x in a] == [x["pos"] for x in b]


def test_q10_q11_histories_survive_reset(oracle_mod):
    O = oracle_mod
    for clear, want_first_step_success in ((0, True), (O.Q_KEEP_CRITERIA, False)):
        sim = _sim(O, clear)
        _run(sim, ZERO * 100)
        sim.reset()
        assert _run(sim, ZERO)[0]["success"] == int(want_first_step_success)     # Q10
    for clear, want in ((0, False), (O.Q_KEEP_REWARD_HIST, True)):
        sim = _sim(O, clear)
        _run(sim, [(0.5, 0.5)] * 3)
        sim.reset()
        first = _run(sim, ZERO)[0]
        assert (first["comp"][4] == 5.0) == want                                 # Q11: smoothness 1.0 only without a previous action

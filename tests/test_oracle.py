"""CPU tests of the oracle: golden trajectories produced by the reference's own Python class
(tests/golden/make_golden.py), known-answer vectors and analytic checks.  No GPU."""
import math
import os

import numpy as np
import pytest

SCENARIOS = ["zero_120", "random_raw", "random_autoreset", "two_episodes", "burnout_1100", "crash_leak"]


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, f"{name}.npz"))


@pytest.mark.parametrize("name", SCENARIOS)
def test_env_layer_matches_reference_python(oracle_mod, golden_dir, name):
    """The C env layer must reproduce the reference class (run over the same physics layer):
    every flag, phase and counter exactly; floats to fp64 round-off (float32 intermediates of
    NumPy 2 in three reward terms allow 1e-6)."""
    O = oracle_mod
    g = _load(golden_dir, name)
    sim = O.OracleSim(O.default_config(O.CONTRACT_R), 1)
    T = len(g["reward"])
    for t in range(T):
        obs, r, term, trunc, outs = sim.step(g["actions"][t:t + 1])
        o = outs[0]
        assert bool(o.terminated) == bool(g["terminated"][t]), (name, t)
        assert bool(o.truncated) == bool(g["truncated"][t]), (name, t)
        assert o.phase == g["phase"][t] and bool(o.success) == bool(g["success"][t]), (name, t)
        assert o.step == g["step"][t] and bool(o.criteria_met) == bool(g["criteria_met"][t]), (name, t)
        assert o.fuel == g["fuel"][t]
        np.testing.assert_allclose(obs[0], g["obs"][t], rtol=0, atol=1e-9)
        np.testing.assert_allclose(r[0], g["reward"][t], rtol=0, atol=2e-6)
        np.testing.assert_allclose(list(o.comp)[:9], g["comp"][t], rtol=0, atol=2e-6)
        e = sim.env(0)
        st = list(e.body.pos) + list(e.body.quat) + list(e.body.vel) + list(e.body.omega)
        np.testing.assert_allclose(st, g["state"][t], rtol=0, atol=1e-9)
        np.testing.assert_allclose(np.degrees(o.tilt), g["tilt_deg"][t], rtol=0, atol=1e-9)
        if g["was_reset"][t]:
            np.testing.assert_array_equal(sim.reset()[0], g["next_obs"][t])


def test_golden_expected_events(golden_dir):
    """Facts SURVEY.md section 8(a) predicts from reading the reference."""
    z = _load(golden_dir, "zero_120")
    assert np.flatnonzero(z["terminated"])[0] == 99            # success after 100 criteria pushes (S10)
    assert z["success"][99] and not z["success"][98]
    two = _load(golden_dir, "two_episodes")
    assert two["terminated"][100] and two["step"][100] == 1    # Q10: episode 2 succeeds on its first step
    b = _load(golden_dir, "burnout_1100")
    assert b["fuel"][198] >= 0.8 > b["fuel"][199]              # S4: first fuel < 0.8 after decrement 200
    assert b["fuel"][898] > 0.1 and not b["fuel"][899] > 0.1   # first not > 0.1 at 900
    assert b["fuel"][998] > 0 and b["fuel"][999] == 0          # reaches 0 at 1000
    assert b["truncated"][999] == False and b["terminated"][999]  # Q16: success masks truncation  # noqa: E712
    c = _load(golden_dir, "crash_leak")
    crash = np.flatnonzero(c["altitude"] < 0.1)
    assert len(crash) and np.all(c["reward"][crash[0]:crash[0] + 10] == -1000.0)  # Q13


def test_philox_known_answers(oracle_mod):
    """Random123 kat_vectors for philox4x32-10."""
    O = oracle_mod
    assert O.philox4x32_10((0, 0, 0, 0), (0, 0)) == (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)
    assert O.philox4x32_10((0xffffffff,) * 4, (0xffffffff,) * 2) == (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)
    assert O.philox4x32_10((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0)) == \
        (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)


def test_fuel_table_thresholds_and_float32_formula(oracle_mod):
    """Row S4: iterated fp64 subtraction decides steps 200/900/1000, and the kernel's
    (float)(1.0 - n*0.001) equals its float32 rounding for every n."""
    L = oracle_mod.lib()
    tab = [L.orc_fuel_table(n) for n in range(1002)]
    f = 1.0
    for n in range(1, 1002):
        f = max(0, f - 0.001)
        assert tab[n] == f
    assert next(n for n in range(1002) if tab[n] < 0.8) == 200
    assert next(n for n in range(1002) if not tab[n] > 0.1) == 900
    assert next(n for n in range(1002) if tab[n] == 0) == 1000
    for n in range(1000):
        assert np.float32(tab[n]) == np.float32(1.0 - n * 0.001)


def test_diversity_threshold_is_integer_compare():
    for L in range(0, 1001):
        for d in range(0, L + 1):
            assert (d > L * 0.8) == (5 * d > 4 * L)


def test_euler_axis_aligned(oracle_mod):
    """Row B8 on known quaternions."""
    O = oracle_mod
    assert O.euler_from_quat((0, 0, 0, 1)) == (0.0, 0.0, 0.0)
    s = math.sin(0.25)
    r, p, y = O.euler_from_quat((s, 0, 0, math.cos(0.25)))
    assert abs(r - 0.5) < 1e-15 and p == 0 and y == 0
    r, p, y = O.euler_from_quat((0, s, 0, math.cos(0.25)))
    assert abs(p - 0.5) < 1e-15 and r == 0 and y == 0
    r, p, y = O.euler_from_quat((0, 0, s, math.cos(0.25)))
    assert abs(y - 0.5) < 1e-15 and r == 0 and p == 0
    h = math.sqrt(0.5)
    r, p, y = O.euler_from_quat((0, h, 0, h))            # gimbal lock branch
    assert r == 0 and abs(p - math.pi / 2) < 1e-12
    # B7: sign canonicalisation
    q = O.reported_quat((0.1, -0.2, 0.3, -math.sqrt(1 - 0.14)))
    assert q[3] > 0 and abs(q[0] + 0.1) < 1e-15


def _free_body(O, **kw):
    import ctypes as C
    p = O.BodyParams()
    O.lib().orc_body_params_default(C.byref(p))
    p.ground = 0
    for k, v in kw.items():
        setattr(p, k, v)
    b = O.Body()
    O.lib().orc_body_init(C.byref(b), O._d((0, 0, 100.0)), O._d((0, 0, 0, 1)))
    return p, b


def test_free_fall_matches_damped_recurrence(oracle_mod):
    """Rows B4/B5: zero torque, gravity only: v_{k+1} = v_k + dt (g - v_k k_L (1 + |v_k|)),
    z_{k+1} = z_k + dt v_{k+1} (semi-implicit); first step from rest gives dv = g*dt exactly."""
    import ctypes as C
    O = oracle_mod
    p, b = _free_body(O)
    v, z, dt = 0.0, 100.0, 0.005
    for step in range(50):
        O.lib().orc_step_simulation(C.byref(p), C.byref(b), None)
        for _ in range(4):
            v = v + (-9.81 - v * (0.01 + 0.01 * abs(v))) * dt
            z = z + dt * v
        assert abs(b.vel[2] - v) < 1e-12 and abs(b.pos[2] - z) < 1e-10
    assert b.quat[3] == 1.0 and b.omega[0] == 0.0


def test_torque_free_rotation(oracle_mod):
    """Row B6: constant omega about z with damping off keeps |q| = 1 and advances the angle by w*t."""
    import ctypes as C
    O = oracle_mod
    p, b = _free_body(O, ang_damp=0.0, lin_damp=0.0)
    p.gravity[2] = 0.0
    b.omega[2] = 1.5
    for _ in range(100):
        O.lib().orc_step_simulation(C.byref(p), C.byref(b), None)
    ang = 2 * math.atan2(b.quat[2], b.quat[3])
    assert abs(ang - 1.5 * 2.0) < 1e-9
    assert abs(sum(x * x for x in b.quat) - 1.0) < 1e-14


def test_resting_contact_is_stable(oracle_mod):
    """Our contact model: zero action -> touches down near step 35, rests upright at z ~ 0.5."""
    O = oracle_mod
    sim = O.OracleSim(O.default_config(O.CONTRACT_R), 1)
    z = []
    for t in range(99):
        _, _, term, _, outs = sim.step(np.zeros((1, 2), np.float32))
        z.append(outs[0].altitude)
        assert not term[0]
    assert 0.499 < z[-1] < 0.502 and abs(z[-1] - z[-20]) < 1e-6
    assert min(z) > 0.49


def test_contract_x_draws_are_reproducible_and_bounded(oracle_mod):
    O = oracle_mod
    a = O.OracleSim(O.default_config(O.CONTRACT_X, init_tilt_max=0.2, init_omega_max=0.1), 256)
    b = O.OracleSim(O.default_config(O.CONTRACT_X, init_tilt_max=0.2, init_omega_max=0.1, env_id_base=128), 128)
    ms = np.array([a.env(i).mass_scale for i in range(256)])
    ts = np.array([a.env(i).thrust_scale for i in range(256)])
    assert ms.min() >= 0.7 and ms.max() <= 1.3 and ms.std() > 0.1
    assert ts.min() >= 0.4 and ts.max() <= 1.6
    for i in range(128):   # global env ids make draws independent of sharding
        assert a.env(128 + i).mass_scale == b.env(i).mass_scale and a.env(128 + i).wind[0] == b.env(i).wind[0]
    acts = a.random_actions(3)
    assert acts.min() >= -1 and acts.max() <= 1 and abs(acts.mean()) < 0.1
    np.testing.assert_array_equal(acts[128:], b.random_actions(3))


def test_fast_diversity_matches_exact_on_goldens(oracle_mod, golden_dir):
    """The fast duplicate bookkeeping (clip + run bits) gives the same bonus decision as len(set())
    on every golden step except where non-adjacent repeats occur (deterministic replays)."""
    O = oracle_mod
    for name, allowed in (("random_raw", 0), ("random_autoreset", 0), ("crash_leak", 0), ("zero_120", 0)):
        g = _load(golden_dir, name)
        ex = O.OracleSim(O.default_config(O.CONTRACT_R, diversity_mode=O.DIV_EXACT), 1)
        fa = O.OracleSim(O.default_config(O.CONTRACT_R, diversity_mode=O.DIV_FAST), 1)
        diff = 0
        for t in range(len(g["reward"])):
            _, r1, _, _, o1 = ex.step(g["actions"][t:t + 1])
            _, r2, _, _, o2 = fa.step(g["actions"][t:t + 1])
            diff += o1[0].comp[11] != o2[0].comp[11]
            if g["was_reset"][t]:
                ex.reset(), fa.reset()
        assert diff <= allowed, (name, diff)


def test_fp32_sensitivity_of_the_model(oracle_mod, golden_dir):
    """How much does evaluating the physics layer in float32 (the device's arithmetic) move one
    teacher-forced step away from fp64?  Uses the -DORC_PHYS_FLOAT build of the same C source.  This
    bounds what the GPU parity tests can demand: free-flight steps stay within K*1e-5; steps in ground
    contact are amplified by the PGS rows (1/I_zz = 400) up to a few 1e-4."""
    import ctypes as C
    O = oracle_mod
    g = _load(golden_dir, "random_raw")
    a = O.OracleSim(O.default_config(O.CONTRACT_R), 1)
    b = O.OracleSim(O.default_config(O.CONTRACT_R), 1, f32_physics=True)
    free, contact = [], []
    for t in range(400):
        C.memmove(C.byref(b.env(0)), C.byref(a.env(0)), C.sizeof(O.Env))
        eb = b.env(0)
        for f in ("pos", "quat", "vel", "omega"):
            arr = getattr(eb.body, f)
            for i in range(len(arr)):
                arr[i] = float(np.float32(arr[i]))
        z_pre = a.env(0).body.pos[2]
        a.step(g["actions"][t:t + 1])
        b.step(g["actions"][t:t + 1])
        ea = a.env(0)
        sa = np.array(list(ea.body.pos) + list(ea.body.quat) + list(ea.body.vel) + list(ea.body.omega))
        sb = np.array(list(eb.body.pos) + list(eb.body.quat) + list(eb.body.vel) + list(eb.body.omega))
        err = float((np.abs(sa - sb) / np.maximum(1, np.abs(sa))).max())
        (free if min(z_pre, ea.body.pos[2]) > 0.62 else contact).append(err)
    assert len(free) >= 10 and len(contact) > 100
    assert max(free) <= 4e-5
    assert max(contact) <= 1e-3 and np.median(contact) < 2e-5


def _one_step_error_vs_converged(O, ci, wi, n=512, steps=90, settle=30):
    """One-step difference between a (ci, wi)-pass contact solve and a (60, 60)-pass solve of the SAME step from the SAME
    state, over the near-ground env-steps of a Contract-X batch in its steady mix (every step the converged sim is re-seeded
    with the other sim's state, so one-step errors are measured, not drift)."""
    import ctypes as C
    over = dict(autoreset=1, init_tilt_max=0.2, init_omega_max=0.1)
    a = O.OracleSim(O.default_config(O.CONTRACT_X, contact_iters=ci, contact_warm_iters=wi, **over), n)
    b = O.OracleSim(O.default_config(O.CONTRACT_X, contact_iters=60, contact_warm_iters=60, **over), n)
    a.reset(), b.reset()
    rng = np.random.default_rng(7)
    errs = []
    for t in range(steps):
        acts = rng.uniform(-1, 1, (n, 2)).astype(np.float32)
        if t >= settle:     # episodes desynchronised: a quarter of the envs is on the ground
            for i in range(n):
                C.memmove(C.addressof(b.env(i)), C.addressof(a.env(i)), C.sizeof(O.Env))
            pre_z = np.array([a.env(i).body.pos[2] for i in range(n)])
        _, _, ta, tra, _ = a.step(acts, threads=4)
        if t < settle:
            continue
        _, _, tb, trb, _ = b.step(acts, threads=4)
        for i in np.flatnonzero((pre_z < 0.75) & ~(ta | tra | tb | trb)):
            ea, eb = a.env(i).body, b.env(i).body
            sa = np.array(list(ea.pos) + list(ea.quat) + list(ea.vel) + list(ea.omega))
            sb = np.array(list(eb.pos) + list(eb.quat) + list(eb.vel) + list(eb.omega))
            errs.append(np.max(np.abs(sa - sb) / np.maximum(1.0, np.abs(sb))))
    a.close(), b.close()
    return np.array(errs)


def test_block_solver_pass_counts(oracle_mod):
    """DESIGN.md section 4, item 6.  The pass counts are part of the model (as Bullet's 50 iterations are part of its): the
    shipped (2 cold, 1 warm) passes agree with a converged solve to 2e-7 in the median near-ground env-step, and differ from
    it where several points bind at once (impacts on the flat cap) -- 18 % of the near-ground env-steps by more than 1e-4.
    More passes converge monotonically; `contact_iters` / `contact_warm_iters` are tvc_config fields on both sides."""
    O = oracle_mod
    assert (O.default_config(O.CONTRACT_X).contact_iters, O.default_config(O.CONTRACT_X).contact_warm_iters) == (2, 1)
    rows = {}
    for ci, wi in ((2, 1), (8, 3), (16, 8)):
        e = _one_step_error_vs_converged(O, ci, wi)
        assert len(e) > 2000
        rows[(ci, wi)] = (np.median(e), np.quantile(e, 0.9), np.quantile(e, 0.99), float((e > 1e-4).mean()))
        print(f"\n[block solver ({ci},{wi}) vs (60,60) passes] {len(e)} near-ground env-steps: median {rows[(ci, wi)][0]:.1e} "
              f"q90 {rows[(ci, wi)][1]:.1e} q99 {rows[(ci, wi)][2]:.1e} share > 1e-4: {rows[(ci, wi)][3]:.3f}")
    assert rows[(2, 1)][0] <= 1e-5 and rows[(2, 1)][3] <= 0.25
    assert rows[(8, 3)][1] < rows[(2, 1)][1] and rows[(16, 8)][1] < rows[(8, 3)][1]      # monotone in the passes
    assert rows[(16, 8)][2] <= 1e-3


def test_fp32_sensitivity_contract_x_batch(oracle_mod):
    """The same question for the workload of the GPU parity test (Contract X, K = 10, 512 envs x 40 steps with DR, delay ring
    and thrust curve): the fp32 build of the ORACLE's own physics layer against its fp64 build, teacher-forced from identical
    fp32 states.  Its contact tail (q99 ~1e-5, max ~1e-3) is what a plain-float kernel shows against the fp64 oracle: it
    belongs to evaluating the model in fp32, not to the CUDA implementation.
    What the tail is (measured here): NOT a discrete decision -- the manifold's entry / reach rules of the worst steps sit
    millimetres from their thresholds (orc_step_out.contact_margin), and both builds take the same solver branches -- but
    gain: the normal targets divide the gap by dt (x500), the gap sees the attitude through the 0.5-0.6 m arm along the body
    axis, and rim friction turns the target into spin through 1 / I.  The fp64 oracle itself answers a 1e-10 m change of the
    input height with up to 2e-6 rad/s on such steps (last part of this test).  That is why the kernel carries the attitude
    and the two cap-centre heights in double near the ground (tvc_device.cuh HpAtt): the THIRD build of the oracle
    (-DORC_PHYS_FLOAT -DORC_GEO_DOUBLE: float everywhere except exactly those quantities) loses most of the tail."""
    import ctypes as C
    O = oracle_mod
    n, K = 512, 10
    over = dict(init_tilt_max=0.2, init_omega_max=0.1, delay_steps=3, thrust_curve=1, propellant_fraction=0.2, cg_burn_shift=0.05,
                autoreset=1, env_id_base=1000)
    a = O.OracleSim(O.default_config(O.CONTRACT_X, **over), n)
    b = O.OracleSim(O.default_config(O.CONTRACT_X, **over), n, f32_physics=True)
    g = O.OracleSim(O.default_config(O.CONTRACT_X, **over), n, f32_physics=True, geo_double=True)
    a.reset(), b.reset(), g.reset()
    free, contact, margin, split = [], [], [], []
    worst = (0.0, None, None, None)
    for t in range(40):
        for i in range(n):
            ea = a.env(i)
            for f in ("pos", "quat", "vel", "omega"):          # identical fp32 inputs on both sides
                arr = getattr(ea.body, f)
                for k in range(len(arr)):
                    arr[k] = float(np.float32(arr[k]))
            C.memmove(C.addressof(b.env(i)), C.addressof(ea), C.sizeof(O.Env))
            C.memmove(C.addressof(g.env(i)), C.addressof(ea), C.sizeof(O.Env))
        pre_z = np.array([a.env(i).body.pos[2] for i in range(n)])
        blobs = [C.string_at(C.addressof(a.env(i)), C.sizeof(O.Env)) for i in range(n)]
        acts = a.random_actions(t)
        _, _, ta, tra, outs = a.step(acts, threads=4)
        _, _, tb, trb, _ = b.step(acts, threads=4)
        _, _, tg, trg, _ = g.step(acts, threads=4)
        for i in np.flatnonzero(~(ta | tra | tb | trb | tg | trg)):
            ea, eb, eg = a.env(i).body, b.env(i).body, g.env(i).body
            sa = np.array(list(ea.pos) + list(ea.quat) + list(ea.vel) + list(ea.omega))
            sb = np.array(list(eb.pos) + list(eb.quat) + list(eb.vel) + list(eb.omega))
            sg = np.array(list(eg.pos) + list(eg.quat) + list(eg.vel) + list(eg.omega))
            err = float((np.abs(sa - sb) / np.maximum(1, np.abs(sa))).max())
            if min(pre_z[i], ea.pos[2]) > 0.75:
                free.append(err)
            else:
                contact.append(err), margin.append(outs[i].contact_margin)
                split.append(float((np.abs(sa - sg) / np.maximum(1, np.abs(sa))).max()))
                if err > worst[0]:
                    worst = (err, int(i), blobs[i], acts[i].copy())
    free, contact, margin, split = np.array(free), np.array(contact), np.array(margin), np.array(split)
    tail = contact > 1e-4
    print(f"\n[fp32 oracle vs fp64 oracle, Contract X] free-flight max {free.max():.2e} ({len(free)} env-steps); near-ground median "
          f"{np.median(contact):.2e} q99 {np.quantile(contact, 0.99):.2e} max {contact.max():.2e} ({len(contact)} env-steps); "
          f"{int(tail.sum())} steps above 1e-4, their manifold margins {margin[tail].min():.1e} .. {margin[tail].max():.1e} m "
          f"(median {np.median(margin[tail]):.1e})")
    print(f"[fp32 oracle with the attitude and the cap heights in double vs fp64 oracle] near-ground median {np.median(split):.2e} "
          f"q99 {np.quantile(split, 0.99):.2e} max {split.max():.2e}; {int((split > 1e-4).sum())} steps above 1e-4")
    assert len(contact) > 2000 and free.max() <= K * 1e-5
    assert np.quantile(contact, 0.99) <= K * 1e-5 and contact.max() <= 2e-2
    # the kernel's precision split: q99 at 1e-5 per STEP, the maximum several times lower, a third of the steps above 1e-4
    assert np.quantile(split, 0.99) <= 1e-5 and split.max() <= 0.3 * contact.max() and (split > 1e-4).sum() * 2 < tail.sum()
    # the tail does not sit on a manifold threshold: a float evaluation moves a gap by ~1e-7 m at most
    assert tail.sum() >= 5 and margin[tail].min() > 1e-6 and np.median(margin[tail]) > 1e-3
    # ... it is gain: the fp64 model's own response of omega to 1e-10 m of input height on the worst step
    err, i, blob, act = worst

    def final_omega(dz):
        s = O.OracleSim(O.default_config(O.CONTRACT_X, **dict(over, env_id_base=1000 + i)), 1)
        s.reset()
        C.memmove(C.addressof(s.env(0)), blob, len(blob))
        s.env(0).body.pos[2] += dz
        s.step(act[None, :].astype(np.float32), threads=1)
        return np.array(list(s.env(0).body.omega))
    w0, w1, w2 = final_omega(0.0), final_omega(1e-10), final_omega(2e-10)
    gain = np.abs(w1 - w0).max() / 1e-10
    print(f"[fp32 oracle vs fp64 oracle, Contract X] worst step (error {err:.1e}): d omega / d height = {gain:.1e} rad/s per m in the "
          f"fp64 oracle, linear ({np.abs(w2 - w0).max() / 2e-10:.1e} at twice the perturbation)")
    assert gain > 200.0                                                       # 1e-7 m of in-step rounding -> > 2e-5 rad/s
    assert abs(np.abs(w2 - w0).max() / 2e-10 / gain - 1.0) < 0.05             # a smooth response, not a branch


def test_free_flight_converges_to_the_continuous_rigid_body_equations(oracle_mod):
    """Rows B3-B6 against an INDEPENDENT formulation: the Newton-Euler equations of the same body (world-frame force and torque
    held constant, inverse inertia R diag(1/I) R^T, k (1 + |.|) damping, quaternion kinematics q' = (w, 0) (x) q / 2, with and
    without the gyroscopic term) integrated by scipy's adaptive Runge-Kutta to 1e-12.  The oracle's semi-implicit substeps must
    approach that solution at first order (error ratio ~4 for substeps 4x smaller) from a general 3-D state: tilted attitude,
    spin about all three axes, off-axis torque -- a wrong frame, sign or inertia axis anywhere does not converge."""
    import ctypes as C
    from scipy.integrate import solve_ivp
    O = oracle_mod
    F, T = np.array([3.0, -2.0, 25.0]), np.array([0.02, -0.015, 0.004])
    q0 = np.array([0.15, -0.1, 0.2, 0.0])
    q0[3] = math.sqrt(1.0 - float(q0[:3] @ q0[:3]))
    v0, w0 = np.array([1.0, -0.5, 2.0]), np.array([0.8, -1.1, 2.5])

    def rot(q):
        x, y, z, w = q
        return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                         [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                         [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])

    for gyro in (0, 1):
        p, _ = _free_body(O)
        m, I = p.mass, np.array(list(p.inertia))
        g, kl, ka = np.array(list(p.gravity)), p.lin_damp, p.ang_damp
        assert abs(I[0] - I[2]) > 0.1 * I[0]          # an axisymmetric but not spherical body: the frames matter

        def rhs(t, y):
            q, v, w = y[3:7], y[7:10], y[10:13]
            R = rot(q / np.linalg.norm(q))
            wl, tl = R.T @ w, R.T @ T
            wdl = tl / I - wl * (ka + ka * np.linalg.norm(wl))
            if gyro:
                wdl = wdl - np.cross(wl, I * wl) / I
            qd = 0.5 * np.array([w[0] * q[3] + w[1] * q[2] - w[2] * q[1], w[1] * q[3] + w[2] * q[0] - w[0] * q[2],
                                 w[2] * q[3] + w[0] * q[1] - w[1] * q[0], -w[0] * q[0] - w[1] * q[1] - w[2] * q[2]])
            return np.concatenate([v, qd, F / m + g - v * (kl + kl * np.linalg.norm(v)), R @ wdl])

        y0 = np.concatenate([[0, 0, 100.0], q0, v0, w0])
        ref = solve_ivp(rhs, (0.0, 0.2), y0, method="DOP853", rtol=1e-12, atol=1e-14).y[:, -1]
        ref[3:7] /= np.linalg.norm(ref[3:7])
        errs = []
        for sub in (4, 16, 64):
            p, b = _free_body(O, use_gyro=gyro, substeps=sub)
            for k in range(3):
                b.vel[k], b.omega[k] = v0[k], w0[k]
            for k in range(4):
                b.quat[k] = q0[k]
            for _ in range(10):                       # ten control steps of 0.02 s
                for k in range(3):
                    b.force[k], b.torque[k] = F[k], T[k]
                O.lib().orc_step_simulation(C.byref(p), C.byref(b), None)
            got = np.array(list(b.pos) + list(b.quat) + list(b.vel) + list(b.omega))
            if got[3:7] @ ref[3:7] < 0:
                got[3:7] = -got[3:7]
            errs.append(float(np.max(np.abs(got - ref) / np.maximum(1.0, np.abs(ref)))))
        print(f"[free flight vs continuous equations] gyro={gyro}: max rel. error {errs[0]:.2e} / {errs[1]:.2e} / {errs[2]:.2e} at 4 / 16 / 64 substeps")
        assert errs[0] < 5e-2 and errs[2] < 1e-3, (gyro, errs)
        assert 3.0 < errs[0] / errs[1] < 5.0 and 3.0 < errs[1] / errs[2] < 5.0, (gyro, errs)     # first order in the substep

"""The C-ABI library loads and exports every symbol include/tvc_b200.h declares; struct layouts of
the ctypes stub match the header.  No compute calls (no GPU here)."""
import ctypes as C
import os
import re
import subprocess
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "tvc_b200.h")


def _declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tvc_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(lib_built):
    from tvc_ai_b200 import _abi
    L = C.CDLL(lib_built)
    names = _declared_functions()
    assert len(names) >= 18
    for n in names:
        assert hasattr(L, n), f"{n} declared in tvc_b200.h but not exported"
    assert set(names) == set(_abi.EXPORTS)
    assert _abi.load().tvc_abi_version() == _abi.ABI_VERSION


def test_struct_layouts_match_header(lib_built):
    from tvc_ai_b200 import _abi, engine
    prog = r'''
#include <stdio.h>
#include <stddef.h>
#include "tvc_b200.h"
int main(void){
  printf("%zu %zu %zu %zu %zu %zu %zu\n", sizeof(tvc_config), sizeof(tvc_stage_conditions), sizeof(tvc_info_soa),
         sizeof(tvc_step_io), sizeof(tvc_env_state), sizeof(tvc_actor_weights), sizeof(tvc_rollout_io));
  printf("%zu %zu %zu %zu\n", offsetof(tvc_config, dt_step), offsetof(tvc_config, seed), offsetof(tvc_config, env_id_base),
         offsetof(tvc_env_state, ring10));
  return 0; }'''
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "s.c")
        open(src, "w").write(prog)
        cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
        subprocess.run([cc, "-I", os.path.join(ROOT, "include"), src, "-o", os.path.join(d, "s")], check=True)
        out = subprocess.run([os.path.join(d, "s")], check=True, capture_output=True, text=True).stdout.split()
    sizes = [int(x) for x in out]
    assert sizes[:7] == [C.sizeof(_abi.TvcConfig), C.sizeof(_abi.TvcStageConditions), C.sizeof(_abi.TvcInfoSoa),
                         C.sizeof(_abi.TvcStepIO), C.sizeof(_abi.TvcEnvState), C.sizeof(_abi.TvcActorWeights),
                         C.sizeof(_abi.TvcRolloutIO)]
    assert sizes[7:] == [_abi.TvcConfig.dt_step.offset, _abi.TvcConfig.seed.offset, _abi.TvcConfig.env_id_base.offset,
                         _abi.TvcEnvState.ring10.offset]
    assert engine.STATE_DTYPE.itemsize == sizes[4]
    assert engine.STATE_DTYPE.fields["ring10"][1] == sizes[10]


def test_defaults_match_reference_constants(lib_built):
    """tvc_config_default(R) == what enhanced_rocket_tvc_env.py hard-codes (ref:340-341, 412-414, 453-454, 463, 471)."""
    from tvc_ai_b200 import _abi
    c = _abi.default_config(_abi.CONTRACT_R)
    assert (c.substeps, c.max_episode_steps, c.dt_step) == (4, 1000, 0.02)
    assert (c.mass, c.thrust) == (2.0, 35.0) and abs(c.radius - 0.05) < 1e-9 and c.length == 1.0
    assert abs(c.gimbal_max_rad - 0.3141592653589793) < 1e-7
    assert abs(c.lin_damp - 0.01) < 1e-9 and abs(c.ang_damp - 0.02) < 1e-9
    assert c.quirks == _abi.Q_ALL_REFERENCE and c.diversity_mode == _abi.DIV_EXACT
    assert abs(c.gradient_penalty - 0.1) < 1e-7 and abs(c.diversity_bonus - 0.05) < 1e-7
    x = _abi.default_config(_abi.CONTRACT_X)
    assert x.substeps == 10 and abs(x.mass_variation - 0.3) < 1e-7 and x.wind_std == 3.0
    assert abs(x.sensor_noise_std - 0.02) < 1e-7 and abs(x.thrust_std - 0.2) < 1e-7
    # Contract X keeps every reference convention except the two cross-episode leaks (Q10, Q11)
    assert x.quirks == _abi.Q_CONTRACT_X == _abi.Q_ALL_REFERENCE & ~(_abi.Q_KEEP_CRITERIA | _abi.Q_KEEP_REWARD_HIST)
    # contact material: Bullet's combination of ref:349-352 (plane) and ref:455-458 (rocket)
    for cc in (c, x):
        assert abs(cc.contact_mu - 0.24) < 1e-7 and abs(cc.contact_mu_spin - 0.11) < 1e-7 and abs(cc.contact_mu_roll - 0.055) < 1e-7
        assert abs(cc.contact_restitution - 0.1) < 1e-7 and abs(cc.contact_erp - 0.2) < 1e-7 and abs(cc.contact_margin - 0.05) < 1e-7
        assert (cc.contact_iters, cc.contact_warm_iters) == (2, 1)


def test_no_cpu_fallback_without_gpu(lib_built):
    """The product path must fail loudly when no CUDA device is present."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from tvc_ai_b200.engine import BatchedEngine
    with pytest.raises(RuntimeError):
        BatchedEngine(4)


def test_product_never_imports_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py may touch oracle/."""
    pkg = os.path.join(ROOT, "tvc_ai_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "tvc_oracle.h" not in txt, f


def test_error_codes_without_a_gpu(lib_built):
    """Error behaviour of the C ABI (include/tvc_b200.h): negative codes + tvc_last_error(), never an exception or exit.
    Argument validation happens before any CUDA call, so these run on the CPU box."""
    from tvc_ai_b200 import _abi as A
    L = A.load()
    h = C.c_void_p()
    cfg = A.default_config(A.CONTRACT_R)
    assert L.tvc_config_default(None, 0) == -1 and b"bad argument" in L.tvc_last_error()
    assert L.tvc_config_default(C.byref(cfg), 7) == -1
    bad = A.default_config(A.CONTRACT_R)
    bad.abi_version = 1
    assert L.tvc_create(C.byref(bad), 0, 16, C.byref(h)) == -4 and b"abi_version" in L.tvc_last_error()      # TVC_E_ABI
    assert L.tvc_create(C.byref(cfg), 0, 0, C.byref(h)) == -1 and b"num_envs" in L.tvc_last_error()          # TVC_E_BADARG
    assert L.tvc_create(C.byref(cfg), 0, 2 ** 31 - 1000, C.byref(h)) == -1 and b"num_envs" in L.tvc_last_error()   # env ids are int32
    assert L.tvc_create(C.byref(cfg), 0, 16, None) == -1
    for field, value in (("substeps", 0), ("substeps", 65), ("max_episode_steps", 0), ("diversity_mode", 3),
                         ("delay_steps", 5), ("contact_iters", -1), ("contract", 2), ("mass", 0.0), ("dt_step", 0.0),
                         ("gimbal_max_rad", 0.9), ("gimbal_max_rad", 0.0), ("mass_variation", 1.0), ("mass_variation", -0.1),
                         ("thrust_lo", 2.0), ("propellant_fraction", 1.0), ("contact_mu", -0.1), ("contact_margin", 0.0),
                         ("contact_restitution", 1.5), ("quirks", 1 << 20)):
        c2 = A.default_config(A.CONTRACT_X)
        setattr(c2, field, value)
        assert L.tvc_create(C.byref(c2), 0, 16, C.byref(h)) == -1, field
        assert not h.value
    # NULL handles are rejected everywhere
    assert L.tvc_step(None, None, None, None, None, None, None, None) == -1 and b"handle" in L.tvc_last_error()
    assert L.tvc_reset(None, None, 0, None, None) == -1
    assert L.tvc_destroy(None) == 0
    assert L.tvc_num_envs(None) == -1 and L.tvc_state_bytes(None) == 0
    import torch
    if not torch.cuda.is_available():
        rc = L.tvc_create(C.byref(cfg), 0, 16, C.byref(h))
        assert rc in (-2, -1) and not h.value          # no device: CUDA error, still no crash

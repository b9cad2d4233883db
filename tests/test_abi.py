"""The C-ABI library loads on a CPU-only box, exports every symbol include/mppi_b200.h declares,
agrees with the ctypes mirror on structure sizes, and fails loudly (no fallback) without a GPU."""
import os
import re

import pytest

import navigation2_b200 as nb
from navigation2_b200 import capi, params as P

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "mppi_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mppi_[a-z_0-9]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    lib = capi.load_library()
    names = declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/mppi_b200.h but not exported"
        assert n in capi.SYMBOLS, f"{n} has no ctypes prototype in navigation2_b200/capi.py"
    for n in capi.SYMBOLS:
        assert n in names, f"{n} bound in capi.py but not declared in the header"


def test_struct_sizes_match():
    capi.check_abi()


def test_defaults_are_the_reference_code_defaults():
    cfg = P.make_config()
    # src/optimizer.cpp:105-129
    assert (cfg.batch_size, cfg.time_steps, cfg.iteration_count) == (1000, 56, 1)
    assert abs(cfg.model_dt - 0.05) < 1e-9 and abs(cfg.temperature - 0.3) < 1e-7 and abs(cfg.gamma - 0.015) < 1e-8
    assert abs(cfg.vx_max - 0.5) < 1e-7 and abs(cfg.vx_min + 0.35) < 1e-7 and abs(cfg.wz_max - 1.9) < 1e-7
    assert abs(cfg.az_max - 3.5) < 1e-7 and cfg.retry_attempt_limit == 1 and cfg.sgf_order == 2
    c = P.critic_params("CostCritic")
    assert abs(c.cost_weight - 3.81) < 1e-6 and c.near_collision_cost == 253 and c.trajectory_point_step == 2
    assert abs(P.critic_params("PathAlignCritic").cost_weight - 10.0) < 1e-6
    assert P.critic_params("PathFollowCritic").offset_from_furthest == 6
    assert abs(P.critic_params("PathAngleCritic").max_angle_to_furthest - 0.785398) < 1e-6
    assert abs(P.critic_params("mppi::critics::ObstaclesCritic").collision_cost - 100000.0) < 1e-3
    with pytest.raises(ValueError):
        P.critic_params("NoSuchCritic")
    lib = capi.load_library()
    for i, n in enumerate(capi.CRITIC_NAMES):
        assert lib.mppi_critic_type_from_name(n.encode()) == i
        assert lib.mppi_critic_name(i).decode() == n


def test_bringup_config_is_baseline_configs_1():
    cfg = P.nav2_bringup_config()
    assert (cfg.batch_size, cfg.time_steps) == (2000, 56)
    names = [capi.CRITIC_NAMES[cfg.critics[i].type] for i in range(cfg.n_critics)]
    assert names == P.NAV2_BRINGUP_CRITICS
    assert abs(cfg.critics[4].cost_weight - 14.0) < 1e-6          # PathAlign yaml weight
    assert abs(cfg.inscribed_radius - 0.22 * 0.9807852804) < 1e-6   # 16-gon apothem (footprint.cpp:158-174)


def test_no_cpu_fallback():
    """Without a CUDA device the product refuses to run (the oracle is never used in its place)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(nb.CudaUnavailable):
        nb.Optimizer(P.nav2_bringup_config())


def test_product_does_not_reference_the_oracle():
    """Nothing under navigation2_b200/ may import, link or mention oracle code."""
    pkg = os.path.join(ROOT, "navigation2_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".hpp")):
                text = open(os.path.join(d, f)).read()
                assert "oracle.binding" not in text and "mppi_oracle" not in text and "liboracle" not in text, \
                    f"{f} references the oracle"


def test_inflation_header_symbols_are_exported_and_bound():
    from navigation2_b200 import inflation
    text = open(os.path.join(ROOT, "include", "costmap_inflation_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = sorted(set(re.findall(r"\b(costmap_inflation_[a-z_0-9]+)\s*\(", text)))
    lib = inflation.load_library()
    assert len(names) >= 9
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/costmap_inflation_b200.h but not exported"
        assert n in inflation.SYMBOLS, f"{n} has no ctypes prototype in navigation2_b200/inflation.py"
    assert sorted(inflation.SYMBOLS) == names
    import ctypes as C
    assert C.sizeof(inflation.CostmapInflationConfig) == 40


def test_inflation_fails_loudly_without_a_gpu():
    import torch
    from navigation2_b200 import inflation
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(RuntimeError):
        inflation.InflationLayer(0.05, 0.22, inflation_radius=0.7)


def test_built_kernels_are_sm_100a_and_use_the_instructions_the_design_claims():
    """The SASS of the in-tree library: sm_100a code; bulk-async staging (UBLKCP) and mbarrier
    transactions in kernel A; the fused kernel's cluster exchanges as asynchronous remote stores (STAS) with no
    GPU-wide fence (MEMBAR.ALL.GPU is what every cluster.sync() costs) and one cluster barrier (start-up only);
    packed FP32 pairs (FADD2 / FMUL2 / FFMA2) in the lean rollout.  Skipped where cuobjdump is not installed."""
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not installed")
    lib = os.path.join(ROOT, "navigation2_b200", "libmppi_b200.so")
    out = subprocess.run([cuobjdump, "-sass", lib], check=True, capture_output=True, text=True).stdout
    assert "arch = sm_100a" in out
    kernels, name = {}, None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            kernels[name] = []
        elif name is not None and "/*" in line:
            kernels[name].append(line)
    def body(fragment):
        found = [k for k in kernels if fragment in k]
        assert found, f"no kernel named *{fragment}* in the library"
        return ["\n".join(kernels[k]) for k in found]
    for sass in body("k_fused_cluster"):
        assert "STAS" in sass and "SYNCS.PHASECHK.TRANS64.TRYWAIT" in sass
        assert "MEMBAR.ALL.GPU" not in sass, "a cluster-scope release crept back into the fused kernel"
        assert sass.count("UCGABAR_ARV") == 1 and sass.count("UCGABAR_WAIT") == 1
        assert "LDGSTS" in sass
    for sass in body("k_rollout_lean"):
        assert "UBLKCP" in sass and "SYNCS.ARRIVE.TRANS64" in sass
        assert "FFMA2" in sass and "FADD2" in sass and "FMUL2" in sass
    assert any("UBLKCP" in sass for sass in body("k_rollout_score"))   # the instantiations with the staged noise ring
    for sass in kernels.values():
        joined = "\n".join(sass)
        assert "HMMA" not in joined and "UTCHMMA" not in joined   # nothing on this path is a contraction

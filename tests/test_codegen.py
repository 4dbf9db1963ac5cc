"""The solver kernel must not be compiled 'divergence-safe' (profiles/README.md): one warp owns one problem,
its control flow is warp-uniform, and se3mpc_core.cuh takes care that ptxas can prove it (uni()/UNI(), ddiv()/dsqrt(), libm-free
rd2quat).  A value that silently becomes 'possibly divergent' again costs ~25% more SASS and ~10% of the throughput without
failing any numerical test, so the property is checked on the built library (static SASS audit, no GPU needed)."""
import os
import shutil
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "learningagileflight_se3_b200", "csrc", "libse3mpc.so")
sys.path.insert(0, os.path.join(ROOT, "profiles"))

UNIFORM_PHASES = ("riccati_factor", "riccati_forward", "riccati_resolve")          # no lane-divergent global access at all
STAGE_PARALLEL = ("apply_step", "eval_compl", "eval_trial", "eval_direction", "eval_constraints", "prepare_stages_exact",
                  "prepare_stages_ident", "eval_kkt")                               # lane-strided stage loops: R2UR only there


@pytest.mark.skipif(shutil.which("cuobjdump") is None or not os.path.exists(LIB), reason="needs cuobjdump and the built library")
@pytest.mark.parametrize("instance", ["2", "3"])
def test_solver_is_not_compiled_divergence_safe(instance):
    from divergence_audit import audit
    a = audit(LIB, instance)
    for f in UNIFORM_PHASES + STAGE_PARALLEL + ("kernel body",):
        assert f in a, (f, sorted(a))
    for f in a:                                                     # every phase, the restoration phase's included (se3mpc_resto.cuh)
        assert a[f]["BRA.DIV"] == 0, (f, dict(a[f]))                # no divergence guards around shuffles
        assert a[f]["WARPSYNC"] <= 4, (f, dict(a[f]))               # (WARPSYNC.ALL of __syncwarp is not emitted as such; COLLECTIVE wrappers are)
    for f in UNIFORM_PHASES + tuple(f + "_soft" for f in UNIFORM_PHASES):
        assert a[f]["R2UR"] == 0 and a[f]["LDCU"] >= 1, (f, dict(a[f]))   # the memory descriptor comes from one LDCU per function
    hot = sum(a[f]["total"] for f in UNIFORM_PHASES + STAGE_PARALLEL)
    assert hot < 8000, hot                                          # 16.7 k instructions (with the kernel body) when everything was divergence-safe
    assert a["kernel body"]["total"] < 12000, a["kernel body"]["total"]   # interior-point loop + inlined restoration-phase control flow + reward

"""The drop-in boundary under the reference's OWN callers: tests/dropin_worker.py swaps three names in the
unmodified reference package and runs its ActorModule / CriticModule / NeuroFlowCT._train_step on the B200 kernels."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*flags):
    from oracle import ref_import

    if not ref_import.reference_available():
        pytest.skip("no reference tree (run baseline/install_ref.py where /root/reference exists)")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "dropin_worker.py"), *flags], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    return json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])


def test_reference_modules_construct_on_the_swapped_classes():
    """CPU: deepcopy + torch.jit.script of the reference's modules (modules.py:370-392) accept the swapped classes; same
    state-dict keys and - the RNG streams being consumed in the same order - the same initial parameters as the golden agent."""
    out = _run("--construct-only")
    assert out["actor_ncp"].startswith("velora_b200.") and out["buffer"].startswith("velora_b200.")
    assert out["scripted"] and out["keys_match"] and out["init_equal"], out


@pytest.mark.gpu
def test_reference_train_step_runs_on_the_b200_kernels():
    """GPU: three `NeuroFlowCT._train_step` calls of the reference's agent code on the fused kernels reproduce the all-reference
    CPU run (tests/golden/train.npz): losses to 2e-5, parameters far below the 3e-4 Adam step."""
    out = _run()
    assert out["actor_ncp"].startswith("velora_b200.") and out["buffer"].startswith("velora_b200.") and out["scripted"]
    assert out["launches"] > 0, "no kernel of this library was launched"
    assert out["loss_ok"], out["losses"]
    assert out["param_max_abs_diff"] < 3e-6 and out["log_alpha_diff"] < 1e-6, out

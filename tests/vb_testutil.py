"""Shared helpers for the test-suite."""
import numpy as np


def liquid_case_params(d, name):
    """state_dict of a golden liquid/ncp case as {key: ndarray}."""
    pre = f"{name}_sd_"
    return {k[len(pre):]: d[k] for k in d.files if k.startswith(pre)}


def rel_err(a, b) -> float:
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    nb = np.linalg.norm(b)
    na = np.linalg.norm(a - b)
    if nb == 0.0:
        return 0.0 if na == 0.0 else float("inf")
    return float(na / nb)


# The numeric contract of BASELINE.json: 1e-5 relative (fp32) on outputs, hidden states and gradients.
FP32_RTOL = 1e-5


def assert_close_to_reference(ours, ref32, ref64, what, rtol=FP32_RTOL):
    """||ours - ref32|| / ||ref32|| <= rtol, and ours is no further from the fp64 arbiter than 4x the fp32
    reference itself is (plus a floor of rtol/4 for quantities the fp32 reference gets almost exactly)."""
    e32 = rel_err(ours, ref32)
    assert e32 <= rtol, f"{what}: rel err vs fp32 reference {e32:.3e} > {rtol:.1e}"
    if ref64 is not None:
        e_ours = rel_err(ours, ref64)
        e_ref = rel_err(ref32, ref64)
        assert e_ours <= max(4.0 * e_ref, rtol / 4), f"{what}: err vs fp64 {e_ours:.3e}, reference's own {e_ref:.3e}"

"""Shared helpers for the test-suite."""


def liquid_case_params(d, name):
    """state_dict of a golden liquid/ncp case as {key: ndarray}."""
    pre = f"{name}_sd_"
    return {k[len(pre):]: d[k] for k in d.files if k.startswith(pre)}

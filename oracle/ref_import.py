"""Import the UNMODIFIED reference hot path from /root/reference (build container) or from its copy under baseline/_ref/
(git-ignored, shipped to the GPU box; written by baseline/install_ref.py).  Test and benchmark infrastructure only.

The reference package cannot be imported normally here: `velora/models/__init__.py:2`,
`velora/utils/__init__.py:1` and `velora/buffer/__init__.py:3` pull in gymnasium / sqlmodel, which are not
installed and cannot be fetched (no network).  This module registers empty package shells for those three
packages and minimal stubs for `gymnasium`, `sqlmodel`, `sqlalchemy`, so that the *unmodified* files
`velora/wiring.py`, `velora/models/lnn/*.py`, `velora/models/weight.py`, `velora/buffer/*.py`,
`velora/utils/{core,torch}.py` are the code that runs.

Users: `oracle/make_golden.py`, tests that skip when no reference tree is present, `bench.py --impl reference` (the
reference arm) and tests/test_dropin_gpu.py (the reference's own agent code driving the B200 kernels).  Nothing in the
product (`velora_b200/`) imports it.  /root/reference does not exist on the GPU box; baseline/_ref/ does.
"""
from __future__ import annotations

import os
import sys
import types

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _find_root() -> str:
    cands = [os.environ.get("VELORA_REFERENCE_ROOT"), "/root/reference", os.path.join(_REPO, "baseline", "_ref")]
    for c in cands:
        if c and os.path.isfile(os.path.join(c, "velora", "wiring.py")):
            return c
    return "/root/reference"


REFERENCE_ROOT = _find_root()


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "velora", "wiring.py"))


def _shell(name: str, path: str) -> None:
    if name in sys.modules:
        return
    m = types.ModuleType(name)
    m.__path__ = [path]  # a package shell: sub-modules resolve from disk, __init__.py is never executed
    sys.modules[name] = m


def _stub_third_party() -> None:
    if "gymnasium" not in sys.modules:
        gym = types.ModuleType("gymnasium")

        class _Env:  # noqa: D401 - placeholder base classes only
            pass

        class _Wrapper:
            def __init__(self, env=None):
                self.env = env

        gym.Env = _Env
        gym.Wrapper = _Wrapper
        vector = types.ModuleType("gymnasium.vector")
        vector.VectorWrapper = _Wrapper
        vector.SyncVectorEnv = _Env
        vector.VectorEnv = _Env
        spaces = types.ModuleType("gymnasium.spaces")
        spaces.Box = type("Box", (), {})
        spaces.Discrete = type("Discrete", (), {})
        wrappers = types.ModuleType("gymnasium.wrappers")
        wrappers.RecordEpisodeStatistics = _Wrapper
        wrappers.NumpyToTorch = _Wrapper
        wrappers.RecordVideo = _Wrapper
        wvec = types.ModuleType("gymnasium.wrappers.vector")
        wvec.NumpyToTorch = _Wrapper
        wvec.RecordEpisodeStatistics = _Wrapper
        wrappers.vector = wvec
        gym.vector = vector
        gym.spaces = spaces
        gym.wrappers = wrappers
        gym.make = lambda *a, **k: (_ for _ in ()).throw(RuntimeError("gymnasium is a stub here"))
        sys.modules["gymnasium"] = gym
        sys.modules["gymnasium.vector"] = vector
        sys.modules["gymnasium.spaces"] = spaces
        sys.modules["gymnasium.wrappers"] = wrappers
        sys.modules["gymnasium.wrappers.vector"] = wvec
    if "sqlmodel" not in sys.modules:
        sm = types.ModuleType("sqlmodel")

        class SQLModel:
            def __init_subclass__(cls, **kw):
                super().__init_subclass__()

            def __init__(self, **kw):
                for k, v in kw.items():
                    setattr(self, k, v)

        sm.SQLModel = SQLModel
        sm.Session = type("Session", (), {})
        sm.create_engine = lambda *a, **k: None
        sm.select = lambda *a, **k: None
        sm.Field = lambda *a, **k: None
        sm.Relationship = lambda *a, **k: None
        sys.modules["sqlmodel"] = sm
        sa = types.ModuleType("sqlalchemy")
        sa.Engine = type("Engine", (), {})
        sa.ScalarResult = type("ScalarResult", (), {"__class_getitem__": classmethod(lambda cls, item: cls)})
        sys.modules["sqlalchemy"] = sa


def import_reference():
    """Returns a namespace with the unmodified reference classes of the hot path."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    _stub_third_party()
    vroot = os.path.join(REFERENCE_ROOT, "velora")
    _shell("velora.models", os.path.join(vroot, "models"))
    _shell("velora.utils", os.path.join(vroot, "utils"))
    _shell("velora.buffer", os.path.join(vroot, "buffer"))

    from velora.models.lnn.cell import NCPLiquidCell
    from velora.models.lnn.ncp import LiquidNCPNetwork, NCPNetwork
    from velora.models.lnn.sparse import SparseLinear
    from velora.utils.core import set_seed
    from velora.wiring import Wiring
    from velora.buffer.replay import ReplayBuffer

    try:   # the agent layer (only needed for the train-step goldens); heavier stubs, optional
        _shell("velora.models.nf", os.path.join(vroot, "models", "nf"))
        from velora.models.nf.modules import ActorModule, CriticModule, EntropyModule
        from velora.models.nf.modules import ActorModuleDiscrete, CriticModuleDiscrete, EntropyModuleDiscrete
        from velora.models.nf.agent import NeuroFlow, NeuroFlowCT
    except Exception as e:   # pragma: no cover
        ActorModule = CriticModule = EntropyModule = NeuroFlowCT = None
        ActorModuleDiscrete = CriticModuleDiscrete = EntropyModuleDiscrete = NeuroFlow = None
        print(f"[ref_import] agent layer not importable: {e!r}")

    ns = types.SimpleNamespace(
        ActorModule=ActorModule,
        CriticModule=CriticModule,
        EntropyModule=EntropyModule,
        NeuroFlowCT=NeuroFlowCT,
        ActorModuleDiscrete=ActorModuleDiscrete,
        CriticModuleDiscrete=CriticModuleDiscrete,
        EntropyModuleDiscrete=EntropyModuleDiscrete,
        NeuroFlow=NeuroFlow,
        NCPLiquidCell=NCPLiquidCell,
        LiquidNCPNetwork=LiquidNCPNetwork,
        NCPNetwork=NCPNetwork,
        SparseLinear=SparseLinear,
        Wiring=Wiring,
        set_seed=set_seed,
        ReplayBuffer=ReplayBuffer,
    )
    return ns

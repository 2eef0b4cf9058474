"""Host-side logic of the product on CPU (no GPU needed): wiring/initialisation parity with the reference goldens,
the reference's structural tests restated against the mirror classes, error behaviour, the C-ABI library."""
import copy
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import velora_b200 as vb
from velora_b200 import _lib
from velora_b200.buffer import BatchExperience, ReplayBuffer
from velora_b200.models.lnn import LiquidNCPNetwork, NCPLiquidCell, NCPNetwork, SparseLinear
from velora_b200.wiring import NeuronCounts, SynapseCounts, Wiring
from vb_testutil import liquid_case_params

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ------------------------------------------------------------------------------------------------ wiring
def test_wiring_bit_exact_with_reference(golden):
    d = golden("wiring")
    for idx, (i, n, o, sp, seed) in enumerate(d["cases"]):
        vb.set_seed(int(seed))
        w = Wiring(int(i), int(n), int(o), sparsity_level=float(sp))
        for k in ("inter", "command", "motor", "recurrent"):
            m = getattr(w.masks, k)
            assert m.dtype == torch.int32
            assert np.array_equal(m.numpy(), d[f"c{idx}_{k}"]), (idx, k)
        c = w.counts
        assert [c.sensory, c.inter, c.command, c.motor] == list(d[f"c{idx}_counts"])
        s = w.n_connections
        assert [s.sensory, s.inter, s.command, s.motor] == list(d[f"c{idx}_synapses"])
        assert np.random.randint(0, 2**31 - 1) == int(d[f"c{idx}_next_np"][0])      # RNG streams left in the same state
        assert int(torch.randint(0, 2**31 - 1, (1,))) == int(d[f"c{idx}_next_torch"][0])


def test_wiring_structure_like_reference_tests():
    # reference tests/models/test_wiring.py:52-54, 56-67, 81-117
    w = Wiring(4, 10, 2, sparsity_level=0.5)
    assert w.n_connections == SynapseCounts(3, 2, 4, 2)
    assert w.counts == NeuronCounts(4, 6, 4, 2)
    for bad in (0.05, 0.95):
        with pytest.raises(ValueError):
            Wiring(4, 10, 2, sparsity_level=bad)
    for sp in (0.1, 0.5, 0.9):
        w = Wiring(8, 30, 3, sparsity_level=sp)
        for k in ("inter", "command", "motor"):
            m = getattr(w.masks, k)
            assert set(m.unique().tolist()) <= {-1, 0, 1}
            assert bool((m != 0).any(dim=0).all()), "every column is connected"
            assert bool((m != 0).any(dim=1).all()), "every row is connected"
    p = Wiring.polarity((5, 3))
    assert p.dtype == torch.int32 and set(p.unique().tolist()) <= {-1, 1}
    masks, counts = w.data()
    assert masks is w.masks and counts is w.counts


# ------------------------------------------------------------------------------------------------ networks (host side)
@pytest.mark.parametrize("name", ["fixture", "actor_t1", "odd", "mid", "wide"])
def test_liquid_init_parity_and_layout(golden, name):
    d = golden("liquid")
    i, n, o, sp, seed = d[f"{name}_meta"][:5]
    vb.set_seed(int(seed))
    net = LiquidNCPNetwork(int(i), int(n), int(o), sparsity_level=float(sp))
    ref = liquid_case_params(d, name)
    sd = net.state_dict()
    assert list(sd.keys()) == list(ref.keys())                      # names and ORDER of the checkpoint format
    for k, v in sd.items():
        assert v.dtype == torch.from_numpy(ref[k]).dtype, k          # inter/motor masks int32, cell masks float32
        assert np.array_equal(v.numpy(), ref[k]), k
    assert [k for k, _ in net.named_parameters()] == [k for k in ref if k.endswith(("weight", "bias"))]


def test_liquid_known_answers_and_attributes():
    vb.set_seed(64)
    net = LiquidNCPNetwork(10, 20, 5)
    assert (net.total_params, net.active_params) == (1017, 657)      # reference tests/models/test_lnn.py:280-309
    assert net.hidden_size == 8 and net.n_units == 25
    assert [n for n, _ in net.named_children()] == ["inter", "command", "motor", "act"]
    assert [n for n, _ in net.command.named_children()] == ["tanh", "sigmoid", "g_head", "h_head", "f_head_to_g", "f_head_to_h", "proj"]
    assert isinstance(net.act, torch.nn.Mish)
    with pytest.raises(ValueError):
        LiquidNCPNetwork(4, 20, 2, sparsity_level=0.95)


def test_ncp_init_parity(golden):
    d = golden("ncp")
    for name in d["names"]:
        i, n, o, sp, seed, _ = d[f"{name}_meta"]
        vb.set_seed(int(seed))
        net = NCPNetwork(int(i), int(n), int(o), sparsity_level=float(sp))
        ref = liquid_case_params(d, str(name))
        assert list(net.state_dict().keys()) == list(ref.keys())
        for k, v in net.state_dict().items():
            assert np.array_equal(v.numpy(), ref[k]), (name, k)
        assert isinstance(net.ncp, torch.nn.Sequential) and len(net.ncp) == 5


def test_no_cpu_fallback_and_errors():
    vb.set_seed(1)
    net = LiquidNCPNetwork(4, 20, 2)
    with pytest.raises(ValueError):
        net(torch.zeros(2, 3, 4))                                    # ncp.py:153-156
    with pytest.raises(RuntimeError, match="no CPU path"):
        net(torch.zeros(2, 4))
    with pytest.raises(RuntimeError, match="no CPU path"):
        net.forward_seq(torch.zeros(2, 3, 4))
    critic = NCPNetwork(5, 128, 1)
    with pytest.raises(ValueError):
        critic(torch.zeros(5))
    with pytest.raises(RuntimeError, match="no CPU path"):
        critic(torch.zeros(2, 5))
    buf = ReplayBuffer(10, 4, 1, 8)
    buf.add_multi(torch.zeros(4, 4), torch.zeros(4, 1), torch.zeros(4), torch.zeros(4, 4), torch.zeros(4), torch.zeros(4, 8))
    with pytest.raises(ValueError):
        buf.sample(5)                                                # replay.py:74-77
    with pytest.raises(RuntimeError, match="no CPU path"):
        buf.sample(2)


def test_deepcopy_and_script_survive():
    """velora/models/nf/modules.py:370-392: networks are deep-copied then torch.jit.script-ed."""
    vb.set_seed(3)
    for net in (LiquidNCPNetwork(4, 20, 2), NCPNetwork(5, 128, 1)):
        twin = torch.jit.script(copy.deepcopy(net))
        assert isinstance(twin, torch.jit.RecursiveScriptModule)
        assert list(twin.state_dict().keys()) == list(net.state_dict().keys())
        assert [tuple(p.shape) for p in twin.parameters()] == [tuple(p.shape) for p in net.parameters()]


def test_layer_modules_formulas(golden):
    """SparseLinear / NCPLiquidCell on their own follow the reference formulas (cell.npz from the reference)."""
    d = golden("cell")
    cell = NCPLiquidCell(8, 6, torch.tensor(d["mask"]))
    assert cell.head_size == 14
    assert np.array_equal(cell.sparsity_mask.numpy(), d["sd_sparsity_mask"])       # _prep_mask, cell.py:91-111
    cell.load_state_dict({k[3:]: torch.tensor(d[k]) for k in d.files if k.startswith("sd_")})
    assert [k for k, _ in cell.named_parameters()] == [k[5:] for k in d.files if k.startswith("grad_")]
    # the cell's arithmetic lives in vb_cell_fwd / vb_cell_bwd (tests/test_liquid_gpu.py checks it against this golden): no CPU path
    with pytest.raises(RuntimeError, match="no CPU path"):
        cell(torch.tensor(d["x"]), torch.tensor(d["h"]))
    # SparseLinear: masked weights are zero after init and the mask is applied in forward (test_lnn.py:429-480)
    mask = torch.tensor([[1, 0, 1], [0, 1, 0]])
    lin = SparseLinear(3, 2, mask)
    assert torch.all(lin.weight[mask == 0] == 0)
    with torch.no_grad():
        lin.weight.fill_(1.0)
        lin.bias.zero_()
    assert torch.equal(lin(torch.ones(1, 3)), torch.tensor([[2.0, 1.0]]))
    lin.update_mask(torch.ones(2, 3, dtype=torch.int64))
    assert torch.equal(lin(torch.ones(1, 3)), torch.tensor([[3.0, 3.0]]))


# ------------------------------------------------------------------------------------------------ buffer (host side)
def test_buffer_ring_semantics_vs_reference(golden):
    d = golden("buffer")
    cap, S, A, H = (int(v) for v in d["dims"])
    buf = ReplayBuffer(cap, S, A, H)
    fi = 0
    while f"feed{fi}_kind" in d.files:
        c = [torch.tensor(d[f"feed{fi}_{ti}"]) for ti in range(6)]
        if str(d[f"feed{fi}_kind"][0]) == "multi":
            buf.add_multi(*c)
        else:
            for r in range(c[0].shape[0]):
                buf.add(c[0][r], c[1][r], float(c[2][r]), c[3][r], bool(c[4][r]), c[5][r])
        fi += 1
    assert [buf.position, buf.size] == list(d["position_size"]) and len(buf) == buf.size
    for k, v in buf.state_dict().items():
        assert np.array_equal(v.numpy(), d[f"state_{k}"]), k
    assert buf.metadata()["capacity"] == cap and buf.config().type == "ReplayBuffer"
    assert BatchExperience.__dataclass_fields__.keys() == {"states", "actions", "rewards", "next_states", "dones", "hiddens"}


def test_buffer_save_load_roundtrip(tmp_path):
    buf = ReplayBuffer(20, 3, 2, 4)
    g = torch.Generator().manual_seed(0)
    buf.add_multi(torch.randn(25, 3, generator=g), torch.randn(25, 2, generator=g), torch.randn(25, generator=g),
                  torch.randn(25, 3, generator=g), torch.zeros(25), torch.randn(25, 4, generator=g))
    buf.save(tmp_path)
    import json

    meta = json.loads((tmp_path / "buffer_metadata.json").read_text())
    again = ReplayBuffer.load(tmp_path / "buffer_state", meta)
    assert (again.position, again.size) == (buf.position, buf.size) == (5, 20)
    for k, v in buf.state_dict().items():
        assert torch.equal(v, again.state_dict()[k])


# ------------------------------------------------------------------------------------------------ C ABI
def test_library_loads_and_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "velora_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(vb_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert os.path.isfile(_lib.LIB_PATH), "run __graft_entry__.build() first"
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(raw, name), name
    lib = _lib.lib()
    abi = int(re.search(r"#define\s+VB_ABI_VERSION\s+(\d+)", header).group(1))
    assert lib.vb_query(0) == abi and lib.vb_query(1) == 100        # ABI version of the header, built for sm_100a
    # pure host-side queries work without a device
    assert lib.vb_liquid_packed_floats(4, 12, 8, 2) == 752 + 48 + 640 + 32
    assert lib.vb_liquid_grad_floats(4, 12, 8, 2) == 752
    assert lib.vb_liquid_has_fast_path(4, 12, 8, 2) == 2 and lib.vb_liquid_has_fast_path(10, 12, 8, 5) == 2 and lib.vb_liquid_has_fast_path(6, 12, 8, 2) == 0 and lib.vb_liquid_has_fast_path(5, 77, 51, 1) == 0
    assert lib.vb_liquid_packed_floats(0, 1, 1, 1) < 0
    # argument errors are reported, not crashed on
    assert lib.vb_liquid_fwd(None, None, None, None, None, 1, 1, 4, 12, 8, 2, None, 0, 0, None) == -1
    assert b"null pointer" in lib.vb_last_error()


def test_graft_entry_build_compiles_and_loads():
    """`__graft_entry__.build()` (the driver's "does it build" check) compiles every CUDA source for sm_100a, loads the library
    and checks its ABI version against the header."""
    import __graft_entry__ as entry

    entry.build()
    assert os.path.isfile(_lib.LIB_PATH)


def test_packed_row_layout_keeps_vector_alignment():
    """Shadow-table row layout: vector-width fields first, every field's offset aligned to its own vector width, stride a
    multiple of 4 floats, nothing overlaps."""
    from velora_b200.functional import packed_row_layout

    for widths in ([4, 1, 1, 4, 1, 8], [3, 2, 1, 3, 1, 8], [17, 6, 1, 17, 1, 40], [1], [8, 8]):
        offsets, stride = packed_row_layout(widths)
        assert stride % 4 == 0 and stride >= sum(widths) and stride - sum(widths) < 4
        spans = sorted((o, o + w) for o, w in zip(offsets, widths))
        assert spans[0][0] == 0 and all(a[1] == b[0] for a, b in zip(spans, spans[1:]))      # contiguous, no overlap
        for o, w in zip(offsets, widths):
            if w % 4 == 0:
                assert o % 4 == 0
            elif w % 2 == 0:
                assert o % 2 == 0
    assert packed_row_layout([4, 1, 1, 4, 1, 8]) == ([0, 16, 17, 4, 18, 8], 20)      # the reference's buffer: 19 floats -> 80-byte rows


def test_common_mask_type_codes():
    """One type code per mask group: transposed views pass as they are, a single-row mask adopts the group's layout, mixed
    layouts or dtypes are reported as such (the caller then copies)."""
    from velora_b200 import _lib

    a = torch.ones(5, 7, dtype=torch.int32).t()           # [7, 5] transposed view, as the reference stores abs(mask.T)
    b = torch.ones(3, 7, dtype=torch.int32).t()
    row = torch.ones(7, 1, dtype=torch.int32).t()          # [1, 7]: both row- and column-major
    assert _lib.mask_type(a) == _lib.VB_MASK_I32 | _lib.VB_MASK_TRANSPOSED
    assert _lib.common_mask_type([a, b, row]) == _lib.VB_MASK_I32 | _lib.VB_MASK_TRANSPOSED
    assert _lib.common_mask_type([a.contiguous(), b.contiguous(), row]) == _lib.VB_MASK_I32
    assert _lib.common_mask_type([a, b.contiguous()]) is None
    assert _lib.common_mask_type([a, b.float()]) is None
    assert _lib.common_mask_type([row]) == _lib.VB_MASK_I32
    f = torch.ones(4, 6).t()
    assert _lib.common_mask_type([f, f]) == _lib.VB_MASK_F32 | _lib.VB_MASK_TRANSPOSED


def test_dataparallel_helpers_single_process():
    from velora_b200 import dataparallel as dp

    assert dp.world_size() == 1 and dp.rank() == 0 and not dp.is_enabled()
    assert dp.shard_bounds(12, rank=0, world=4) == (0, 3) and dp.shard_bounds(12, rank=3, world=4) == (9, 12)
    assert dp.shard_bounds(8, rank=1, world=2) == (4, 8)
    t = torch.ones(3)
    dp.maybe_allreduce(t)                                  # disabled: no-op
    assert torch.equal(t, torch.ones(3))


def test_fused_adam_argument_contract():
    """FusedAdam is the plain Adam the reference constructs: other variants are refused at construction; the optimizer state
    layout is torch's capturable one; without CUDA the step fails loudly instead of falling back."""
    from velora_b200.optim import FusedAdam

    p = torch.nn.Parameter(torch.ones(3))
    for bad in (dict(weight_decay=0.1), dict(amsgrad=True), dict(maximize=True)):
        with pytest.raises(ValueError):
            FusedAdam([p], **bad)
    opt = FusedAdam([p], lr=1e-3)
    assert opt.defaults["capturable"] is True and opt.defaults["weight_decay"] == 0.0
    p.grad = torch.ones(3)
    with pytest.raises(RuntimeError, match="CUDA"):
        opt.step()


def test_state_dicts_round_trip_with_the_reference_classes():
    """Checkpoint compatibility: state dicts of the mirrored networks and SAC heads load into the
    unmodified reference classes with strict=True and back, key for key and bit for bit."""
    from oracle import ref_import

    if not ref_import.reference_available():
        pytest.skip("reference tree not present (GPU box)")
    import velora_b200 as vb
    from velora_b200.models.lnn import LiquidNCPNetwork, NCPNetwork
    from velora_b200.models.sac import SACActor, SACActorDiscrete, SACCriticNCP, SACCriticNCPDiscrete

    ref = ref_import.import_reference()
    import velora.models.sac.continuous as rc
    import velora.models.sac.discrete as rd

    one, zero = torch.tensor([1.0]), torch.tensor([0.0])
    pairs = [
        (lambda: LiquidNCPNetwork(10, 20, 5), lambda: ref.LiquidNCPNetwork(10, 20, 5)),
        (lambda: NCPNetwork(5, 128, 1), lambda: ref.NCPNetwork(5, 128, 1)),
        (lambda: SACActor(4, 20, 1, one, zero), lambda: rc.SACActor(4, 20, 1, one, zero)),
        (lambda: SACCriticNCP(4, 128, 1), lambda: rc.SACCriticNCP(4, 128, 1)),
        (lambda: SACActorDiscrete(4, 20, 2), lambda: rd.SACActorDiscrete(4, 20, 2)),
        (lambda: SACCriticNCPDiscrete(4, 128, 2), lambda: rd.SACCriticNCPDiscrete(4, 128, 2)),
    ]
    for make_ours, make_ref in pairs:
        vb.set_seed(3)
        ours = make_ours()
        ref.set_seed(4)
        theirs = make_ref()
        sd_ours, sd_theirs = ours.state_dict(), theirs.state_dict()
        assert list(sd_ours.keys()) == list(sd_theirs.keys())
        theirs.load_state_dict(sd_ours, strict=True)          # ours -> reference
        for k, v in theirs.state_dict().items():
            assert torch.equal(v, sd_ours[k]), k
        ref.set_seed(5)
        other = make_ref()
        ours.load_state_dict(other.state_dict(), strict=True)  # reference -> ours
        for k, v in ours.state_dict().items():
            assert torch.equal(v, other.state_dict()[k]), k


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the oracle's CPU port timed on the host cores) prints ONE JSON line with the keys the driver
    reads; runs without a GPU."""
    import json
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                        "--batch", "256", "--seq", "4"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "liquid_cell_fwd_bwd_sample_steps_per_s" and d["unit"] == "sample-steps/s"
    for key in ("value", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config",
                "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["value"] > 0 and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]


def test_frozen_parameters_flag_is_scoped_and_thread_local():
    """functional.frozen_parameters(): nests, restores the previous state on exit (also on an exception) and is invisible to other threads
    (autograd's backward thread never sees the caller's flag)."""
    import threading

    from velora_b200 import functional as VF

    assert not VF.parameters_frozen()
    with VF.frozen_parameters():
        assert VF.parameters_frozen()
        with VF.frozen_parameters():
            assert VF.parameters_frozen()
        assert VF.parameters_frozen()
        seen = []
        t = threading.Thread(target=lambda: seen.append(VF.parameters_frozen()))
        t.start()
        t.join()
        assert seen == [False]
    assert not VF.parameters_frozen()
    with pytest.raises(ValueError):
        with VF.frozen_parameters():
            raise ValueError("boom")
    assert not VF.parameters_frozen()


def test_relayout_leaves_small_or_non_contiguous_tensors_alone():
    """functional._relayout transposes only contiguous batch-major tensors of at least RELAYOUT_MIN_ROWS sample-steps, and only when the
    kernel path accepts time-major storage; everything else is passed through untouched (no kernel call: this runs without a GPU)."""
    from velora_b200 import functional as VF

    small = torch.zeros(8, 4, 3)
    assert VF._relayout(small, False, True)[0] is small
    assert VF._relayout(None, False, True) == (None, False)
    big = torch.zeros(VF.RELAYOUT_MIN_ROWS, 2, 3)
    assert VF._relayout(big, False, False)[0] is big                      # the kernel path takes batch-major only
    assert VF._relayout(big, True, True)[0] is big                        # already time-major
    view = torch.zeros(2, VF.RELAYOUT_MIN_ROWS, 3).permute(1, 0, 2)
    assert VF._relayout(view, False, True)[0] is view                     # not contiguous: the caller's _seq() decides

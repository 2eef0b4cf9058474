"""The oracle against the golden vectors produced by the unmodified reference (CPU, no GPU needed)."""
import numpy as np
import pytest
import torch

from oracle import buffer_oracle, liquid_oracle as lo, wiring_oracle
from vb_testutil import liquid_case_params


def _seed(v):
    # what velora/utils/core.py:7-25 (set_seed) does
    import random

    random.seed(v)
    torch.manual_seed(v)
    np.random.seed(v)


def test_wiring_oracle_bit_exact(golden):
    d = golden("wiring")
    for idx, (i, n, o, sp, seed) in enumerate(d["cases"]):
        _seed(int(seed))
        masks, counts, syn = wiring_oracle.wiring(int(i), int(n), int(o), float(sp))
        for k in ("inter", "command", "motor", "recurrent"):
            assert masks[k].dtype == np.int32
            assert np.array_equal(masks[k], d[f"c{idx}_{k}"]), (idx, k)
        assert list(counts) == list(d[f"c{idx}_counts"])
        assert list(syn) == list(d[f"c{idx}_synapses"])
        # same amount of both RNG streams consumed
        assert np.random.randint(0, 2**31 - 1) == int(d[f"c{idx}_next_np"][0])
        assert int(torch.randint(0, 2**31 - 1, (1,))) == int(d[f"c{idx}_next_torch"][0])


def test_wiring_reference_known_answers():
    # reference tests/models/test_wiring.py:52-54: (4,10,2,0.5) -> SynapseCounts(3,2,4,2)
    _, ni, nc, _ = wiring_oracle.neuron_counts(4, 10, 2)
    assert wiring_oracle.synapse_counts(ni, nc, 0.5) == (3, 2, 4, 2)
    with pytest.raises(ValueError):
        wiring_oracle.wiring(4, 10, 2, 0.05)


@pytest.mark.parametrize("name", ["fixture", "actor_t1", "actor_seq", "actor_seq_allh", "fixture_seq", "odd", "mid", "wide"])
def test_liquid_numpy_oracle_vs_reference(golden, name):
    d = golden("liquid")
    p = liquid_case_params(d, name)
    x, h0, ry, rh = d[f"{name}_x"], d[f"{name}_h0"], d[f"{name}_ry"], d[f"{name}_rh"]
    # fp64 oracle against the fp64 deep copy of the reference: agreement to rounding
    y, ht, caches = lo.liquid_seq_forward(p, x, h0, np.float64)
    dx, dh0, grads = lo.liquid_seq_backward(p, caches, ry, rh, np.float64)
    assert lo.rel_err(y, d[f"{name}_f64_y"]) < 1e-13
    assert lo.rel_err(ht, d[f"{name}_f64_htraj"]) < 1e-13
    assert lo.rel_err(dx, d[f"{name}_f64_dx"]) < 1e-12
    assert lo.rel_err(dh0, d[f"{name}_f64_dh0"]) < 1e-12
    for k, g in grads.items():
        assert lo.rel_err(g, d[f"{name}_f64_grad_{k}"]) < 1e-12, k
        mask = p[k.replace("weight", "mask")] if k.endswith("weight") else None
        if mask is not None:
            assert np.all(g[mask == 0] == 0.0)   # reference tests/models/test_lnn.py:570-576
    # fp32 oracle is within fp32 rounding of the fp32 reference
    y32, ht32, c32 = lo.liquid_seq_forward(p, x, h0, np.float32)
    assert lo.rel_err(y32, d[f"{name}_f32_y"]) < 5e-6
    assert lo.rel_err(ht32, d[f"{name}_f32_htraj"]) < 5e-6


@pytest.mark.parametrize("name", ["fixture", "actor_seq", "odd", "wide"])
def test_liquid_torch_port_bit_identical(golden, name):
    """The torch CPU port issues the reference's op sequence: identical bits on the same parameters."""
    d = golden("liquid")
    torch.set_num_threads(1)
    p = lo.torch_params(liquid_case_params(d, name), requires_grad=True)
    x = torch.tensor(d[f"{name}_x"], requires_grad=True)
    h0 = torch.tensor(d[f"{name}_h0"], requires_grad=True)
    y, ht = lo.liquid_seq_forward_torch(p, x, h0)
    assert np.array_equal(y.detach().numpy(), d[f"{name}_f32_y"])
    assert np.array_equal(ht.detach().numpy(), d[f"{name}_f32_htraj"])
    ((y * torch.tensor(d[f"{name}_ry"])).sum() + (ht * torch.tensor(d[f"{name}_rh"])).sum()).backward()
    assert lo.rel_err(x.grad.numpy(), d[f"{name}_f32_dx"]) < 1e-6
    for k, t in p.items():
        if t.requires_grad:
            assert lo.rel_err(t.grad.numpy(), d[f"{name}_f32_grad_{k}"]) < 1e-6, k


def test_liquid_anchor_and_h0_none(golden):
    d = golden("liquid")
    # SURVEY.md 8c anchors for the (10,20,5) fixture; reference tests/models/test_lnn.py:280-309 (1017 / 657)
    assert list(d["anchor_total_active"]) == [1017, 657]
    p = liquid_case_params(d, "fixture")
    n_total = sum(v.size for k, v in p.items() if k.endswith("weight") or k.endswith("bias"))
    n_active = sum(int((v != 0).sum()) for k, v in p.items() if k.endswith("weight") or k.endswith("bias"))
    assert (n_total, n_active) == (1017, 657)
    y, ht, _ = lo.liquid_seq_forward(p, d["fixture_x"], d["fixture_h0"], np.float32)
    assert abs(float(y.sum()) - float(d["anchor_sum_y"][0])) < 1e-5
    assert np.allclose(y[0, 0], d["anchor_y0"], atol=1e-6)
    y0, h1, _ = lo.liquid_seq_forward(p, d["fixture_x"][:, :1], None, np.float32)
    assert lo.rel_err(y0[:, 0], d["fixture_f32_y_h0none"]) < 5e-6
    assert lo.rel_err(h1[:, 0], d["fixture_f32_h_h0none"]) < 5e-6


def test_cell_oracle_vs_reference(golden):
    """NCPLiquidCell alone (cell.py:147-169): drive the oracle's cell with identity inter/motor stripped."""
    d = golden("cell")
    mask = d["mask"]
    # _prep_mask: abs(cat([mask, ones]).T)  (cell.py:91-111)
    prep = np.abs(np.concatenate([mask, np.ones((mask.shape[1],) * 2, mask.dtype)]).T).astype(np.float32)
    assert np.array_equal(prep, d["sd_sparsity_mask"])
    x, h = d["x"].astype(np.float64), d["h"].astype(np.float64)
    z = np.concatenate([x, h], 1)
    s = {k: lo.masked_linear(z, d[f"sd_{k}.weight"].astype(np.float64), d[f"sd_{k}.mask"], d[f"sd_{k}.bias"].astype(np.float64)) for k in lo.HEADS}
    gate = lo.sigmoid(s["f_head_to_g"] + s["f_head_to_h"])
    hn = np.tanh(s["g_head"]) * (1 - gate) + gate * np.tanh(s["h_head"])
    y = s["proj"] + hn
    assert lo.rel_err(hn, d["hn"]) < 2e-6
    assert lo.rel_err(y, d["y"]) < 2e-6


def test_gating_formula_known_answer():
    # reference tests/models/test_lnn.py:103-138: f heads mocked to 0.5 and 0.3 -> gate = sigmoid(0.8)
    g_out, h_out = np.array([[0.2, -1.0]]), np.array([[1.5, 0.1]])
    gate = lo.sigmoid(np.array([[0.8, 0.8]]))
    expect = np.tanh(g_out) * (1 - gate) + gate * np.tanh(h_out)
    assert np.allclose(expect, np.tanh(g_out) * (1 - 1 / (1 + np.exp(-0.8))) + np.tanh(h_out) / (1 + np.exp(-0.8)))


@pytest.mark.parametrize("name", ["critic_ct", "critic_disc", "small", "odd"])
def test_ncp_oracle_vs_reference(golden, name):
    d = golden("ncp")
    p = liquid_case_params(d, name)
    x, r = d[f"{name}_x"], d[f"{name}_r"]
    y, cache = lo.ncp_forward(p, x, np.float64)
    dx, grads = lo.ncp_backward(p, cache, r, np.float64)
    assert lo.rel_err(y, d[f"{name}_f64_y"]) < 1e-13
    assert lo.rel_err(dx, d[f"{name}_f64_dx"]) < 1e-12
    for k, g in grads.items():
        assert lo.rel_err(g, d[f"{name}_f64_grad_{k}"]) < 1e-12, k
    pt = lo.torch_params(p)
    torch.set_num_threads(1)
    assert np.array_equal(lo.ncp_forward_torch(pt, torch.tensor(x)).numpy(), d[f"{name}_f32_y"])


def test_buffer_oracle_vs_reference(golden):
    d = golden("buffer")
    cap, S, A, H = (int(v) for v in d["dims"])
    buf = buffer_oracle.RingBufferOracle(cap, S, A, H)
    fi = 0
    while f"feed{fi}_kind" in d.files:
        c = [d[f"feed{fi}_{ti}"] for ti in range(6)]
        if str(d[f"feed{fi}_kind"][0]) == "multi":
            buf.add_multi(*c)
        else:
            for r in range(c[0].shape[0]):
                buf.add(c[0][r], c[1][r], float(c[2][r]), c[3][r], bool(c[4][r]), c[5][r])
        fi += 1
    assert [buf.position, buf.size] == list(d["position_size"])
    for k in buffer_oracle.FIELDS:
        assert np.array_equal(getattr(buf, k), d[f"state_{k}"]), k
    for B, seed in ((16, 1), (50, 2)):
        torch.manual_seed(seed)
        idx = torch.randint(0, buf.size, (B,)).numpy()      # replay.py:79
        assert np.array_equal(idx, d[f"sample{B}_idx"])
        out = buf.gather(idx)
        for k in buffer_oracle.FIELDS:
            assert np.array_equal(out[k], d[f"sample{B}_{k}"]), k
    with pytest.raises(ValueError):
        buf.gather(np.zeros(cap + 1, np.int64))


def test_train_oracle_vs_reference(golden):
    """The CPU restatement of the NeuroFlowCT train step reproduces three reference `_train_step` calls."""
    from oracle.train_oracle import TrainOracle

    d = golden("train")
    pre = lambda p: {k[len(p):]: d[k] for k in d.files if k.startswith(p)}   # noqa: E731
    torch.set_num_threads(1)
    o = TrainOracle(pre("init_actor_"), pre("init_critic1_"), pre("init_critic2_"))
    eps, idx = torch.tensor(d["eps"]), d["idx"]
    for s in range(int(d["dims"][3])):
        batch = {k: torch.tensor(d[f"buf_{k}"][idx[s]]) for k in ("states", "actions", "rewards", "next_states", "dones", "hiddens")}
        losses = o.train_step(batch, [eps[2 * s], eps[2 * s + 1]])
        assert np.allclose(losses, d["losses"][s], rtol=1e-5, atol=1e-6), (s, losses, d["losses"][s])
    for k, v in pre("final_actor_ncp.").items():
        assert np.allclose(o.actor[k].detach().numpy(), v, atol=1e-6), k
    assert abs(float(o.log_alpha) - float(d["final_log_alpha"])) < 1e-6


def test_train_discrete_golden_regenerates(golden):
    """tests/golden/train_discrete.npz is what the unmodified reference's discrete agent produces here (no oracle port of
    that step exists: the GPU parity test compares NeuroFlowCore with these reference outputs directly)."""
    from oracle import make_golden, ref_import

    if not ref_import.reference_available():
        pytest.skip("reference tree not present (GPU box)")

    d = golden("train_discrete")
    n_threads = torch.get_num_threads()
    torch.set_num_threads(1)
    try:
        fresh = make_golden.gen_train_discrete(ref_import.import_reference())
    finally:
        torch.set_num_threads(n_threads)
    assert set(fresh) == set(d.keys())
    for k, v in fresh.items():
        assert np.array_equal(np.asarray(v), d[k]), k


def test_bf16_rounding_helper_matches_torch():
    a = (np.random.default_rng(0).standard_normal(4096) * np.array([1e-3, 1.0, 300.0, 1e6]).repeat(1024)).astype(np.float32)
    assert np.array_equal(lo.bf16_round(a), torch.tensor(a).to(torch.bfloat16).float().numpy())


def test_bf16_contract_oracle_against_the_fp64_oracle_and_the_reference(golden):
    """`liquid_seq_forward_bf16` (the arithmetic contract of the tcgen05 bf16 forward: bf16-rounded head operands, exact
    accumulation) stays within BASELINE's 2e-2 of the fp64 oracle and of the REFERENCE's fp64 deep copy at (8,512,1); the fp64
    oracle itself reproduces the reference at that shape to 1e-12."""
    from vb_testutil import liquid_case_params, rel_err

    d = golden("liquid_big")
    p = liquid_case_params(d, "big")
    x, h0 = d["big_x"], d["big_h0"]
    y64, h64, caches = lo.liquid_seq_forward(p, x, h0, np.float64)
    assert rel_err(y64, d["big_f64_y"]) < 1e-12 and rel_err(h64, d["big_f64_htraj"]) < 1e-12
    dx, dh0, grads = lo.liquid_seq_backward(p, caches, d["big_ry"], d["big_rh"], np.float64)
    assert rel_err(dx, d["big_f64_dx"]) < 1e-11 and rel_err(dh0, d["big_f64_dh0"]) < 1e-11
    for k, g in grads.items():
        assert rel_err(g, d[f"big_f64_grad_{k}"]) < 1e-11, k
    yb, hb = lo.liquid_seq_forward_bf16(p, x, h0)
    assert rel_err(yb, y64) < 2e-2 and rel_err(hb, h64) < 2e-2
    assert rel_err(yb, d["big_f64_y"]) < 2e-2 and rel_err(hb, d["big_f64_htraj"]) < 2e-2

"""ReplayBuffer on the GPU: the fused gather is bit-exact with torch indexing / the reference goldens."""
import numpy as np
import pytest
import torch

from oracle import buffer_oracle

pytestmark = pytest.mark.gpu
FIELDS = buffer_oracle.FIELDS


def _replay(d):
    from velora_b200.buffer import ReplayBuffer

    cap, S, A, H = (int(v) for v in d["dims"])
    buf = ReplayBuffer(cap, S, A, H, device=torch.device("cuda"))
    fi = 0
    while f"feed{fi}_kind" in d.files:
        c = [torch.tensor(d[f"feed{fi}_{ti}"]) for ti in range(6)]
        if str(d[f"feed{fi}_kind"][0]) == "multi":
            buf.add_multi(*[t.cuda() for t in c])
        else:
            for r in range(c[0].shape[0]):
                buf.add(c[0][r].cuda(), c[1][r].cuda(), float(c[2][r]), c[3][r].cuda(), bool(c[4][r]), c[5][r].cuda())
        fi += 1
    return buf


def test_ring_writes_and_gather_vs_reference(golden):
    d = golden("buffer")
    buf = _replay(d)
    assert [buf.position, buf.size] == list(d["position_size"])
    for k in FIELDS:
        assert np.array_equal(getattr(buf, k).cpu().numpy(), d[f"state_{k}"]), k
    for B in (16, 50):
        out = buf.gather(torch.tensor(d[f"sample{B}_idx"], device="cuda"))
        for k in FIELDS:
            assert np.array_equal(getattr(out, k).cpu().numpy(), d[f"sample{B}_{k}"]), k


def test_sample_indices_bit_exact_with_torch_on_device(golden):
    """replay.py:79: indices come from torch.randint on the buffer's device; same seed => same rows."""
    d = golden("buffer")
    buf = _replay(d)
    for B in (1, 7, 50):
        torch.manual_seed(123)
        batch = buf.sample(B)
        torch.manual_seed(123)
        idx = torch.randint(0, buf.size, (B,), device="cuda")
        for k in FIELDS:
            assert torch.equal(getattr(batch, k), getattr(buf, k)[idx]), k
    with pytest.raises(ValueError):
        buf.sample(buf.size + 1)


@pytest.mark.parametrize("S,A,H", [(4, 1, 8), (3, 2, 5), (17, 6, 24)])
def test_gather_full_size_vs_oracle(S, A, H):
    """BASELINE config 5 shape: 1M-transition buffer, 1M sampled rows; checked against torch indexing on the
    device and (on a slice) the numpy oracle, plus a checksum property."""
    from velora_b200.buffer import ReplayBuffer

    cap = 1_000_000 if S == 4 else 50_000
    buf = ReplayBuffer(cap, S, A, H, device=torch.device("cuda"))
    g = torch.Generator(device="cuda").manual_seed(0)
    for k in FIELDS:
        getattr(buf, k).copy_(torch.randn(getattr(buf, k).shape, device="cuda", generator=g))
    buf.size = cap
    B = 1_048_576 if S == 4 else 70_001
    torch.manual_seed(1)
    idx = torch.randint(0, cap, (B,), device="cuda")
    out = buf.gather(idx)
    for k in FIELDS:
        assert torch.equal(getattr(out, k), getattr(buf, k)[idx]), k
    sl = idx[:4096].cpu().numpy()
    oracle = buffer_oracle.gather6([getattr(buf, k).cpu().numpy() for k in FIELDS], sl)
    for k, o in zip(FIELDS, oracle):
        assert np.array_equal(getattr(out, k)[:4096].cpu().numpy(), o), k
    # gathering the identity permutation returns the buffer itself
    ident = buf.gather(torch.arange(cap, device="cuda"))
    for k in FIELDS:
        assert torch.equal(getattr(ident, k), getattr(buf, k))


def test_shadow_table_tracks_every_write_path():
    """The packed shadow table the gather reads stays identical to torch indexing of the six arrays through add, add_multi
    (with wrap-around), direct tensor writes and state reassignment; use_shadow=False reads the arrays themselves."""
    from velora_b200.buffer import ReplayBuffer

    dev = torch.device("cuda")
    g = torch.Generator(device=dev).manual_seed(0)
    buf = ReplayBuffer(100, 4, 1, 8, device=dev)

    def check():
        idx = torch.randint(0, buf.size, (257,), device=dev, generator=g)
        for shadow in (True, False):
            buf.use_shadow = shadow
            got = buf.gather(idx)
            for name in ("states", "actions", "rewards", "next_states", "dones", "hiddens"):
                assert torch.equal(getattr(got, name), getattr(buf, name)[idx]), (name, shadow)
        buf.use_shadow = True

    def rnd(*shape):
        return torch.randn(*shape, device=dev, generator=g)

    buf.add_multi(rnd(30, 4), rnd(30, 1), rnd(30), rnd(30, 4), (rnd(30) > 0), rnd(30, 8))
    check()                                      # first gather builds the table
    assert buf._shadow_current()
    for _ in range(5):
        buf.add(rnd(4), rnd(1), 1.5, rnd(4), False, rnd(8))
    assert buf._shadow_current()                 # row-wise refresh, no rebuild
    check()
    buf.add_multi(rnd(90, 4), rnd(90, 1), rnd(90), rnd(90, 4), (rnd(90) > 0), rnd(90, 8))   # wraps around the ring
    assert buf.size == 100 and buf._shadow_current()
    check()
    buf.states.mul_(2.0)                         # written behind the table's back
    assert not buf._shadow_current()
    check()
    buf.hiddens = buf.hiddens.clone() + 1.0      # reassigned (what load() does)
    check()


def test_packed_gather_bit_exact_at_scale():
    from velora_b200.buffer import ReplayBuffer

    dev = torch.device("cuda")
    g = torch.Generator(device=dev).manual_seed(1)
    buf = ReplayBuffer(1_000_000, 4, 1, 8, device=dev)
    for name in ("states", "actions", "rewards", "next_states", "dones", "hiddens"):
        getattr(buf, name).copy_(torch.randn(getattr(buf, name).shape, device=dev, generator=g))
    buf.size = buf.capacity
    idx = torch.randint(0, buf.size, (1_000_000,), device=dev, generator=g)
    got = buf.gather(idx)
    for name in ("states", "actions", "rewards", "next_states", "dones", "hiddens"):
        assert torch.equal(getattr(got, name), getattr(buf, name)[idx]), name


def test_gather_empty_index_vector():
    from velora_b200.buffer import ReplayBuffer

    buf = ReplayBuffer(16, 4, 1, 8, device=torch.device("cuda"))
    buf.size = 16
    for shadow in (True, False):
        buf.use_shadow = shadow
        got = buf.gather(torch.zeros(0, dtype=torch.int64, device="cuda"))
        assert got.states.shape == (0, 4) and got.hiddens.shape == (0, 8) and got.rewards.shape == (0, 1)


def test_add_is_one_launch_and_equals_the_indexing_writes():
    """BufferBase.add (base.py:73-101) through vb_buffer_add_row: same array contents as the reference's six indexing assignments for
    every value kind the agents pass (float32 rows, an int64 action index, 0-dim tensors, Python numbers, host scalar tensors), wrap-around
    included; values on another device fall back to the indexing writes."""
    from velora_b200 import functional as VF
    from velora_b200.buffer import ReplayBuffer

    dev = torch.device("cuda")
    g = torch.Generator(device=dev).manual_seed(3)
    ours = ReplayBuffer(7, 4, 1, 8, device=dev)
    ref = {k: torch.zeros_like(getattr(ours, k)) for k in FIELDS}
    pos = 0
    for n in range(17):
        s, s2, h = (torch.randn(w, device=dev, generator=g) for w in (4, 4, 8))
        kind = n % 4
        action = [torch.randn(1, device=dev, generator=g), torch.tensor(n % 3, device=dev), torch.tensor([n % 3], device=dev),
                  torch.randn((), device=dev, generator=g)][kind]
        reward = [1.5, torch.tensor(0.25 * n), float(n), torch.tensor(2.0, device=dev)][kind]
        done = [False, True, torch.tensor(True), torch.tensor(0.0, device=dev)][kind]
        before = VF.LAUNCHES
        ours.add(s, action, reward, s2, done, h)
        assert VF.LAUNCHES - before == 1 + (1 if action.dtype != torch.float32 else 0) * 0      # one kernel of ours (dtype casts are torch's)
        ref["states"][pos] = s
        ref["actions"][pos] = action
        ref["rewards"][pos] = reward
        ref["next_states"][pos] = s2
        ref["dones"][pos] = done
        ref["hiddens"][pos] = h
        pos = (pos + 1) % 7
        if n == 9:
            ours.gather(torch.arange(ours.size, device=dev))      # builds the shadow table: later adds must keep it current
    assert ours.position == pos and ours.size == 7 and ours._shadow_current()
    for k in FIELDS:
        assert torch.equal(getattr(ours, k), ref[k]), k
    got = ours.gather(torch.arange(7, device=dev))
    for k in FIELDS:
        assert torch.equal(getattr(got, k), ref[k]), k
    # CPU rows: the reference's own indexing writes (they copy across devices)
    before = VF.LAUNCHES
    ours.add(torch.ones(4), torch.ones(1), 0.0, torch.ones(4), False, torch.ones(8))
    assert torch.equal(ours.states[pos], torch.ones(4, device=dev)) and ours._shadow_current()

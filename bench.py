"""Benchmark of the liquid-cell hot path (BASELINE.json metric: liquid-cell fwd+bwd sample-steps/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE config 3): set_seed(64); LiquidNCPNetwork(4, 20, 2); x ~ N(0,1) [65536, 32, 4]
fp32 per GPU, h0 = None; one step = forward_seq + backward of  sum(y * r_y) + sum(h_T * r_h)  (r ~ N(0,1)),
including the parameter-gradient reduction (and, for N > 1, the one NCCL all-reduce of the flat gradient block).
Weak scaling: every rank owns 65536 batch rows.

One JSON line on stdout (rank 0).  `value` is measured with inputs resident in HBM; `e2e` goes through the same
public API with pinned HOST inputs, H2D/D2H copies inside the timed region; `roofline` is the dominant (backward)
kernel against the measured HBM peak; `cpu_baseline` is the UNMODIFIED reference (its package copied to the git-ignored
baseline/_ref/ by baseline/install_ref.py, which gpurun ships) timed on this box's host cores - or, when that copy is absent, the
oracle's torch-CPU port of it (kind "port").  `--impl reference` times that alone and, on a GPU box, also launches the
same reference modules eagerly / jit-scripted on the GPU (`reference_on_this_gpu`: the incumbent a GPU user has today).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "liquid_cell_fwd_bwd_sample_steps_per_s"
UNIT = "sample-steps/s"
SHAPE = dict(I=4, N=20, O=2, Ni=12, Nc=8)
ALGO_BYTES_FWD = 4 * (SHAPE["I"] + SHAPE["O"] + SHAPE["Nc"])            # read x_t, write y_t, write h_t
ALGO_BYTES_BWD = 4 * (2 * SHAPE["I"] + SHAPE["O"] + SHAPE["Nc"])        # read x_t, h_{t-1}, dy_t; write dx_t
SAVED_BYTES = 4 * (5 * SHAPE["Nc"])                                      # round 2: gate factors + mish(c) the forward saves for the backward


def workload_config(B: int, T: int) -> dict:
    """Identical on both arms (the driver compares the two `config` objects)."""
    return {"workload": f"LiquidNCPNetwork(4,20,2) fwd+bwd, batch {B} x seq {T} per GPU, fp32 (BASELINE config 3)",
            "batch_per_gpu": B, "seq_len": T}


def algo_bytes_step(T: int) -> float:
    return ALGO_BYTES_FWD + ALGO_BYTES_BWD + 12.0 * SHAPE["Nc"] / T      # + h0/dh_T/dh0 once per sequence (131 B @ T=32)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed regions run."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        self.active = False
        self.windows = []          # (t0, t1) host times of the timed regions; only samples inside them are reported
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self.stop_flag:
            if self.active:
                try:
                    self.samples.append((time.perf_counter(), nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                    try:
                        mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                    except Exception:
                        mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                    now = time.perf_counter()
                    for bit, name in names.items():
                        if mask & bit:
                            self.reasons.add((now, name))
                except Exception:
                    pass
            time.sleep(0.004)

    def summary(self):
        inside = lambda t: any(a <= t <= b for a, b in self.windows)   # noqa: E731
        s = sorted(v for t, v in self.samples if inside(t))
        return {
            "sm_mhz": s[len(s) // 2] if s else None,
            "sm_max_mhz": self.max_mhz,
            "reasons": sorted({name for t, name in self.reasons if inside(t)}),
            "samples": len(s),
        }


def bind_to_gpu_numa_node(index: int):
    """Pin this rank's host threads to the CPUs NVML reports as local to its GPU, BEFORE the pinned staging buffers are allocated
    (first touch then places them on that NUMA node).  Round 1's end-to-end leg scaled 4.1x at 8 GPUs: eight ranks pulled 33.5 MB per
    step through whatever socket their pages happened to be on.  Returns the CPU list (None when NVML gives no answer)."""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        cpus = [64 * w + b for w, mask in enumerate(words) for b in range(64) if (mask >> b) & 1 and 64 * w + b < n_cpu]
        if cpus:
            os.sched_setaffinity(0, cpus)
            return cpus
    except Exception:
        pass
    return None


def cpu_port_step_fn(B: int, T: int, threads: int):
    """The oracle's torch-CPU port of the reference (op for op the reference's ATen sequence), fwd + bwd."""
    import torch

    import velora_b200 as vb
    from oracle import liquid_oracle as lo
    from velora_b200.models.lnn import LiquidNCPNetwork

    torch.set_num_threads(threads)
    vb.set_seed(64)
    net = LiquidNCPNetwork(SHAPE["I"], SHAPE["N"], SHAPE["O"])      # CPU parameter container only
    params = lo.torch_params({k: v.detach() for k, v in net.state_dict().items()}, requires_grad=True)
    g = torch.Generator().manual_seed(1234)
    x = torch.randn(B, T, SHAPE["I"], generator=g)
    ry = torch.randn(B, T, SHAPE["O"], generator=g)
    rh = torch.randn(B, SHAPE["Nc"], generator=g)
    leaves = [t for t in params.values() if t.requires_grad]

    def step():
        for t in leaves:
            t.grad = None
        y, ht = lo.liquid_seq_forward_torch(params, x, None)
        ((y * ry).sum() + (ht[:, -1] * rh).sum()).backward()

    return step


def reference_step_fn(B: int, T: int, threads: int, device=None, script: bool = False):
    """The UNMODIFIED reference (baseline/_ref, loaded by oracle/ref_import.py): LiquidNCPNetwork(4,20,2) stepped T times the way
    its callers do (`y, h = net(x_t, h)`, ncp.py:129-178), fwd + bwd of the config-3 loss.  None when no reference tree travelled."""
    import torch

    from oracle import ref_import

    if not ref_import.reference_available():
        return None
    ref = ref_import.import_reference()
    torch.set_num_threads(threads)
    ref.set_seed(64)
    net = ref.LiquidNCPNetwork(SHAPE["I"], SHAPE["N"], SHAPE["O"], device=device)
    if device is not None:
        net = net.to(device)
    params = list(net.parameters())
    if script:
        net = torch.jit.script(net)      # what the reference's modules do with their networks (modules.py:388-392)
    g = torch.Generator().manual_seed(1234)
    x = torch.randn(B, T, SHAPE["I"], generator=g).to(device)
    ry = torch.randn(B, T, SHAPE["O"], generator=g).to(device)
    rh = torch.randn(B, SHAPE["Nc"], generator=g).to(device)

    def step():
        for p in params:
            p.grad = None
        h, loss = None, None
        for t in range(T):
            y, h = net(x[:, t], h)
            term = (y * ry[:, t]).sum()
            loss = term if loss is None else loss + term
        (loss + (h * rh).sum()).backward()

    return step


def time_cpu(B: int, T: int, steps: int, warmup: int):
    """(times, threads, torch threads, kind): the unmodified reference when it is present, else the oracle's port of it."""
    import torch

    threads = os.cpu_count() or 1
    step = reference_step_fn(B, T, threads)
    kind = "reference"
    if step is None:
        step, kind = cpu_port_step_fn(B, T, threads), "port"
    for _ in range(warmup):
        step()
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
    return times, threads, torch.get_num_threads(), kind


def cpu_sample_text(kind: str, B: int, T: int, what: str, used: int) -> str:
    if kind == "reference":
        from oracle import ref_import

        src = f"unmodified reference modules from {os.path.relpath(ref_import.REFERENCE_ROOT, ROOT)} (LiquidNCPNetwork stepped T times, eager)"
    else:
        src = "oracle torch-CPU port of the reference ops"
    return f"{src}, batch {B} x seq {T}, {what}, {used} threads"


def time_reference_on_gpu(B: int, T: int, iters: int = 3):
    """The incumbent a GPU user of the reference has today: the same unmodified modules launched eagerly (and jit-scripted) on
    this GPU.  Reported beside the CPU arm; not the driver's ratio."""
    import torch

    if not torch.cuda.is_available():
        return None
    out = {}
    for name, script in (("eager", False), ("jit_script", True)):
        try:
            step = reference_step_fn(B, T, os.cpu_count() or 1, device=torch.device("cuda", 0), script=script)
            if step is None:
                return None
            for _ in range(3 if script else 1):
                step()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(iters):
                step()
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / iters
            out[name] = {"value": B * T / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms}
        except Exception as e:      # the incumbent is informational: never fail the arm on it
            out[name] = {"error": repr(e)[:200]}
    out["sample"] = f"unmodified reference LiquidNCPNetwork(4,20,2) on cuda:0, batch {B} x seq {T}, fwd+bwd, {iters} steps"
    return out


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    B, T = args.batch, args.seq
    # bounded sample: full batch is ~0.2-1 s per step on the host cores; shrink it when the requested run would take minutes
    sample_B = B
    if (args.steps + args.warmup) > 60:
        sample_B = max(4096, B // 4)
    times, threads, used, kind = time_cpu(sample_B, T, args.steps, max(args.warmup, 1))
    total = sum(times)
    value = sample_B * T * len(times) / total
    line = {
        "impl": "reference",
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": workload_config(B, T),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind,
                         "sample": cpu_sample_text(kind, sample_B, T, f"{len(times)} steps", used)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if kind == "reference" and not args.no_gpu_incumbent:
        inc = time_reference_on_gpu(B, T)
        if inc is not None:
            line["reference_on_this_gpu"] = inc
    print(json.dumps(line), flush=True)


def run_ours(args) -> None:
    import torch
    import torch.distributed as dist

    import velora_b200 as vb
    from velora_b200 import functional as VF
    from velora_b200.models.lnn import LiquidNCPNetwork

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (velora_b200 has no CPU fallback)")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = bind_to_gpu_numa_node(local) if world > 1 else None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        vb.dataparallel.configure(True, reduce="sum", oneshot=False if args.nccl_allreduce else None)
    B, T = args.batch, args.seq
    K, W = args.steps, max(args.warmup, 3)

    vb.set_seed(64)                                    # same seed on every rank => identical masks and parameters
    net = LiquidNCPNetwork(SHAPE["I"], SHAPE["N"], SHAPE["O"], device=dev)
    params = list(net.parameters())
    # R rotating input sets (> L2: 4 x 52 MB of inputs, ~170 MB touched per step).  Sequence tensors are stored time-major
    # ([T,B,F]: the stack of the per-step batches the reference's forward consumes) and handed over as [B,T,F] views.
    R = 4
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    tm = not args.batch_major

    def seq_tensor(F):
        if tm:
            return torch.randn(T, B, F, device=dev, generator=g).permute(1, 0, 2)
        return torch.randn(B, T, F, device=dev, generator=g)

    xs = [seq_tensor(SHAPE["I"]) for _ in range(R)]
    rys = [seq_tensor(SHAPE["O"]) for _ in range(R)]
    rhs = [torch.randn(B, SHAPE["Nc"], device=dev, generator=g) for _ in range(R)]

    def step(i: int) -> None:
        for p in params:
            p.grad = None
        y, _, h_last = net.forward_seq(xs[i % R], None, return_last=True)
        torch.autograd.backward([y, h_last], [rys[i % R], rhs[i % R]])   # d/dy, d/dh_T of sum(y r_y) + sum(h_T r_h)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    tick = torch.zeros(1, device=dev)

    def align():
        """After the host-side barrier the ranks' Python threads resume up to milliseconds apart (8 processes share the host).
        A tiny all-reduce enqueued right before the start event makes every GPU reach its start event when the LAST rank
        has arrived, so that this host skew is not billed to the first of the K timed steps.  Outside the timed region."""
        if world > 1:
            dist.all_reduce(tick)

    sampler = ClockSampler(local)
    sampler.start()
    for i in range(W):
        step(i)
    # ------------------------------------------------------------------ eager pass: per-kernel times (events around each C-ABI call)
    VF.PROFILE = {}
    launches0 = VF.LAUNCHES
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(K):
        step(W + i)
    e1.record()
    barrier()
    gpu_launches = VF.LAUNCHES - launches0
    prof, VF.PROFILE = VF.PROFILE, None
    eager_ms = e0.elapsed_time(e1)
    # ------------------------------------------------------------------ device-resident timing (the headline `value`)
    # One CUDA graph per input set (pack, forward, backward + partial-sum reduction, unpack and the framework's small ops),
    # replayed round-robin: at 0.4 ms of kernels per step the Python dispatch of the eager loop is as long as the GPU work.
    graphs = None
    if not args.no_graph:
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for i in range(R):
                step(i)
        torch.cuda.current_stream(dev).wait_stream(side)
        barrier()
        graphs, pool = [], None
        try:
            for i in range(R):
                gr = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gr, pool=pool):
                    step(i)
                pool = gr.pool()
                graphs.append(gr)
        except Exception as e:       # a failed capture must not cost the measurement: time the eager loop instead
            if world > 1:
                raise                # ranks must not diverge; rerun with --no-graph
            print(f"[bench] CUDA graph capture failed ({e!r}); timing the eager loop", file=sys.stderr)
            graphs = None
            torch.cuda.synchronize()

    def timed_step(i: int) -> None:
        if graphs is None:
            step(i)
        else:
            graphs[i % R].replay()

    # Warm-up of the timed code path itself.  With several ranks the first few dozen replays of a graph that contains the
    # all-reduce run ~15-40 % slower than steady state (measured at 8 GPUs: 0.43-0.54 ms for the first 50 steps, 0.371 ms
    # for every later block of 50), so the multi-rank run gets extra untimed replays on top of the W requested ones.
    extra_warm = 100 if (world > 1 and graphs is not None) else 0
    # NVML polling starts during the warm-up: the first queries of a process take milliseconds inside the driver and, with
    # 8 ranks in lock step, showed up as 0.43-0.81 ms steps when they fell into the timed region
    sampler.active = True
    for i in range(W + extra_warm):
        timed_step(i)
    barrier()
    align()
    w0 = time.perf_counter()
    e0.record()
    for i in range(K):
        timed_step(W + i)
    e1.record()
    barrier()
    sampler.windows.append((w0, time.perf_counter()))
    sampler.active = False
    ms_total = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms_total, eager_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total, eager_ms = float(t[0].item()), float(t[1].item())
    value = world * B * T * K / (ms_total * 1e-3)

    if os.environ.get("VB_BENCH_DIAG"):      # diagnostics for the multi-GPU step time (stderr, rank 0): not part of the result
        def region(n_steps, sample):
            barrier()
            sampler.active = sample
            align()
            e0.record()
            for i in range(n_steps):
                timed_step(W + i)
            e1.record()
            barrier()
            sampler.active = False
            t = torch.tensor([e0.elapsed_time(e1) / n_steps], device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        diag = {"K_nvml_on": region(K, True), "K_nvml_off": region(K, False), "4K_nvml_off": region(4 * K, False),
                "K_nvml_off_again": region(K, False)}
        if world > 1 and graphs is not None:      # the same graphs without the all-reduce: ranks run unsynchronised
            vb.dataparallel.configure(False)
            g2, pool2 = [], None
            for i in range(R):
                gr = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gr, pool=pool2):
                    step(i)
                pool2 = gr.pool()
                g2.append(gr)
            vb.dataparallel.configure(True, reduce="sum")
            nodp = []
            for rep in range(2):
                barrier(); align(); e0.record()
                for i in range(K):
                    g2[i % R].replay()
                e1.record(); barrier()
                t = torch.tensor([e0.elapsed_time(e1) / K], device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                nodp.append(float(t.item()))
            diag["graph_without_allreduce_max_over_ranks"] = nodp
            del g2
        if rank == 0:
            print("[bench diag] ms/step:", json.dumps(diag), file=sys.stderr, flush=True)

    # ------------------------------------------------------------------ the same loop with the sequence tensors stored batch-major
    # ([B,T,F] contiguous: what a caller holding per-sample sequences passes); `value` above is on time-major storage
    bm = None
    if tm and graphs is not None and not args.no_batch_major:
        xs_bm = [t.contiguous() for t in xs]
        rys_bm = [t.contiguous() for t in rys]

        def step_bm(i: int) -> None:
            for p in params:
                p.grad = None
            y, _, h_last = net.forward_seq(xs_bm[i % R], None, return_last=True)
            torch.autograd.backward([y, h_last], [rys_bm[i % R], rhs[i % R]])

        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for i in range(R):
                step_bm(i)
        torch.cuda.current_stream(dev).wait_stream(side)
        barrier()
        g_bm, pool_bm = [], None
        for i in range(R):
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr, pool=pool_bm):
                step_bm(i)
            pool_bm = gr.pool()
            g_bm.append(gr)
        for i in range(W + extra_warm):
            g_bm[i % R].replay()
        barrier()
        align()
        e0.record()
        for i in range(K):
            g_bm[(W + i) % R].replay()
        e1.record()
        barrier()
        t_bm = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(t_bm, op=dist.ReduceOp.MAX)
        bm = {"value": world * B * T * K / (float(t_bm.item()) * 1e-3), "unit": UNIT, "ms_per_step": float(t_bm.item()) / K}
        del g_bm, xs_bm, rys_bm

    def avg_ms(name):
        ev = prof.get(name, [])
        return sum(a.elapsed_time(b) for a, b in ev) / max(len(ev), 1)

    eager_bwd_ms, eager_fwd_ms = avg_ms("liquid_bwd"), avg_ms("liquid_fwd")

    # Kernel durations for the roofline: the two C-ABI entry points launched back to back (20 launches between one pair of
    # events on the launching stream), rotating over the R input sets.  The per-call events of the eager pass above include
    # the host gap before each launch and over-state a 0.1-0.3 ms kernel.
    def direct_kernel_ms(iters: int = 20):
        from velora_b200 import _lib
        L = _lib.lib()
        I, Ni, Nc, O = SHAPE["I"], SHAPE["Ni"], SHAPE["Nc"], SHAPE["O"]
        saved = []
        for i in range(R):
            y, _, h_last = net.forward_seq(xs[i], None, return_last=True)
            x_s, _, h_traj, packed, gates = y.grad_fn.saved_tensors
            saved.append((x_s, h_traj, packed, torch.empty_like(y), torch.empty_like(h_traj), torch.empty_like(x_s), gates))
        dpacked = torch.empty(L.vb_liquid_grad_floats(I, Ni, Nc, O), dtype=torch.float32, device=dev)
        ws = torch.empty(L.vb_liquid_bwd_workspace_bytes(I, Ni, Nc, O, B, T, 1), dtype=torch.uint8, device=dev)
        tm_out = not saved[0][1].is_contiguous()
        lay_f = (VF.X_TM if tm else 0) | ((VF.Y_TM | VF.H_TM) if tm_out else 0)
        lay_b = (VF.X_TM | VF.DX_TM | VF.Y_TM if tm else 0) | (VF.H_TM if tm_out else 0)
        st = _lib.stream_ptr(dev)
        out = []
        for which in ("fwd", "bwd"):
            def call(i):
                x_s, h_traj, packed, y2, h2, dx, gates = saved[i % R]
                if which == "fwd":          # the training forward: it also writes the saved gate factors
                    _lib.check(L.vb_liquid_fwd(_lib.ptr(packed), _lib.ptr(x_s), None, _lib.ptr(y2), _lib.ptr(h2), B, T, I, Ni, Nc, O,
                                               _lib.ptr(gates), gates.numel(), lay_f, st), "vb_liquid_fwd")
                else:
                    _lib.check(L.vb_liquid_bwd(_lib.ptr(packed), _lib.ptr(x_s), None, _lib.ptr(h_traj), _lib.ptr(rys[i % R]), None,
                                               _lib.ptr(rhs[i % R]), _lib.ptr(gates), gates.numel(), _lib.ptr(dx), None, _lib.ptr(dpacked),
                                               B, T, I, Ni, Nc, O,
                                               _lib.ptr(ws), ws.numel(), lay_b, st), "vb_liquid_bwd")
            for i in range(3):
                call(i)
            torch.cuda.synchronize()
            a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for i in range(iters):
                call(i)
            b_.record()
            torch.cuda.synchronize()
            out.append(a.elapsed_time(b_) / iters)
        return out

    fwd_ms, bwd_ms = direct_kernel_ms()      # bwd_ms includes the 2 us partial-sum reduction launched by vb_liquid_bwd
    peaks, peak_kind = measured_peaks()
    hbm_peak = float(peaks["hbm_gbs"])
    bwd_gbs = ALGO_BYTES_BWD * B * T / (bwd_ms * 1e-3) / 1e9 if bwd_ms > 0 else 0.0
    fwd_gbs = ALGO_BYTES_FWD * B * T / (fwd_ms * 1e-3) / 1e9 if fwd_ms > 0 else 0.0
    step_gbs = algo_bytes_step(T) * B * T / (ms_total / K * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")      # dram bytes per launch from the committed ncu capture
    if os.path.isfile(tpath):
        with open(tpath) as f:
            traffic = json.load(f).get("liquid_bwd_tc_kernel_bytes_per_launch")

    # ------------------------------------------------------------------ end to end: pinned host inputs, H2D + D2H inside
    # Every step's INPUT x comes from pinned host memory (double-buffered on a copy stream, overlapping the previous step's
    # compute); the loss weights r_y, r_h are the fixed tensors of BASELINE config 3 ("loss = sum y r_y + sum h_T r_h with fixed
    # random r") and stay on the device; the step's result (loss + every parameter gradient) is read back to pinned host memory.
    # The compute part is the same public-API code, captured once per input slot in a CUDA graph like the device-resident loop.
    def host_seq(F):
        return torch.randn(T, B, F).pin_memory().permute(1, 0, 2) if tm else torch.randn(B, T, F).pin_memory()

    hx = [host_seq(SHAPE["I"]) for _ in range(2)]
    n_grad = sum(p.numel() for p in params)
    hout = torch.empty(n_grad + 1).pin_memory()
    copy_stream = torch.cuda.Stream(device=dev)
    main_stream = torch.cuda.current_stream(dev)
    dbuf = [torch.empty_like(xs[0]) for _ in range(2)]
    r_y, r_h = rys[0], rhs[0]
    ready = [torch.cuda.Event() for _ in range(2)]
    freed = [torch.cuda.Event() for _ in range(2)]

    def upload(i: int) -> None:
        s = i % 2
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(freed[s])
            dbuf[s].copy_(hx[s], non_blocking=True)
            ready[s].record(copy_stream)

    def e2e_compute(s: int) -> torch.Tensor:
        for p in params:
            p.grad = None
        y, _, h_last = net.forward_seq(dbuf[s], None, return_last=True)
        loss = (y * r_y).sum() + (h_last * r_h).sum()
        loss.backward()
        return torch.cat([loss.detach().reshape(1)] + [p.grad.reshape(-1) for p in params])

    e2e_graphs, e2e_out = None, [None, None]
    if graphs is not None:
        try:
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(main_stream)
            with torch.cuda.stream(side):
                for s in range(2):
                    e2e_compute(s)
            main_stream.wait_stream(side)
            barrier()
            e2e_graphs, pool = [], None
            for s in range(2):
                gr = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gr, pool=pool):
                    e2e_out[s] = e2e_compute(s)
                pool = gr.pool()
                e2e_graphs.append(gr)
        except Exception as e:
            if world > 1:
                raise
            print(f"[bench] CUDA graph capture of the end-to-end step failed ({e!r}); running it eagerly", file=sys.stderr)
            e2e_graphs = None
            torch.cuda.synchronize()

    def e2e_step(i: int) -> None:
        s = i % 2
        upload(i + 1)                                   # next step's input travels while this step computes
        main_stream.wait_event(ready[s])
        if e2e_graphs is not None:
            e2e_graphs[s].replay()
            flat = e2e_out[s]
        else:
            flat = e2e_compute(s)
        freed[s].record(main_stream)
        hout.copy_(flat, non_blocking=True)             # D2H of the step's result (loss + all parameter gradients)

    for s in range(2):
        freed[s].record(main_stream)
    upload(0)
    n_warm = 3 + (extra_warm if e2e_graphs is not None else 0)      # see the device-resident loop: the first replays of a graph
    for i in range(n_warm):                                          # holding the all-reduce are slow
        e2e_step(i)
    torch.cuda.synchronize()
    barrier()
    sampler.active = True
    align()
    w0 = time.perf_counter()
    e0.record()
    for i in range(n_warm, n_warm + K):
        e2e_step(i)
    e1.record()
    barrier()
    sampler.windows.append((w0, time.perf_counter()))
    sampler.active = False
    e2e_ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([e2e_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    e2e_value = world * B * T * K / (e2e_ms * 1e-3)
    h2d = 4 * B * T * SHAPE["I"]
    d2h = 4 * (n_grad + 1)
    sampler.stop_flag = True

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        times, threads, used, kind = time_cpu(B, T, 3, 1)
        best = min(times)
        cpu = {"value": B * T / best, "unit": UNIT, "cores": threads, "kind": kind,
               "sample": cpu_sample_text(kind, B, T, "full size, best of 3 steps", used)}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": workload_config(B, T),
            "details": {"global_batch": B * world,
                        "parallelism": f"dp{world}" if world > 1 else "single",
                        "l2": f"{R} rotating input sets ({R * 4 * (B * T * (SHAPE['I'] + SHAPE['O']) + B * SHAPE['Nc']) // 2**20} MiB) + ~170 MB touched per step > 126 MB L2",
                        "e2e_inputs": "x from pinned host memory every step; loss weights r_y, r_h fixed on the device (config 3); loss + gradients read back",
                        "layout": "time-major storage [T,B,F] viewed as [B,T,F]" if tm else "batch-major [B,T,F]",
                        "launch": f"{R} CUDA graphs (one per input set) replayed round-robin" if graphs is not None else "eager",
                        "eager_ms_per_step": eager_ms / K, "extra_untimed_graph_replays": extra_warm,
                        "gradient_exchange": ("one-shot NVLink kernel (vb_allreduce_oneshot)" if vb.dataparallel._state["oneshot"] is not None
                                              else "NCCL all-reduce") if world > 1 else None,
                        "host_cpus_bound_to_gpu_numa_node": len(numa) if numa else None},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms / K},
            "gpu_launches": gpu_launches,
            "batch_major": bm,
            "roofline": {"bound": "hbm", "kernel": "liquid_bwd_tc_kernel", "achieved": bwd_gbs, "peak": hbm_peak,
                         "unit": "GB/s", "frac": bwd_gbs / hbm_peak, "traffic": traffic, "peak_kind": peak_kind,
                         "algorithmic_bytes_per_sample_step": ALGO_BYTES_BWD, "kernel_ms": bwd_ms,
                         "saved_activation_bytes_per_sample_step": SAVED_BYTES,
                         "traffic_note": "traffic exceeds the algorithmic bytes by design: the forward saves the gate factors (160 B per "
                                         "sample-step) and the backward reads them back instead of re-running the five heads",
                         "fwd_kernel_ms": fwd_ms, "fwd_achieved": fwd_gbs,
                         "timing": "20 back-to-back launches between CUDA events on the launching stream, inputs rotating over 4 sets",
                         "eager_event_ms": {"bwd": eager_bwd_ms, "fwd": eager_fwd_ms},
                         "step_achieved": step_gbs, "step_frac": step_gbs / hbm_peak,
                         "note": "frac = the dominant (backward) kernel; step_frac = the whole timed step at SURVEY 8(d)'s 131 B per sample-step. "
                                 "kernel_ms + fwd_kernel_ms can exceed ms_per_step slightly: they are timed as 20 back-to-back launches of one entry "
                                 "point, the step is a graph of five kernels whose tails overlap the next launch",
                         "step_algorithmic_bytes_per_sample_step": algo_bytes_step(T)},
            "cpu_baseline": cpu,
            "clocks": sampler.summary(),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        # captured graphs hold NCCL work: drop them before the communicator, and do not let interpreter teardown order
        # (graphs / communicator / sampler thread) decide whether the process exits
        graphs = None
        import gc
        gc.collect()
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=65536, help="batch rows per GPU")
    ap.add_argument("--seq", type=int, default=32)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-gpu-incumbent", action="store_true", help="reference arm: skip timing the reference modules on the GPU")
    ap.add_argument("--nccl-allreduce", action="store_true", help="exchange the gradient block with NCCL instead of the one-shot NVLink kernel")
    ap.add_argument("--no-graph", action="store_true", help="time the eager Python loop instead of CUDA-graph replays")
    ap.add_argument("--batch-major", action="store_true", help="store the sequence tensors [B,T,F] instead of time-major")
    ap.add_argument("--no-batch-major", action="store_true", help="skip the additional batch-major measurement")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

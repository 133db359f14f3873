#!/usr/bin/env python
"""bench.py — fused ITC + ITM + fusion-head forward+backward throughput (samples/s) on B200.

    python bench.py --gpus N --steps K --warmup W [--workload c2|c3|c4|itc:<B>x<d>] [--impl reference]

A "step" is one pass of the hot path (forward, losses, backward) over one batch of synthetic encoder outputs.
Workloads (BASELINE.json `configs`):
  c2  (default, configs[1]) BERT-base + ViT-B/16 late concat fusion + ITC + ITM, batch 256 per GPU, E=768, P=512, bf16
  c3  ITC only, synthetic d=768 embeddings, 4096 rows per GPU against the global batch (32k at 8 GPUs)
  c4  attention fusion (128 text tokens x 197 patches) + ITC + ITM with similarity-weighted HARD-NEGATIVE mining,
      batch 4096 per GPU (configs[3])
  itc:<B>x<d>  ITC only sweep point (c5)
One JSON line is printed by rank 0 (see the task contract for the keys).
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", type=str, default="c2")
    ap.add_argument("--impl", type=str, default="tic", choices=["tic", "reference"])
    ap.add_argument("--no-graph", action="store_true", help="launch kernels eagerly instead of replaying a CUDA graph")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile", action="store_true", help="short run for ncu: no e2e leg, no per-kernel timing, no CPU leg")
    ap.add_argument("--dist", type=str, default="peer", choices=["peer", "nccl"],
                    help="multi-GPU exchange: peer = own kernels over CUDA-IPC peer memory (default), nccl = collectives (baseline)")
    ap.add_argument("--no-scale-point", action="store_true", help="skip the extra ITC-only B=16384 roofline point")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget for the cpu_baseline leg")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------ workloads
def workload_spec(name, world):
    E = 768
    if name == "c2":
        return dict(workload="c2: concat fusion + ITC + ITM, B=256/GPU, E=768, P=512, C=4", B=256, E=E, P=512, C=4,
                    fusion="concat", use_itc=True, use_itm=True, Lt=128, Lv=197)
    if name == "c4":
        return dict(workload="c4: attention fusion (128x197 tokens) + ITC + ITM with hard-negative mining (similarity-weighted "
                             "multinomial sampling from the ITC tiles + pair gather), B=4096/GPU, E=768, P=512, C=4", B=4096,
                    E=E, P=512, C=4, fusion="attention", use_itc=True, use_itm=True, Lt=128, Lv=197, itm_mode="hard")
    if name == "c3":
        return dict(workload="c3: ITC only, d=768, 4096 rows/GPU vs global batch", B=4096, E=E, P=None, d=768, C=4, fusion=None,
                    use_itc=True, use_itm=False, Lt=0, Lv=0)
    if name.startswith("itc:"):
        b, d = name[4:].split("x")
        return dict(workload="c5: ITC only, B=%s/GPU, d=%s" % (b, d), B=int(b), E=E, P=None, d=int(d), C=4, fusion=None,
                    use_itc=True, use_itm=False, Lt=0, Lv=0)
    raise SystemExit("unknown workload %r" % name)


def make_inputs(spec, seed=40, rank=0):
    """Synthetic encoder outputs: N(0,1) hidden states, tanh-squashed pools, one-hot labels, U[0,1) uniforms.
    Returned on the HOST (fp32 master copies); the caller rounds to bf16 / moves to the device."""
    g = torch.Generator().manual_seed(seed + 1000 * rank)
    B, E = spec["B"], spec["E"]
    inp = {}
    if spec["fusion"] is None:
        d = spec["d"]
        inp["t_pool"] = torch.randn(B, d, generator=g)
        inp["v_pool"] = torch.randn(B, d, generator=g) + 0.25 * inp["t_pool"]
        return inp
    inp["t_pool"] = torch.tanh(torch.randn(B, E, generator=g))
    inp["v_pool"] = torch.tanh(torch.randn(B, E, generator=g))
    # only the CLS row of x_t is consumed by every fusion variant; x_v is read in full by `attention` only
    inp["x_t"] = torch.randn(B, 1, E, generator=g)
    inp["x_v"] = torch.randn(B, spec["Lv"] if spec["fusion"] == "attention" else 1, E, generator=g)
    y = torch.randint(0, spec["C"], (B,), generator=g)
    inp["y_soft"] = torch.eye(spec["C"])[y]
    inp["u_coin"] = torch.rand(B, generator=g)
    inp["u_pick"] = torch.rand(B, generator=g)
    inp["ids"] = torch.randint(5, 30000, (B, spec["Lt"]), generator=g)
    inp["mask"] = torch.ones(B, spec["Lt"], dtype=torch.int64)
    return inp


BF16_KEYS = ("t_pool", "v_pool", "x_t", "x_v")


def synthetic_params(num_labels, E=768, P=512, seed=40):
    """Random-init weights of the head (nn.Linear-style U(-1/sqrt(in), 1/sqrt(in)); mm_late.py:71-89 + the two HF projection
    layers, logit_scale = 2.6592) on numpy's legacy MT19937 stream.  bench.py's own generator: the GPU arm does not touch
    oracle/ (the CPU legs use the oracle's own, which draws the same stream, so both arms see the same weights)."""
    rs = np.random.RandomState(seed)

    def lin(o, i, bias=True):
        bound = 1.0 / math.sqrt(i)
        w = torch.from_numpy(rs.uniform(-bound, bound, size=(o, i))).float()
        b = torch.from_numpy(rs.uniform(-bound, bound, size=(o,))).float() if bias else None
        return w, b

    p = {}
    if P is not None:
        p["dual_encoder.visual_projection.weight"], _ = lin(P, E, False)
        p["dual_encoder.text_projection.weight"], _ = lin(P, E, False)
    p["dual_encoder.logit_scale"] = torch.tensor(2.6592)
    for name, (o, i) in (("fc_Q", (E, E)), ("fc_K", (E, E)), ("fc_V", (E, E)), ("aspectattention", (1, E)),
                         ("linear_fusion", (E, 2 * E)), ("linear_cls", (num_labels, E)), ("linear_tim", (2, E)),
                         ("linear_iadds", (2, E)), ("linear_gmu_t", (2 * E, E)), ("linear_gmu_v", (2 * E, E))):
        p[name + ".weight"], p[name + ".bias"] = lin(o, i)
    return p


def algorithmic_work(spec, n_global):
    """FLOPs / bytes per step per GPU as defined in SURVEY.md §8(d) (recompute is NOT counted)."""
    B, E = spec["B"], spec["E"]
    d = spec["P"] if spec["P"] is not None else spec.get("d", E)
    w = {"itc_flops": 6.0 * B * n_global * d}
    if spec["P"] is not None:
        w["proj_flops"] = 10.0 * B * E * spec["P"]
    if spec["fusion"] == "concat":
        passes = 2 if spec["use_itm"] else 1
        w["fusion_flops"] = passes * 3 * 2.0 * (2 * E) * E * B
    if spec["fusion"] == "attention":
        w["attn_bytes"] = 2.0 * spec["Lv"] * E * 2 * B
    return w


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def wait_first(self, timeout_s=4.0):
        """nvidia-smi takes ~1 s to come up (longer with 8 ranks starting at once): do not enter a sub-second timed
        region before the sampler delivers."""
        t0 = time.time()
        while self.proc is not None and not self.rows and time.time() - t0 < timeout_s:
            time.sleep(0.05)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        busy = sorted(x for x in sm if x > 0)
        # median of the upper half = clocks under load (the sampler also sees the idle gaps between steps)
        med = busy[len(busy) * 3 // 4] if busy else None
        return {"sm_mhz": med, "sm_max_mhz": (max(mx) if mx else None), "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ reference / CPU arm
def cpu_reference_run(spec, steps, warmup, budget_s, threads=None, world=1, min_timed_s=2.0):
    """The reference's CPU implementation of the path (oracle restatement of models/mm_late.py + utils.clip_loss, pinned
    against the unmodified reference), fp32, all host threads, on the GLOBAL batch of the GPU arm (B per GPU x world: ITC
    is O(B^2), so a per-GPU batch would be a different problem).  At least max(warmup, 10) untimed steps, then at least
    `steps` timed steps AND `min_timed_s` seconds, all within `budget_s`.  Returns (samples/s, ms/step, sample, threads)."""
    from oracle import restatement as R
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    B_full = spec["B"] * world
    B = B_full
    note = "full global batch B=%d" % B
    if spec["fusion"] is None and B > 8192:
        B = 8192
        note = "B=8192 of %d rows (ITC is O(B^2): CPU throughput at the full size is LOWER than this sample's)" % B_full
    if spec["fusion"] == "attention" and B > 256:
        B = 256
        note = "B=256 of %d samples (attention fusion is per-sample; the ITC part is O(B^2): throughput at the full size is lower)" % B_full
    if spec["fusion"] == "concat" and B > 2048:
        B = 2048
        note = "B=2048 of %d samples" % B_full
    s2 = dict(spec, B=B)
    inp = make_inputs(s2)
    inp = {k: (v.to(torch.bfloat16).float() if k in BF16_KEYS else v) for k, v in inp.items()}
    hard = spec.get("itm_mode") == "hard"
    if spec["fusion"] is None:
        ls = torch.tensor(2.6592, requires_grad=True)

        def step():
            T = inp["t_pool"].clone().requires_grad_(True)
            V = inp["v_pool"].clone().requires_grad_(True)
            loss = R.clip_loss(R.itc_logits(T, V, ls))
            loss.backward()
            return float(loss.detach())
    else:
        p = {k: v.clone().requires_grad_(True) for k, v in R.init_params(spec["C"], seed=40).items()}
        u_coin, u_pick = inp["u_coin"].numpy(), inp["u_pick"].numpy()
        lbl, src = R.itm_sample_uniform(u_coin, u_pick)
        base = dict(inp)
        if spec["fusion"] == "attention":
            base["x_t"] = inp["x_t"].expand(B, 1, spec["E"])  # literal reference attention needs all Lt rows; use collapse

        def step():
            cur = dict(base)
            cur["x_t"] = base["x_t"].clone().requires_grad_(True)
            cur["t_pool"] = base["t_pool"].clone().requires_grad_(True)
            if hard:   # similarity-weighted negatives: sample from this step's logits (oracle spec), then gather ids / mask
                with torch.no_grad():
                    S = R.itc_logits(R.project(cur["t_pool"], p["dual_encoder.text_projection.weight"]),
                                     R.project(base["v_pool"], p["dual_encoder.visual_projection.weight"]),
                                     p["dual_encoder.logit_scale"])
                l_, s_ = R.itm_sample_hard(S.numpy(), u_coin, u_pick, ref=float(np.float32(math.exp(2.6592))))
                idx = torch.from_numpy(s_)
                _tim = (inp["ids"][idx].clone(), inp["mask"][idx].clone())
                cur["lbl_tim"], cur["src_idx"] = torch.from_numpy(l_), idx
            else:      # mm_late.py:389-414 host loop + row copies are part of the path
                R.prepare_itm_inputs_stream(inp["ids"], inp["mask"], np.random.RandomState(0))
                cur["lbl_tim"], cur["src_idx"] = torch.from_numpy(lbl), torch.from_numpy(src)
            for v in p.values():
                v.grad = None
            out = R.head_step(cur, p, fusion_name=spec["fusion"], use_itc=True, use_itm=spec["use_itm"])
            out["loss"].backward()
            return float(out["loss"].detach())
    t_begin = time.perf_counter()
    n_warm = 0
    while n_warm < max(warmup, 10) and (n_warm < 2 or time.perf_counter() - t_begin < 0.25 * budget_s):
        step()
        n_warm += 1
    times = []
    t_timed = time.perf_counter()
    while True:
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
        el = time.perf_counter() - t_timed
        if (len(times) >= steps and el >= min_timed_s) or time.perf_counter() - t_begin > budget_s:
            break
    ms = 1e3 * float(np.median(times))
    return B / (ms / 1e3), ms, "%s, %d warm-up + %d timed steps (%.1f s), median, fp32 torch CPU" % (
        note, n_warm, len(times), sum(times)), threads


# ------------------------------------------------------------------------------------------------ stdout hygiene
class _QuietStdout:
    """Rank 0 must print ONE JSON line.  Libraries write to file descriptor 1 behind Python's back (NCCL prints its version
    banner there when NCCL_DEBUG is set by the environment): while the benchmark runs, fd 1 points at stderr; `emit` puts it
    back and prints the line."""

    def __init__(self):
        self.saved = None
        try:
            sys.stdout.flush()
            self.saved = os.dup(1)
            os.dup2(2, 1)
        except OSError:
            self.saved = None

    def emit(self, line):
        sys.stdout.flush()
        if self.saved is not None:
            try:
                os.dup2(self.saved, 1)
                os.close(self.saved)
            except OSError:
                pass
            self.saved = None
        print(line, flush=True)


# ------------------------------------------------------------------------------------------------ main
def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    spec = workload_spec(args.workload, world)
    metric = "fused ITC+ITM+fusion fwd+bwd samples/sec"
    base = {"metric": metric, "unit": "samples/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "data": "synthetic",
            "config": {"workload": spec["workload"], "per_gpu_batch": spec["B"], "global_batch": spec["B"] * world,
                       "parallelism": "dp%d" % world if world == 1 else "dp%d (%s exchange)" % (world, args.dist)}}

    if args.impl == "reference":
        if rank != 0:
            return
        val, ms, sample, threads = cpu_reference_run(spec, max(args.steps, 3), args.warmup, 150.0, world=world)
        line = dict(base, impl="reference", value=val, ms_per_step=ms, dtype="f32", gpu_launches=0,
                    cpu_baseline={"value": val, "unit": "samples/s", "cores": threads, "kind": "port", "sample": sample},
                    e2e={"value": val, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0})
        print(json.dumps(line))
        return

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback)"
    out = _QuietStdout()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist_mod.init_process_group("nccl", device_id=dev)
        dist = dist_mod

    import tic_b200.plan as P
    from tic_b200 import capi

    B = spec["B"]
    host = make_inputs(spec, rank=rank)
    dev_in = {k: (v.to(torch.bfloat16) if k in BF16_KEYS else v).to(dev) for k, v in host.items()}
    n_global = B * world
    if world > 1:
        if args.dist == "peer":
            from tic_b200.peer import PeerHeadPlan as DistHeadPlan
        else:
            from tic_b200.dist import DistHeadPlan
        plan = DistHeadPlan(B, world=world, rank=rank, E=spec["E"], P=spec["P"] if spec["P"] is not None else None,
                            d=spec.get("d"), C=spec["C"], fusion=spec["fusion"], use_itc=spec["use_itc"],
                            use_itm=spec["use_itm"], Lv=spec["Lv"], device=dev)
    else:
        plan = P.HeadPlan(B, E=spec["E"], P=spec["P"], C=spec["C"], fusion=spec["fusion"], use_itc=spec["use_itc"],
                          use_itm=spec["use_itm"], Lv=max(spec["Lv"], 1), device=dev, itm_mode=spec.get("itm_mode", "uniform"))
        if spec["P"] is None:
            plan.itc = P.ItcPlan(B, B, spec["d"], dev)
            plan.itc.scale_dev = plan.scale_t
            plan.Pe = spec["d"]
            plan.out["d_t_emb"] = torch.empty(B, spec["d"], device=dev)
            plan.out["d_v_emb"] = torch.empty(B, spec["d"], device=dev)
    # fp32 MASTER weights resident on the device; the step itself (first node of the captured graph) refreshes the bf16
    # working copies and exp(logit_scale) from them, so the timed step is a training step's forward+backward, not a step on
    # pre-cast constants.  (The NCCL comparison plan takes a snapshot: set_weights; TIC_BENCH_SNAPSHOT=1 forces it: A/B.)
    master = {k: v.to(dev) for k, v in synthetic_params(spec["C"], seed=40).items()}
    if (world == 1 or args.dist == "peer") and os.environ.get("TIC_BENCH_SNAPSHOT", "0") != "1":
        plan.bind_params(master, live=True)
    else:
        plan.set_weights(master)

    # ---- count launches of one step (kernels per C-ABI call are fixed)
    KPC = {"tic_itc_lse_loss": 2, "tic_heads_fwd_bwd": 1, "tic_ce_bidir_fwd": 2, "tic_gemm_rowss_parts": 0,
           "tic_itc_row_parts": 0, "tic_itc_col_parts": 0}
    KPC.update({"tic_peer_alloc": 0, "tic_peer_export": 0, "tic_peer_open": 0, "tic_peer_close": 0, "tic_gemm_plan": 0})
    counter = {"n": 0}

    def hook(name):
        counter["n"] += KPC.get(name, 1)

    capi.call_hook = hook
    plan.step(dev_in)
    # + the accumulator memsets (small block on the chain, large block on a side branch)
    launches_per_step = counter["n"] + 1 + (1 if getattr(plan, "_z_pending", False) else 0)
    capi.call_hook = None
    torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        plan.step(dev_in)
    torch.cuda.synchronize()

    use_graph = (not args.no_graph) and (world == 1 or args.dist == "peer")
    graph = None
    if use_graph:
        graph = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            plan.step(dev_in)
        torch.cuda.current_stream().wait_stream(s)
        with torch.cuda.graph(graph):
            plan.step(dev_in)
        for _ in range(3):
            graph.replay()
        torch.cuda.synchronize()

    def run_step():
        if graph is not None:
            graph.replay()
        else:
            plan.step(dev_in)

    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)  # > 126 MB L2

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def align():
        """Multi-GPU: line the ranks up ON THE DEVICE right before a start event (peer barrier kernel; NCCL barrier in the
        nccl mode), so host launch skew between the ranks is not billed to the step."""
        if dist is None:
            return
        if hasattr(plan, "pg"):
            plan.pg.exchange("align")
        else:
            dist.barrier()

    # ---- timed region: K steps, each bracketed by its own event pair, L2 flushed between steps (outside the pair)
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
        clocks.wait_first()
    barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for k in range(args.steps):
        flush.fill_(float(k))
        align()
        ev[k][0].record()
        run_step()
        ev[k][1].record()
    barrier()
    total_ms = sum(a.elapsed_time(b) for a, b in ev)
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = n_global / (ms_per_step / 1e3)

    if args.profile:
        if rank == 0:
            clocks.stop()
            out.emit(json.dumps(dict(base, value=value, ms_per_step=ms_per_step, profile_run=True)))
        return
    # ---- end-to-end: HOST buffers in, loss out, through the public host-facing API.
    # (1) synchronous: every call = pinned arena -> H2D -> step -> D2H losses -> host sync; L2 flushed outside the brackets.
    runner = P.HostStep(plan, host, bf16_keys=BF16_KEYS, use_graph=use_graph)
    for _ in range(3):
        runner()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2e_ms = 0.0
    for k in range(args.steps):
        flush.fill_(float(k))
        torch.cuda.synchronize()
        align()
        e0.record()
        loss_host = runner()       # pinned host arena -> H2D -> step -> D2H losses (synchronous)
        e1.record()
        e1.synchronize()
        e2e_ms += e0.elapsed_time(e1)
    t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_sync_ms = float(t.item()) / args.steps
    # (2) pipelined (the headline): two slots, the H2D of step k+1 overlaps the kernels of step k; every step still does
    # its own H2D + D2H, and the L2 flush runs INSIDE the timed region between consecutive steps.
    e2e_ms, e2e_mode = e2e_sync_ms, "synchronous call per step"
    if use_graph:
        flush2 = torch.empty(160 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)  # > 126 MB L2
        pipe = P.HostPipeline(plan, host, bf16_keys=BF16_KEYS, between=lambda: flush2.fill_(1.0))
        for _ in range(3):
            pipe.submit()
            pipe.result()
        barrier()
        align()
        e0.record()
        for k in range(args.steps):
            pipe.submit()
            if k >= 1:
                loss_host = pipe.result()
        loss_host = pipe.result()
        e1.record()
        e1.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        pipe_ms = float(t.item()) / args.steps
        if pipe_ms < e2e_ms:
            e2e_ms, e2e_mode = pipe_ms, "pipelined, 2 slots: H2D of step k+1 overlaps step k; 160 MiB L2 flush inside the timed region"
    clk = clocks.stop() if rank == 0 else None

    # ---- the drop-in API itself (c2, one GPU): MM_Model.head() -> the reference's loss code -> loss.backward()
    dropin = None
    if world == 1 and args.workload == "c2":
        dropin = dropin_leg(spec, host, dev, master, flush, args.steps)
    # ---- multi-GPU: the global ITC loss of the peer-memory step against an NCCL all-gather + plain torch fp32 evaluation
    check = global_loss_check(plan, dist, dev, n_global) if (world > 1 and spec["use_itc"]) else None

    # ---- per-kernel timing of the dominant kernels (live, CUDA events on the launching stream, L2 flushed)
    roof, kernels = kernel_rooflines(P, plan, spec, dev_in, n_global, flush)
    if world > 1 and hasattr(plan, "pg"):
        kernels += exchange_points(plan, flush)
    hbm_points = hbm_kernel_points(P, dev, flush) if world == 1 else None
    # the default workload (B=256) is launch-latency-bound: also time the same ITC kernels where a roofline applies
    at_scale = None
    if world == 1 and spec["B"] < 4096 and spec["use_itc"] and not args.no_scale_point:
        at_scale = scale_point(P, dev, flush)

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    line = dict(base, value=value, ms_per_step=ms_per_step, dtype="bf16", gpu_launches=launches_per_step * args.steps,
                clocks=clk, e2e={"value": n_global / (e2e_ms / 1e3), "unit": "samples/s", "ms_per_step": e2e_ms,
                                 "h2d_bytes_per_step": runner.h2d_bytes, "d2h_bytes_per_step": runner.d2h_bytes,
                                 "loss": loss_host, "mode": e2e_mode, "sync_ms_per_step": e2e_sync_ms,
                                 "returns": "the 4 loss scalars only (mix, cls, itc, itm): gradients and outputs stay on the "
                                            "device for a device-side optimiser"})
    if dropin is not None:
        line["e2e_dropin"] = dropin
    if check is not None:
        line["global_loss_check"] = check
    if hbm_points is not None:
        line["kernels_hbm_4096"] = [finalize_roofline(k, peaks) for k in hbm_points]
    line["config"].update({"l2": "flushed between timed steps (256 MiB write)", "cuda_graph": bool(use_graph),
                           "launches_per_step": launches_per_step, "loss": float(plan.out["loss"][0])})
    line["roofline"] = finalize_roofline(roof, peaks)
    line["kernels"] = [finalize_roofline(k, peaks) for k in kernels]
    if at_scale is not None:
        line["roofline_at_scale"] = {"workload": at_scale["workload"], "why": at_scale["why"],
                                     "step": finalize_roofline(at_scale["step"], peaks),
                                     "kernels": [finalize_roofline(k, peaks) for k in at_scale["kernels"]]}
    if not args.no_cpu_baseline:
        val, ms, sample, threads = cpu_reference_run(spec, 20, 10, args.cpu_seconds + 8.0, world=world,
                                                     min_timed_s=args.cpu_seconds)
        line["cpu_baseline"] = {"value": val, "unit": "samples/s", "cores": threads, "kind": "port", "sample": sample,
                                "ms_per_step": ms}
    out.emit(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def finalize_roofline(r, peaks):
    if r is None:
        return None
    if r["bound"] == "nvlink":
        out = dict(r)
        out.update(peak=900.0, frac=r["achieved"] / 900.0, peak_source="NVLink 5 via NVSwitch: 900 GB/s per direction per GPU (nominal)",
                   traffic=None, traffic_source=None)
        return out
    if r["bound"] == "tensor":
        peak = peaks.get("bf16_tflops", 1590.0)
        src = "MEASURED_PEAKS.json bf16_tflops (burst; kernel timed alone)" if "bf16_tflops" in peaks else "fallback 1590"
    else:
        peak = peaks.get("hbm_gbs", 6650.0)
        src = "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650"
    out = dict(r)
    key = out.pop("traffic_key", None)
    traffic, tsrc = None, None
    ent = _traffic_db().get(key) if key else None
    if ent:   # dram__bytes_read.sum + dram__bytes_write.sum of this kernel from one committed `ncu --set full` capture
        traffic, tsrc = ent.get("traffic_bytes"), "profiles/ncu_traffic.json[%s] (%s launch(es), ncu, bytes)" % (key, ent.get("launches", 1))
    out.update(peak=peak, frac=r["achieved"] / peak, peak_source=src, traffic=traffic, traffic_source=tsrc)
    return out


_TRAFFIC = None


def _traffic_db():
    global _TRAFFIC
    if _TRAFFIC is None:
        try:
            _TRAFFIC = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        except Exception:
            _TRAFFIC = {}
    return _TRAFFIC


def time_kernel(fn, flush, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for k in range(iters):
        flush.fill_(float(k))
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / iters


def time_kernel_graph(fn, flush, reps=8, rounds=5):
    """Duration of a SHORT kernel with a cold L2 and without the host's launch gap: a CUDA graph of `reps` x (L2 flush, fn)
    minus a graph of `reps` x (L2 flush), best of `rounds`.  (Eager event pairs around a < 10 us kernel mostly time the
    ctypes call that launches it.)"""
    def build(with_fn):
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            if with_fn:
                fn()
        torch.cuda.current_stream().wait_stream(s)
        with torch.cuda.graph(g):
            for r in range(reps):
                flush.fill_(float(r))
                if with_fn:
                    fn()
        return g
    g1, g0 = build(True), build(False)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def run(g):
        best = 1e30
        for _ in range(rounds):
            e0.record(); g.replay(); e1.record(); e1.synchronize()
            best = min(best, e0.elapsed_time(e1))
        return best
    run(g1); run(g0)
    return max(run(g1) - run(g0), 1e-6) / reps


def kernel_rooflines(P, plan, spec, dev_in, n_global, flush):
    """Times the hot kernels of this workload one by one; returns (dominant, all)."""
    B, E = spec["B"], spec["E"]
    ks = []
    if spec["use_itc"] and hasattr(plan, "itc") and plan.itc is not None and n_global == B:
        it = plan.itc
        if spec["P"] is not None:
            Yt, Yv = plan.Y[:B], plan.Y[B:]
        else:
            Yt, Yv = dev_in["t_pool"], dev_in["v_pool"]
        d = it.P
        ldt, ldv = Yt.stride(0), Yv.stride(0)
        ms = time_kernel(lambda: it.fwd_tiles(Yt, ldt, Yv, ldv, plan.scale), flush)
        shp = "%dx%d" % (B, d)
        ks.append(dict(kernel="tic_itc_fwd (tcgen05 similarity tiles + fused bidirectional softmax-CE)", bound="tensor",
                       ms=ms, achieved=2.0 * B * B * d / ms * 1e-9, unit="TFLOP/s", algorithmic="2*B^2*d FLOP",
                       traffic_key="itc_fwd_" + shp))
        ms = time_kernel(lambda: it.bwd_operands(Yt, ldt, Yv, ldv, plan.scale, 1.0 / (2 * B)), flush)
        ks.append(dict(kernel="tic_itc_bwd_g (tile recompute -> bf16 gradient operands)", bound="tensor", ms=ms,
                       achieved=2.0 * B * B * d / ms * 1e-9, unit="TFLOP/s",
                       algorithmic="recompute: 2*B^2*d executed FLOP, 0 algorithmic (reported as executed)",
                       traffic_key="itc_bwd_" + shp))
        ms = time_kernel(lambda: it.grad_gemms(Yt, ldt, Yv, ldv), flush)
        ks.append(dict(kernel="tic_gemm_bf16 x2 (dT = GA*V, dV = GBT*T)", bound="tensor", ms=ms,
                       achieved=4.0 * B * B * d / ms * 1e-9, unit="TFLOP/s", algorithmic="4*B^2*d FLOP",
                       traffic_key="gemm_dtdv_" + shp))
        if getattr(plan, "fuse_itc_small", False) and it.can_fuse_small and spec["P"] is not None:
            # what the step itself launches at this size instead of the forward / gradient-operand pair above (timed with the
            # step's own operands: split-precision embeddings)
            Ft, Fv, Ftl, Fvl = plan._itc_operands(dev_in)
            ms = time_kernel_graph(lambda: it.fwd_bwd_fused(Ft, Ft.stride(0), Fv, Fv.stride(0), plan.scale, 1.0 / (2 * B), T_lo=Ftl,
                                                            V_lo=Fvl), flush)
            ks.append(dict(kernel="tic_itc_fwd_bwd_small (ONE cluster launch: similarity tiles, softmax statistics, gradient operands; "
                                  "the form the step uses at this size)", bound="tensor", ms=ms,
                           achieved=2.0 * B * B * d / ms * 1e-9, unit="TFLOP/s",
                           algorithmic="2*B^2*d FLOP (the k-loop runs once; split-precision operands execute 3x that)"))
    if spec["use_itc"] and getattr(plan, "itc_mode", None) == "rowblock":
        # multi-GPU row block (this rank's m = B rows against the n_global gathered columns): the rank-local tensor kernels
        blk, T, V_all = plan.rb, plan.Y[:B], plan.V_all
        d, ldt, ldv = blk.P, plan.Y.stride(0), plan.V_all.stride(0)
        shp = "%dx%dx%d" % (B, n_global, d)
        ms = time_kernel(lambda: blk.fwd_tiles(T, ldt, V_all, ldv, plan.scale), flush)
        ks.append(dict(kernel="tic_itc_fwd row block %s (per rank)" % shp, bound="tensor", ms=ms,
                       achieved=2.0 * B * n_global * d / ms * 1e-9, unit="TFLOP/s", algorithmic="2*m*N*d FLOP per rank"))
        ms = time_kernel(lambda: blk.bwd_operands(T, ldt, V_all, ldv, plan.scale, 1.0 / (2 * n_global)), flush)
        ks.append(dict(kernel="tic_itc_bwd_g row block %s (per rank)" % shp, bound="tensor", ms=ms,
                       achieved=2.0 * B * n_global * d / ms * 1e-9, unit="TFLOP/s",
                       algorithmic="recompute: 2*m*N*d executed FLOP per rank, 0 algorithmic (reported as executed)"))
        ms = time_kernel(lambda: (blk.grad_gemm_t(V_all, ldv), blk.grad_gemm_v(T, ldt)), flush)
        ks.append(dict(kernel="tic_gemm_bf16 x2 row block %s (dT_r = GA*V_all, dV contributions = GA^T*That_r)" % shp, bound="tensor",
                       ms=ms, achieved=4.0 * B * n_global * d / ms * 1e-9, unit="TFLOP/s", algorithmic="4*m*N*d FLOP per rank"))
    if spec["fusion"] == "attention":
        x_v = dev_in["x_v"]
        R_, Lv, Ea = plan.R, spec["Lv"], E + 8
        npass = 2 if spec["use_itm"] else 1
        st = lambda: torch.cuda.current_stream().cuda_stream  # noqa: E731
        ms = time_kernel(lambda: P.call("tic_attn_pool_fwd", x_v.data_ptr(), x_v.stride(0), x_v.stride(1), plan.kq.data_ptr(),
                                        Ea, B, npass, Lv, E, float(E) ** -0.5, plan.xbar_b.data_ptr(), plan.xbar_lo.data_ptr(),
                                        E, plan.xbar_f.data_ptr(), E, plan.attn.data_ptr(), Lv, st()), flush)
        ks.append(dict(kernel="tic_attn_pool_fwd (one streaming read of x_v, main+ITM pass)", bound="hbm", ms=ms,
                       achieved=B * Lv * E * 2.0 / ms * 1e-6, unit="GB/s", algorithmic="B*Lv*E*2 bytes",
                       traffic_key="attn_fwd_%d" % B))
        ms = time_kernel(lambda: P.call("tic_attn_pool_bwd", x_v.data_ptr(), x_v.stride(0), x_v.stride(1), plan.attn.data_ptr(),
                                        Lv, plan.dxbar.data_ptr(), E, plan.xbar_f.data_ptr(), E, B, npass, Lv, E,
                                        float(E) ** -0.5, plan.dkq.data_ptr(), plan.dkq_lo.data_ptr(), Ea, st()), flush)
        ks.append(dict(kernel="tic_attn_pool_bwd (second streaming read of x_v)", bound="hbm", ms=ms,
                       achieved=B * Lv * E * 2.0 / ms * 1e-6, unit="GB/s", algorithmic="B*Lv*E*2 bytes",
                       traffic_key="attn_bwd_%d" % B))
    if spec["fusion"] in ("concat",) and getattr(plan, "pairwise", False):
        E2, x_t = 2 * E, dev_in["x_t"]
        ms = time_kernel(lambda: P.gemm(x_t, x_t.stride(0), 0, plan.w["W_f"], E2, 0, plan.Pt, E, 0, B, E, E), flush)
        ks.append(dict(kernel="tic_gemm_bf16 (linear_fusion, text half Pt = x_t W_f[:, :E]^T; the image half runs beside it)",
                       bound="tensor", ms=ms, achieved=2.0 * B * E * E / ms * 1e-9, unit="TFLOP/s", algorithmic="2*B*E*E FLOP",
                       traffic_key="gemm_fusion_half_%d" % B))
    elif spec["fusion"] in ("concat",):
        R_, E2 = plan.R, 2 * E
        ms = time_kernel(lambda: P.gemm(plan.Xcat, E2, 0, plan.w["W_f"], E2, 0, plan.H, E, 0, R_, E, E2, bias=plan.w["b_f"],
                                        relu=True), flush)
        ks.append(dict(kernel="tic_gemm_bf16 (linear_fusion forward, bias+ReLU epilogue)", bound="tensor", ms=ms,
                       achieved=2.0 * R_ * E * E2 / ms * 1e-9, unit="TFLOP/s", algorithmic="2*R*E*2E FLOP",
                       traffic_key="gemm_fusion_%dx%d" % (B, spec["P"] or 0)))
    if not ks:
        return None, []
    dom = max(ks, key=lambda k: k["ms"])
    return dom, ks


def dropin_leg(spec, host, dev, master, flush, steps):
    """The reference-facing API timed the way the reference's loop drives it (models/mm_late.py:459-490): HOST encoder outputs
    -> device -> MM_Model.head() (autograd Function over the HeadPlan) -> nn.CrossEntropyLoss / clip_loss / ITM CE ->
    loss.backward() -> loss.item().  No CUDA graph, torch's autograd and allocator in the loop: this is what a user who only
    swaps the import gets; `value` / `e2e` are what the fused plan API gets."""
    import torch.nn as nn
    from tic_b200.mm_late import MM_Model
    from tic_b200.utils import clip_loss
    B, E, C = spec["B"], spec["E"], spec["C"]

    class _Stub(nn.Module):   # the three non-tower parameters of the HF dual encoder (the towers are outside this path)
        def __init__(self):
            super().__init__()
            self.visual_projection = nn.Linear(E, spec["P"], bias=False)
            self.text_projection = nn.Linear(E, spec["P"], bias=False)
            self.logit_scale = nn.Parameter(torch.tensor(2.6592))

    model = MM_Model(C, "bert", "vit", 0.0, fusion_name=spec["fusion"], dual_encoder=_Stub())
    sd = model.state_dict()
    for k, v in master.items():
        if k in sd:
            sd[k].copy_(v)
    model = model.to(dev).train()
    loss_fn, tim_loss_fn = nn.CrossEntropyLoss(), nn.CrossEntropyLoss()
    pinned = {k: v.to(torch.bfloat16 if k in BF16_KEYS else v.dtype).pin_memory() for k, v in host.items()
              if k in ("x_t", "x_v", "t_pool", "v_pool", "y_soft", "u_coin", "u_pick")}
    h2d = sum(v.numel() * v.element_size() for v in pinned.values())

    def step():
        d = {k: v.to(dev, non_blocking=True) for k, v in pinned.items()}
        x_t = d["x_t"].float().requires_grad_(True)
        t_pool = d["t_pool"].float().requires_grad_(True)
        lbl = (d["u_coin"] >= 0.5).long()                         # uniform rule on the device (no host loop)
        k = torch.clamp((d["u_pick"] * (B - 1)).floor().long(), max=B - 2)
        i = torch.arange(B, device=dev)
        src = torch.where(lbl == 1, i, torch.where(k < i, k, k + 1)).to(torch.int32)
        model.zero_grad(set_to_none=True)
        out_cls, logits, out_tim, _ = model.head(x_t, d["x_v"].float(), t_pool, d["v_pool"].float(), tim_src=src)
        loss = 0.8 * loss_fn(out_cls, d["y_soft"]) + 0.1 * clip_loss(logits) + 0.1 * tim_loss_fn(out_tim, lbl)
        loss.backward()
        return float(loss.item())

    for _ in range(5):
        loss = step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tot = 0.0
    n = max(steps, 10)
    for kk in range(n):
        flush.fill_(float(kk))
        torch.cuda.synchronize()
        e0.record()
        loss = step()
        e1.record()
        e1.synchronize()
        tot += e0.elapsed_time(e1)
    ms = tot / n
    return {"value": B / (ms * 1e-3), "unit": "samples/s", "ms_per_step": ms, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
            "loss": loss, "api": "tic_b200.mm_late.MM_Model.head() + nn.CrossEntropyLoss / tic_b200.utils.clip_loss + "
                                 "loss.backward() (eager autograd, no CUDA graph; mm_late.py:459-490)"}


def global_loss_check(plan, dist, dev, n_global):
    """SCALE correctness: the ITC loss of the multi-GPU step (own peer-memory kernels) against NCCL all_gather of the ranks'
    projected embeddings + a plain torch fp32 (TF32 off) clip_loss on the global matrix, row-chunked.  Raises on > 1e-3."""
    b = plan.B
    if hasattr(plan, "_rows") and getattr(plan, "itc_mode", None) == "symmetric":
        yt, yv, ytl, yvl = plan._rows()
    else:
        yt, yv = plan.Y[:b], plan.Y[b:]
        ytl, yvl = (plan.Y_lo[:b], plan.Y_lo[b:]) if getattr(plan, "Y_lo", None) is not None else (None, None)
    lo = getattr(plan, "has_lo", False)
    Yt = yt.float() + (ytl.float() if (lo and ytl is not None) else 0)
    Yv = yv.float() + (yvl.float() if (lo and yvl is not None) else 0)
    world = dist.get_world_size()
    Ta = [torch.empty_like(Yt) for _ in range(world)]
    Va = [torch.empty_like(Yv) for _ in range(world)]
    dist.all_gather(Ta, Yt.contiguous())
    dist.all_gather(Va, Yv.contiguous())
    T, V = torch.cat(Ta), torch.cat(Va)
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        Tn, Vn = torch.nn.functional.normalize(T, dim=1), torch.nn.functional.normalize(V, dim=1)
        s = float(plan.scale)
        N = T.shape[0]
        diag = s * (Tn * Vn).sum(1)
        lse_r = torch.empty(N, device=dev)
        col_m = torch.full((N,), -float("inf"), device=dev)
        col_s = torch.zeros(N, device=dev)
        for i in range(0, N, 4096):
            S = s * Tn[i:i + 4096] @ Vn.t()
            lse_r[i:i + 4096] = torch.logsumexp(S, 1)
            m = torch.maximum(col_m, S.max(0).values)
            col_s = col_s * torch.exp(col_m - m) + torch.exp(S - m).sum(0)
            col_m = m
        ref = float(0.5 * ((lse_r - diag).mean() + (col_m + torch.log(col_s) - diag).mean()))
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old
    got = float(plan.global_loss()[2]) if hasattr(plan, "global_loss") else float("nan")
    rel = abs(got - ref) / max(abs(ref), 1e-12)
    out = {"itc_loss_global": got, "itc_loss_nccl_allgather_torch_fp32": ref, "rel_err": rel, "tolerance": 1e-3,
           "n_global": n_global}
    if not rel < 1e-3:
        raise SystemExit("multi-GPU ITC loss differs from the all-gathered fp32 evaluation: %r" % out)
    return out


def exchange_points(plan, flush):
    """Peer-exchange kernels timed alone (all ranks call in lock step): barrier + pull over NVLink; achieved = bytes this
    rank pulls from REMOTE peers / time (the local segment is an HBM copy)."""
    pg = plan.pg
    ks = []
    for name, ph in pg._phases.items():
        n, so, nb, dst, ds, _keep = ph
        remote = sum(int(nb[i]) for i in range(n)) * (pg.world - 1)
        ms = time_kernel(lambda: pg.exchange(name), flush, iters=10)
        ks.append(dict(kernel="tic_peer_exchange[%s] (system-scope flag barrier + pull of %d range(s) from %d peers)" % (name, n, pg.world - 1),
                       bound="nvlink", ms=ms, achieved=remote / ms * 1e-6 if remote else 0.0, unit="GB/s",
                       algorithmic="%d bytes pulled from remote peers per rank" % remote))
    return ks


def hbm_kernel_points(P, dev, flush, B=4096, Lt=128):
    """The HBM-bound kernels north_star names, at the c4 batch size (4096), timed alone (CUDA events, L2 flushed):
    materialised clip_loss forward/backward (models/utils.py:225-231; BASELINE.md target 12*B^2 bytes -> 31 us) and the ITM
    sampler + pair gather (mm_late.py:389-414: ids and mask rows, int64 x 128 tokens).  Timed inside CUDA graphs
    (time_kernel_graph): these kernels are shorter than the host's launch gap."""
    st = lambda: torch.cuda.current_stream().cuda_stream  # noqa: E731
    g = torch.Generator().manual_seed(7)
    S = (torch.randn(B, B, generator=g) * 3).to(dev)
    dS = torch.empty_like(S)
    lse_r, lse_c = torch.empty(B, device=dev), torch.empty(B, device=dev)
    loss, gl = torch.zeros(1, device=dev), torch.ones(1, device=dev)
    ws = torch.empty(int(P.capi.load().tic_ce_bidir_workspace_bytes(B)), dtype=torch.uint8, device=dev)
    ks = []
    ms_f = time_kernel_graph(lambda: P.call("tic_ce_bidir_fwd", S.data_ptr(), B, B, lse_r.data_ptr(), lse_c.data_ptr(), loss.data_ptr(),
                                      ws.data_ptr(), st()), flush)
    ks.append(dict(kernel="tic_ce_bidir_fwd (materialised clip_loss: one read of S)", bound="hbm", ms=ms_f,
                   achieved=4.0 * B * B / ms_f * 1e-6, unit="GB/s", algorithmic="4*B^2 bytes (S read once)", traffic_key="ce_fwd_%d" % B))
    ms_b = time_kernel_graph(lambda: P.call("tic_ce_bidir_bwd", S.data_ptr(), B, B, lse_r.data_ptr(), lse_c.data_ptr(), gl.data_ptr(),
                                      dS.data_ptr(), B, st()), flush)
    ks.append(dict(kernel="tic_ce_bidir_bwd (one read of S, one write of dS)", bound="hbm", ms=ms_b,
                   achieved=8.0 * B * B / ms_b * 1e-6, unit="GB/s", algorithmic="8*B^2 bytes", traffic_key="ce_bwd_%d" % B,
                   fwd_plus_bwd_us=(ms_f + ms_b) * 1e3, baseline_md_target_us_for_12B2_bytes=12.0 * B * B / 6455.6e9 * 1e6))
    ids = torch.randint(0, 30000, (B, Lt), generator=g).to(dev)
    mask = torch.ones(B, Lt, dtype=torch.int64, device=dev)
    t_ids, t_mask = torch.empty_like(ids), torch.empty_like(mask)
    uc, up = torch.rand(B, generator=g).to(dev), torch.rand(B, generator=g).to(dev)
    lbl, src = torch.empty(B, dtype=torch.int64, device=dev), torch.empty(B, dtype=torch.int32, device=dev)
    rb = Lt * 8
    ms = time_kernel_graph(lambda: P.call("tic_itm_sample_gather", uc.data_ptr(), up.data_ptr(), B, 0, None, 0, 0.0, None, ids.data_ptr(),
                                    mask.data_ptr(), rb, t_ids.data_ptr(), t_mask.data_ptr(), lbl.data_ptr(), src.data_ptr(), st()),
                     flush)
    ks.append(dict(kernel="tic_itm_sample_gather (uniform rule + gather of ids and mask, one launch)", bound="hbm", ms=ms,
                   achieved=4.0 * B * rb / ms * 1e-6, unit="GB/s", algorithmic="2 row sets x (read + write) x B x 1024 bytes",
                   traffic_key="itm_sample_gather_%d" % B))
    ms = time_kernel_graph(lambda: P.call("tic_gather_rows", ids.data_ptr(), rb, t_ids.data_ptr(), rb, rb, src.data_ptr(), B, st()), flush)
    ks.append(dict(kernel="tic_gather_rows (one row set, 1024-byte rows)", bound="hbm", ms=ms, achieved=2.0 * B * rb / ms * 1e-6,
                   unit="GB/s", algorithmic="(read + write) x B x 1024 bytes", traffic_key="gather_rows_%d" % B))
    # hard-negative sampler, tile-stream form, P=512: weight sums ride on the forward tiles; locate; pick tiles
    d = 512
    T = torch.randn(B, d, generator=g).to(torch.bfloat16).to(dev)
    V = (torch.randn(B, d, generator=g) + 0.3 * T.float().cpu()).to(torch.bfloat16).to(dev)
    sc = math.exp(2.6592)
    it0, it1 = P.ItcPlan(B, B, d, dev), P.ItcPlan(B, B, d, dev, hard=True)
    for it in (it0, it1):
        it.norms(T, d, V, d)
    ms0 = time_kernel_graph(lambda: it0.fwd_tiles(T, d, V, d, sc), flush)
    ms1 = time_kernel_graph(lambda: it1.fwd_tiles(T, d, V, d, sc), flush)
    msl = time_kernel_graph(lambda: it1.hard_locate(uc, up, lbl, src), flush)
    msp = time_kernel_graph(lambda: it1.hard_pick(T, d, V, d, sc, src), flush)
    ks.append(dict(kernel="hard-negative sampler, tile-stream form (weight sums in tic_itc_fwd + tic_itm_hard_locate + tic_itc_pick)",
                   bound="hbm", ms=(ms1 - ms0) + msl + msp, achieved=(2.0 * B * it1.nqp * 8 + 20.0 * B) / ((ms1 - ms0) + msl + msp) * 1e-6,
                   unit="GB/s", algorithmic="8 bytes per (row, part) written + read, 20 bytes per row; S itself never reaches HBM",
                   fwd_tiles_plain_us=ms0 * 1e3, fwd_tiles_with_weight_sums_us=ms1 * 1e3, locate_us=msl * 1e3, pick_tiles_us=msp * 1e3,
                   materialised_alternative_bytes=8.0 * B * B))
    return ks


def scale_point(P, dev, flush, B=16384, d=768):
    """ITC-only step at B=16384, d=768 on one GPU (a c5 sweep point): the same kernels as the default workload, at a size
    where the tensor-core roofline applies.  Reported beside the default workload's (latency-bound) kernel numbers."""
    g = torch.Generator().manual_seed(41)
    T = torch.randn(B, d, generator=g).to(torch.bfloat16).to(dev)
    V = (torch.randn(B, d, generator=g) + 0.25 * T.float().cpu()).to(torch.bfloat16).to(dev)
    it = P.ItcPlan(B, B, d, dev)
    sums, rsum = torch.zeros(2, device=dev), torch.zeros(1, device=dev)
    dT, dV = torch.empty(B, d, device=dev), torch.empty(B, d, device=dev)
    scale = math.exp(2.6592)
    ld = T.stride(0)
    it.run(T, V, scale, 1.0, sums, rsum, dT_f32=dT, dV_f32=dV)
    ks = []
    ms = time_kernel(lambda: it.fwd_tiles(T, ld, V, ld, scale), flush)
    shp = "%dx%d" % (B, d)
    ks.append(dict(kernel="tic_itc_fwd", bound="tensor", ms=ms, achieved=2.0 * B * B * d / ms * 1e-9, unit="TFLOP/s",
                   algorithmic="2*B^2*d FLOP", traffic_key="itc_fwd_" + shp))
    ms = time_kernel(lambda: it.bwd_operands(T, ld, V, ld, scale, 1.0 / (2 * B)), flush)
    ks.append(dict(kernel="tic_itc_bwd_g", bound="tensor", ms=ms, achieved=2.0 * B * B * d / ms * 1e-9, unit="TFLOP/s",
                   algorithmic="recompute: 2*B^2*d executed FLOP, 0 algorithmic (reported as executed)",
                   traffic_key="itc_bwd_" + shp))
    ms = time_kernel(lambda: it.grad_gemms(T, ld, V, ld), flush)
    ks.append(dict(kernel="tic_gemm_bf16 x2 (dT, dV)", bound="tensor", ms=ms, achieved=4.0 * B * B * d / ms * 1e-9,
                   unit="TFLOP/s", algorithmic="4*B^2*d FLOP", traffic_key="gemm_dtdv_" + shp))
    ms = time_kernel(lambda: it.run(T, V, scale, 1.0, sums, rsum, dT_f32=dT, dV_f32=dV), flush, iters=5)
    step = dict(kernel="whole ITC fwd+bwd step (norms, tiles, lse, recompute, 2 GEMMs, finalize)", bound="tensor", ms=ms,
                achieved=6.0 * B * B * d / ms * 1e-9, unit="TFLOP/s", algorithmic="6*B^2*d FLOP (recompute not counted)",
                samples_per_s=B / (ms * 1e-3))
    del it
    return dict(workload="c5 point: ITC only, B=%d, d=%d, 1 GPU" % (B, d), kernels=ks, step=step,
                why="the default workload (B=256) is launch-latency-bound; this is the same path at a roofline-relevant size")


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench.py — mask-build throughput of the B200 pruning hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

Workload (BASELINE.json configs[1]): ResNet-50 SNIP to 90 % sparsity, 8 synthetic mini-batches of
gradients accumulated into the score, global k-th-smallest threshold, bit-packed mask emit.
One "step" = one complete mask build (score accumulate x 8 -> radix select -> emit).
`value` = prunable parameters masked per second (Gparams/s), inputs resident in HBM.
`e2e`   = the same metric through the host-buffer C-ABI entry point
          (b200p_snip_mask_build_host): weights and all gradient sets start in pinned HOST memory,
          the packed mask and the result block come back to the host inside the timed region.
`roofline` is for the dominant kernel of the step: k_score_multi (default --score-mode fused: all 8 resident
          gradient sets folded in one pass, 4*(B+2) B/param) or k_score_accumulate (--score-mode streaming,
          16 B/param/batch).
`cpu_baseline` / `--impl reference`: the reference's own torch-CPU operator sequence
          (oracle/torch_port.py) on the box's host cores.

N > 1 (launched under torchrun), default --dist-mode replicas: mask builds are independent units (one
per model replica / sparsity level), so every rank builds the mask of its own replica from its own 8
gradient sets with no data-path collective (weak scaling); value = N mask builds / max-over-ranks time.
--dist-mode sharded runs ONE mask build over the ranks (SURVEY §8e, strong scaling): local accumulate
-> NCCL all-to-all of score slices summed in rank order on the owner -> parameter-sharded select with
a histogram all-reduce per radix pass -> every rank emits its slice of the packed mask -> all-reduce of
the mask words.  For ResNet-50 that exchange (~100 MB per GPU) costs more than the whole single-GPU
build (0.26 ms), see DESIGN.md §4.  Timing is the max over ranks of CUDA-event time.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

MODEL = "resnet50"
TARGET_SPARSITY = 0.9
N_BATCHES = 8
SCORE_BYTES_PER_PARAM = 16.0          # read w, g, acc; write acc
SCORE_ASSIGN_BYTES_PER_PARAM = 12.0   # first batch: no acc read
STEP_BYTES_PER_PARAM = 16.0 * N_BATCHES + 8.125


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-lost", action="store_true", help="skip the LOST images/s leg")
    ap.add_argument("--dist-mode", default="replicas", choices=["replicas", "sharded"],
                    help="N > 1: replicas = every rank builds the mask of its own model replica from its own 8 gradient "
                         "sets (independent units, no collective, weak scaling); sharded = ONE mask build split over the "
                         "ranks (batches shard the scores, parameters shard the select; NCCL exchange; strong scaling)")
    ap.add_argument("--score-mode", default="sweep", choices=["sweep", "fused", "streaming"],
                    help="fused: one pass over all resident gradient sets (4*(B+2) B/param); "
                         "streaming: one accumulate launch per mini-batch (16 B/param/batch)")
    ap.add_argument("--no-clocks", action="store_true", help="skip the clock sampler and its keep-busy loops (ncu runs)")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            with open(path) as f:
                return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ---------------------------------------------------------------------------------------------
# synthetic workload
def model_numels():
    from pruning_for_vision_representation_b200.shapes import prunable_numels
    return prunable_numels(MODEL)


def make_weights_cpu(numels):
    """ResNet-50 default init, torch.manual_seed(1); synthetic fan-in-scaled
    normal weights of the same shapes if torchvision is unavailable."""
    import torch
    try:
        import torchvision
        torch.manual_seed(1)
        m = torchvision.models.get_model(MODEL, weights=None, num_classes=1000)
        ws = [mod.weight.detach().reshape(-1).clone() for mod in m.modules()
              if isinstance(mod, (torch.nn.Conv2d, torch.nn.Linear))]
        assert [w.numel() for w in ws] == list(numels)
        return torch.cat(ws), "torchvision resnet50 default init, seed 1"
    except Exception as e:      # pragma: no cover
        g = torch.Generator().manual_seed(1)
        return torch.randn(sum(numels), generator=g) * 0.02, f"0.02*randn (torchvision unavailable: {e})"


def make_grads(n_total, batch_ids, device):
    """g_b = 1e-3 * randn, seed 300+b, generated on `device` (values differ between cpu and cuda
    generators; each arm is self-consistent)."""
    import torch
    out = []
    for b in batch_ids:
        g = torch.Generator(device=device).manual_seed(300 + b)
        out.append(torch.randn(n_total, generator=g, device=device).mul_(1e-3))
    return out


def split_views(flat, numels):
    out, off = [], 0
    for n in numels:
        out.append(flat[off:off + n])
        off += n
    return out


# ---------------------------------------------------------------------------------------------
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(prefix="clocks_", suffix=".csv")
            os.close(fd)
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.QUERY}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        try:
            self.proc.terminate()
            self.proc.wait(timeout=5)
        except Exception:
            pass
        try:
            self.f.close()
            rows = [r.strip().split(",") for r in open(self.path) if r.strip()]
            os.unlink(self.path)
        except Exception:
            return out
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            r = [c.strip() for c in r]
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except ValueError:
                continue
            for name, col in zip(names, r[5:9]):
                if col.lower().startswith("active"):
                    reasons.add(name)
        if sm:
            if len(sm) > 3:
                sm = sm[2:]                          # the first samples may predate the load
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


# ---------------------------------------------------------------------------------------------
def cpu_reference_run(steps, warmup, numels, verbose=False):
    """Times oracle/torch_port.snip_mask_build (the reference's torch-CPU operator sequence) on the
    host cores.  Returns (Gparams/s, ms per step, sample description, cores)."""
    import torch
    from oracle import torch_port as TP
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    n_total = sum(numels)
    # bounded sample: the full parameter set costs ~5-8 s per mask build on a server CPU; when the
    # requested step count would run past a few minutes, use a leading subset of the tensors.
    budget_full_steps = 24
    frac = min(1.0, budget_full_steps / max(1, steps + warmup))
    use, acc = [], 0
    for n in numels:
        if acc >= frac * n_total and use:
            break
        use.append(n); acc += n
    n_used = sum(use)
    w_flat, wsrc = make_weights_cpu(numels)
    w = split_views(w_flat[:n_used], use)
    grads = [split_views(g, use) for g in make_grads(n_used, range(N_BATCHES), "cpu")]
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        masks, thr = TP.snip_mask_build(w, grads, TARGET_SPARSITY)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
        if verbose:
            print(f"[reference] step {i}: {dt:.3f} s thr={thr}", file=sys.stderr, flush=True)
    total = sum(times)
    sample = (f"{len(use)}/{len(numels)} prunable tensors ({n_used} of {n_total} params), {N_BATCHES} batches, "
              f"{steps} timed mask builds, torch {torch.__version__} CPU ops, {cores} threads; weights: {wsrc}")
    return n_used * len(times) / total / 1e9, 1e3 * total / len(times), sample, cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    numels = model_numels()
    val, ms, sample, cores = cpu_reference_run(args.steps, args.warmup, numels, verbose=True)
    line = {
        "impl": "reference", "metric": "mask-build Gparams/s (score+global top-k)", "value": val,
        "unit": "Gparams/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{MODEL} SNIP {TARGET_SPARSITY} sparsity, {N_BATCHES} mini-batches of synthetic "
                               "gradients, score accumulate + full sort threshold + mask (CPU, torch ops)"},
        "cpu_baseline": {"value": val, "unit": "Gparams/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "Gparams/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    GUARD.emit(json.dumps(line))


# ---------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist
    from pruning_for_vision_representation_b200 import _lib as L
    from pruning_for_vision_representation_b200.plan import ParamPlan

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    L.require_cuda()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")      # keep stdout for the one JSON line
        dist.init_process_group("nccl", device_id=dev)

    numels = model_numels()
    n_total = sum(numels)
    k = int(n_total * TARGET_SPARSITY)                      # train.py:299
    sharded = world > 1 and args.dist_mode == "sharded"
    if sharded:
        assert N_BATCHES % world == 0, "mini-batches must split evenly across ranks"
        my_batches = list(range(rank * N_BATCHES // world, (rank + 1) * N_BATCHES // world))
    else:
        my_batches = [b + 1000 * rank for b in range(N_BATCHES)]     # every rank: its own full set of 8 gradient sets

    w_flat_cpu, wsrc = make_weights_cpu(numels)
    w_flat = w_flat_cpu.to(dev)
    g_flat = make_grads(n_total, my_batches, dev)           # > L2: 8 x 102 MB streamed once per step
    s_flat = torch.empty(n_total, device=dev)
    plan = ParamPlan(numels, dev)
    plan.bind(L.SLOT_W, split_views(w_flat, numels)).bind(L.SLOT_SCORE, split_views(s_flat, numels))
    g_tables = [plan.pointer_table(L.SLOT_G, split_views(g, numels)) for g in g_flat]
    mask = plan.new_mask()

    if sharded:
        from pruning_for_vision_representation_b200.distributed import ShardedMaskBuilder
        builder = ShardedMaskBuilder(plan, dist.group.WORLD)

    launches = [0]
    score_events = []

    sweep = args.score_mode == "sweep" and not sharded      # score pass + bracket sweep in one kernel (b200p_snip_mask_build)
    fused = args.score_mode == "fused" or (args.score_mode == "sweep" and sharded)
    if sweep:
        plan.time_sweep(True)

    def step(record):
        if sweep:
            plan.snip_mask_build(g_tables, k, mask); launches[0] += 4      # sample, score+sweep, finish, emit
            return
        if fused:
            if record:
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                e0.record()
            plan.score_accumulate_multi(g_tables, accumulate=False); launches[0] += 1
            if record:
                e1.record(); score_events.append((True, e0, e1))
        else:
            for i, tbl in enumerate(g_tables):
                plan.bind_table(tbl)                     # host-side table swap, no launch
                if record:
                    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                    e0.record()
                plan.score_accumulate(i > 0); launches[0] += 1
                if record:
                    e1.record(); score_events.append((i > 0, e0, e1))
        if not sharded:
            plan.mask_build(L.KEY_SCORE, k, L.MODE_SNIP_STRICT, mask); launches[0] += 4      # sample, sweep, finish, emit
        else:
            launches[0] += builder.snip_select_emit(s_flat, k, mask)

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step(False)
    sync_all()
    # clocks: sampled every 100 ms while the same step loop keeps the GPU busy before, during and
    # after the (short) timed region, so every sample is taken under this load
    clocks = ClockSampler(local)
    if rank == 0 and not args.no_clocks:
        clocks.start()

    def keep_busy(seconds):
        if args.no_clocks:
            return
        t_end = time.time() + seconds
        flag = torch.zeros(1, device=dev)
        while True:
            for _ in range(8):
                step(False)
            torch.cuda.synchronize()
            flag[0] = 1.0 if time.time() < t_end else 0.0
            if world > 1:
                dist.broadcast(flag, 0)
            if flag.item() == 0.0:
                break

    keep_busy(0.5)
    launches[0] = 0
    sync_all()
    if sweep:
        plan.kernel_time_ms()                 # discard the launches timed during warm-up
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(args.steps):
        step(True)
    t1.record()
    sync_all()
    elapsed_ms = t0.elapsed_time(t1)
    n_launches = launches[0]
    if world > 1:
        t = torch.tensor([elapsed_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
        tl = torch.tensor([n_launches], device=dev, dtype=torch.int64)
        dist.all_reduce(tl)
        n_launches = int(tl.item())
    res = plan.result()
    sweep_ms, sweep_n = plan.kernel_time_ms() if sweep else (0.0, 0)     # the launches of the timed region only
    keep_busy(0.4)
    clk = clocks.stop() if rank == 0 else None

    acc_ms = [a.elapsed_time(b) for is_acc, a, b in score_events if is_acc]
    kernel_ms = sum(acc_ms) / max(1, len(acc_ms))
    n_timed = len(acc_ms)
    if sweep:
        kernel_ms, n_timed = sweep_ms, sweep_n           # CUDA events recorded by the library around k_snip_score_sweep, same stream
    peak, peak_src = peaks()
    nb_local = len(my_batches)
    # algorithmic bytes per launch of the dominant kernel: fused pass reads w and nb_local
    # gradient sets and writes the score once; the streaming pass reads w, g, acc and writes acc
    kernel_bytes = n_total * (4.0 * (nb_local + 2) + 0.125 if sweep else 4.0 * (nb_local + 2) if fused else SCORE_BYTES_PER_PARAM)
    kernel_name = (f"k_snip_score_sweep<ACCUMULATE=0,B={nb_local}>" if sweep else
                   f"k_score_multi<ACCUMULATE=0,B={nb_local}>" if fused else "k_score_accumulate<ACCUMULATE=1,VEC=1>")
    # whole step: the fused score+sweep never re-reads the scores (no 4 B select read, no 4 B emit read)
    step_bytes_per_param = (4.0 * (N_BATCHES + 2) + 0.125 if sweep else
                            (4.0 * (N_BATCHES + 2) if fused else 16.0 * N_BATCHES) + 8.125)
    achieved = kernel_bytes / (kernel_ms * 1e-3) / 1e9
    ms_per_step = elapsed_ms / args.steps
    units = 1 if sharded else world                      # mask builds completed per step over all ranks
    value = units * n_total / (ms_per_step * 1e-3) / 1e9

    # ---- e2e: host buffers through the C-ABI (rank-local; N = 1 headline) -----------------------
    e2e = None
    if not args.no_e2e:
        nb = len(my_batches)
        w_host = w_flat_cpu.pin_memory()
        g_host = [g.cpu().pin_memory() for g in g_flat]
        mask_host = torch.empty(plan.mask_words, dtype=torch.int32).pin_memory()
        if not sharded:
            for _ in range(1):
                plan.snip_mask_build_host(w_host, g_host, k, mask_host)
            sync_all()
            tt = time.perf_counter()
            for _ in range(args.e2e_steps):
                r2 = plan.snip_mask_build_host(w_host, g_host, k, mask_host)    # synchronises
            e2e_s = (time.perf_counter() - tt) / args.e2e_steps
            assert torch.equal(mask_host, mask.cpu()), "host-buffer path and resident path disagree"
            if world > 1:
                t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                e2e_s = float(t.item())
            e2e = {"value": world * n_total / e2e_s / 1e9, "unit": "Gparams/s",
                   "h2d_bytes_per_step": world * (1 + nb) * n_total * 4,
                   "d2h_bytes_per_step": world * (plan.mask_words * 4 + 64),
                   "ms_per_step": e2e_s * 1e3, "steps": args.e2e_steps,
                   "api": "b200p_snip_mask_build_host (pinned host buffers in, packed mask out)"}
        else:
            e2e_ms = builder.e2e_host_steps(w_host, g_host, k, mask_host, args.e2e_steps)
            e2e = {"value": n_total / (e2e_ms * 1e-3) / 1e9, "unit": "Gparams/s",
                   "h2d_bytes_per_step": (1 + nb) * n_total * 4 * world,
                   "d2h_bytes_per_step": (plan.mask_words * 4 + 64) * world,
                   "ms_per_step": e2e_ms, "steps": args.e2e_steps,
                   "api": "ShardedMaskBuilder.snip_mask_build_host (pinned host buffers in, packed mask out)"}

    lost = None
    if not args.no_lost:
        lost = lost_leg(args, dev, world, rank, dist)

    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, ms, sample, cores = cpu_reference_run(2, 1, numels)
        cpu_base = {"value": v, "unit": "Gparams/s", "cores": cores, "kind": "port", "sample": sample,
                    "ms_per_step": ms}

    if rank == 0:
        line = {
            "metric": "mask-build Gparams/s (score+global top-k)", "value": value, "unit": "Gparams/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "strong" if sharded else "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{MODEL} SNIP mask build, target sparsity {TARGET_SPARSITY}, {N_BATCHES} mini-batches "
                                   f"of synthetic gradients (1e-3*randn) accumulated, N={n_total} params in {len(numels)} tensors, "
                                   f"k={k}", "weights": wsrc,
                       "score_mode": ("sweep: one pass folds all resident gradient sets into the score (bit-identical to per-batch "
                                      "accumulation) and classifies it against the sampled bracket; the scores are never read back"
                                      if sweep else
                                      "fused: all resident gradient sets folded in one pass, bit-identical to per-batch accumulation"
                                      if fused else "streaming: one accumulate launch per mini-batch"),
                       "l2": f"inputs larger than L2: {N_BATCHES // world} gradient sets x {n_total * 4 >> 20} MiB streamed per step",
                       "parallelism": ("1 GPU" if world == 1 else
                                       f"batches split over {world} ranks, NCCL all-to-all score exchange + rank-order sum, "
                                       "parameter-sharded radix select with histogram all-reduce (one mask build, strong scaling)" if sharded
                                       else f"{world} independent mask builds, one model replica with its own 8 gradient sets per "
                                            "rank, no data-path collective (weak scaling); --dist-mode sharded runs the NCCL path")},
            "roofline": {"bound": "hbm", "kernel": kernel_name, "achieved": achieved,
                         "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": ncu_traffic("k_snip_score_sweep" if sweep else "k_score_multi" if fused else "k_score_accumulate"),
                         "peak_source": peak_src, "kernel_ms": kernel_ms,
                         "algorithmic_bytes_per_launch": kernel_bytes,
                         "algorithmic_bytes_per_param": kernel_bytes / n_total,
                         "launches_timed": n_timed,
                         "step_algorithmic_bytes_per_param": step_bytes_per_param,
                         "step_algorithmic_GBps": (n_total * step_bytes_per_param / (ms_per_step * 1e-3) / 1e9
                                                   if world == 1 else None)},
            "e2e": e2e, "cpu_baseline": cpu_base, "gpu_launches": n_launches, "clocks": clk, "lost": lost,
            "result": {"threshold": res["threshold"], "n_less": res["n_less"], "n_equal": res["n_equal"],
                       "n_kept": res["n_kept"], "passes_full": res["passes_full"]},
        }
        GUARD.emit(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------
LOST_N, LOST_D, LOST_DIMS, LOST_B = 900, 384, [30, 30], 256
LOST_FLOP_PER_IMAGE = 2.0 * LOST_N * LOST_N * LOST_D          # 622 080 000


def lost_leg(args, dev, world, rank, dist):
    """LOST ViT-S/16 images/s (BASELINE.json configs[2]): synthetic patch keys randn(B, 900, 384), seed 0
    (+ rank), dims 30x30, scales 16, image 480x480, k_patches 100.  Images shard over ranks, no collective.
    value = images/s with the keys resident in HBM; e2e = keys in pinned host memory -> boxes on the host."""
    import numpy as np
    import torch
    from pruning_for_vision_representation_b200 import object_discovery as OD
    g = torch.Generator().manual_seed(rank)
    feats_host = torch.randn(LOST_B, LOST_N, LOST_D, generator=g).pin_memory()
    feats = feats_host.to(dev)
    size = (3, 480, 480)
    run = lambda f: OD.lost_batched(f, LOST_DIMS, [16, 16], size, k_patches=100)
    for _ in range(3):
        out = run(feats)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    steps = max(5, min(50, args.steps))
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(steps):
        out = run(feats)
    t1.record()
    torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / steps
    # e2e: pinned host keys in, boxes / seeds / status back on the host
    box_h = torch.empty(LOST_B, 4).pin_memory(); seed_h = torch.empty(LOST_B, dtype=torch.int32).pin_memory()
    st_h = torch.empty(LOST_B, dtype=torch.int32).pin_memory()
    def e2e_once():
        f = feats_host.to(dev, non_blocking=True)
        o = run(f)
        box_h.copy_(o["box"], non_blocking=True); seed_h.copy_(o["seed"], non_blocking=True); st_h.copy_(o["status"], non_blocking=True)
        torch.cuda.synchronize()
    e2e_once()
    tt = time.perf_counter()
    for _ in range(3):
        e2e_once()
    e2e_ms = (time.perf_counter() - tt) / 3 * 1e3
    if world > 1:
        t = torch.tensor([ms, e2e_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_ms = float(t[0]), float(t[1])
    peak_tf32 = None
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peak_tf32 = float(json.load(f)["bf16_tflops"]) / 2.0
    except Exception:
        peak_tf32 = 1590.0 / 2.0
    img_s = world * LOST_B / (ms * 1e-3)
    achieved = img_s / world * LOST_FLOP_PER_IMAGE / 1e12
    # flop the tensor cores execute per image: upper-triangular 256x256 tiles (256 rows each), the last column tile
    # trimmed to the next multiple of 16 columns, three tf32 MMAs per product (3xTF32)
    t2 = (LOST_N + 255) // 256
    last_cols = min(256, (LOST_N - (t2 - 1) * 256 + 15) // 16 * 16)
    exec_elems = sum(256 * (256 if tj < t2 - 1 else last_cols) for ti in range(t2) for tj in range(ti, t2))
    exec_flop = 3 * 2 * exec_elems * LOST_D
    prof = {}
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            prof = json.load(f)
    except Exception:
        pass
    leg = {"metric": "LOST ViT-S/16 images/sec", "value": img_s, "unit": "images/s", "ms_per_step": ms, "steps": steps,
           "config": {"workload": f"LOST on synthetic patch keys randn({LOST_B},{LOST_N},{LOST_D}) per GPU, dims 30x30, "
                                  "k_patches 100, keys -> Gram -> degree -> seed -> expansion -> box; ViT forward excluded",
                      "l2": f"{LOST_B} Gram matrices = {LOST_B * LOST_N * LOST_N * 4 >> 20} MiB written per step (> L2)"},
           "roofline": {"bound": "tensor", "kernel": "k_lost_gram_tc2<direct> (TMA on the caller's keys + tcgen05.mma cta_group::2 kind::tf32, 3xTF32 with "
                                                     "the lo tiles derived in shared memory, TMEM epilogue with fused degree)",
                        "achieved": achieved, "peak": peak_tf32, "unit": "TFLOP/s", "frac": achieved / peak_tf32,
                        "executed_tflops": img_s / world * exec_flop / 1e12,
                        "executed_frac": img_s / world * exec_flop / 1e12 / peak_tf32,
                        "tensor_pipe_active_pct_ncu": prof.get("k_lost_gram_tc2_tensor_pipe_active_pct"),
                        "traffic": prof.get("k_lost_gram_tc2_traffic_bytes"),
                        "peak_source": "0.5 x measured bf16 burst (nominal tf32:bf16 ratio; no measured tf32 peak); achieved counts the "
                                       "algorithmic 2*N^2*d flop per image over the WHOLE LOST step (Gram + finish + launches); executed_tflops "
                                       "counts what the tensor cores run: 3 tf32 MMAs per product on the upper-triangular 256-row tiles"},
           "e2e": {"value": world * LOST_B / (e2e_ms * 1e-3), "unit": "images/s", "ms_per_step": e2e_ms,
                   "h2d_bytes_per_step": LOST_B * LOST_N * LOST_D * 4, "d2h_bytes_per_step": LOST_B * (16 + 4 + 4)},
           "gpu_launches_per_step": 3 + (LOST_B + 255) // 256,   # set_meta x ceil(B/256), tile table, Gram, finish (+ 1 memset node)
           "seed0_box0": [int(out["seed"][0].item()), out["box"][0].tolist()]}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import lost_oracle as LO
        n_img = 24
        fs = feats_host[:n_img].numpy()
        LO.lost(fs[0], LOST_DIMS, [16, 16], size, 100)
        tt = time.perf_counter()
        for i in range(n_img):
            LO.lost(fs[i], LOST_DIMS, [16, 16], size, 100)
        dt = time.perf_counter() - tt
        leg["cpu_baseline"] = {"value": n_img / dt, "unit": "images/s", "cores": os.cpu_count(), "kind": "port",
                               "sample": f"{n_img} of the same images through oracle/lost_oracle.lost (numpy BLAS Gram + python flood fill)"}
    return leg


def ncu_traffic(kernel):
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture
    (profiles/roofline_traffic.json), or None."""
    p = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    try:
        with open(p) as f:
            return json.load(f).get(kernel + "_traffic_bytes")
    except Exception:
        return None


class StdoutGuard:
    """Everything libraries print to fd 1 (e.g. NCCL's version banner) goes to stderr; stdout carries only
    the one JSON line the driver parses."""

    def __init__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)

    def emit(self, text):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        print(text, flush=True)
        os.dup2(2, 1)


GUARD = None


def main():
    global GUARD
    GUARD = StdoutGuard()
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()

"""Per-kernel timing on a synthetic parameter set (CUDA events, L2 flushed between launches)."""
import argparse
import json
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pruning_for_vision_representation_b200 import _lib as L
from pruning_for_vision_representation_b200.plan import ParamPlan
from pruning_for_vision_representation_b200.shapes import prunable_numels


def timeit(fn, flush, iters=10):
    ts = []
    for _ in range(iters):
        flush()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e-3)
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="resnet50")
    ap.add_argument("--sparsity", type=float, default=0.9)
    ap.add_argument("--no-flush", action="store_true")
    a = ap.parse_args()
    dev = "cuda:0"
    numels = prunable_numels(a.model)
    N = sum(numels)
    plan = ParamPlan(numels, dev)
    g = torch.Generator(device=dev); g.manual_seed(1)
    mk = lambda scale=1.0: [torch.randn(n, device=dev, generator=g) * scale for n in numels]
    w, gr, sc, buf = mk(0.02), mk(1e-3), mk(), mk()
    weff = [torch.empty(n, device=dev) for n in numels]
    w16 = [torch.empty(n, device=dev, dtype=torch.bfloat16) for n in numels]
    plan.bind(L.SLOT_W, w).bind(L.SLOT_G, gr).bind(L.SLOT_SCORE, sc).bind(L.SLOT_BUF, buf).bind(L.SLOT_WEFF, weff).bind(L.SLOT_WEFF16, w16)
    junk = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    flush = (lambda: None) if a.no_flush else (lambda: junk.fill_(1))
    mask = plan.new_mask(); old = plan.new_mask()
    k = int(N * a.sparsity)
    out = {"model": a.model, "N": N, "segments": len(numels)}

    def rep(name, fn, bytes_per_param):
        med, best = timeit(fn, flush)
        out[name] = {"us": round(med * 1e6, 1), "best_us": round(best * 1e6, 1), "GBps": round(N * bytes_per_param / med / 1e9, 1),
                     "Gparams_s": round(N / med / 1e9, 2)}
        print(name, out[name], flush=True)

    rep("score_assign", lambda: plan.score_accumulate(False), 12)
    rep("score_accumulate", lambda: plan.score_accumulate(True), 16)
    plan.score_accumulate(False)
    def pass0():
        plan.select_begin(k, L.MODE_SNIP_STRICT); plan.select_hist(0, L.KEY_SCORE)
    rep("select_pass0_hist(+init)", pass0, 4)
    rep("select_kth_score", lambda: plan.select_kth(L.KEY_SCORE, k, L.MODE_SNIP_STRICT), 4)
    print(plan.result())
    rep("emit_snip", lambda: plan.emit_masks(L.KEY_SCORE, L.MODE_SNIP_STRICT, mask), 4.125)
    rep("select_kth_absw", lambda: plan.select_kth(L.KEY_ABS_W, N // 2, L.MODE_EXACT_K), 4)
    print(plan.result())
    rep("emit_magnitude", lambda: plan.emit_masks(L.KEY_ABS_W, L.MODE_EXACT_K, mask), 4.125)
    def mag():
        plan.select_kth(L.KEY_ABS_W, N // 2, L.MODE_EXACT_K); plan.emit_masks(L.KEY_ABS_W, L.MODE_EXACT_K, mask)
    rep("magnitude_mask_build", mag, 8.125)
    old.copy_(mask)
    def mag2():
        plan.select_kth(L.KEY_ABS_W, N // 10, L.MODE_EXACT_K, old); plan.emit_masks(L.KEY_ABS_W, L.MODE_EXACT_K, mask, old)
    rep("magnitude_round2", mag2, 8.375)
    rep("masked_sgd_bf16", lambda: plan.masked_sgd_step(mask, 0.1, 0.9, 0.0, 1e-4, L.SGD_EMIT_WEFF16), 22.125)
    rep("masked_sgd_f32", lambda: plan.masked_sgd_step(mask, 0.1, 0.9, 0.0, 1e-4, L.SGD_EMIT_WEFF), 24.125)
    rep("count_zeros", lambda: plan.lib.b200p_count_zeros(plan.handle, mask.data_ptr(), plan._counts.data_ptr(), 1, torch.cuda.current_stream().cuda_stream), 4.125)
    cp_src = torch.empty(N, device=dev); cp_dst = torch.empty(N, device=dev)
    rep("torch_copy(ref)", lambda: cp_dst.copy_(cp_src), 8)
    print(json.dumps(out))


if __name__ == "__main__":
    main()

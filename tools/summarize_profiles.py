"""Turn the scratch captures of tools/gpu_profile_round.sh (gpurun_out/) into the committed summaries under profiles/:
launch list -> per-kernel shares, `ncu --set full` reports -> one table of the metrics DESIGN.md quotes, and the
per-launch DRAM traffic / tensor-pipe figures bench.py reads (profiles/roofline_traffic.json)."""
import csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out"); PROF = os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r1_v3"


def launches(path):
    rows = [r for r in csv.reader(l for l in open(path) if not l.startswith("=="))]
    h = rows[0]; ki, vi = h.index("Kernel Name"), h.index("Metric Value")
    agg = {}
    for r in rows[1:]:
        name = r[ki].split("(")[0].replace("void ", "").replace("b200p::", "")
        t = agg.setdefault(name, [0, 0.0]); t[0] += 1; t[1] += float(r[vi]) / 1e3
    return agg


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h = rows[0]
    return [dict(zip(h, r)) for r in rows[2:]]


def f(d, k, default=0.0):
    try:
        return float(d.get(k, default) or default)
    except ValueError:
        return default


WANT = [("us", "gpu__time_duration.sum"), ("dram read MB", "dram__bytes_read.sum"), ("dram write MB", "dram__bytes_write.sum"),
        ("dram % (read+write)", "DRAMPCT"), ("L2 %", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
        ("tensor pipe active %", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"),
        ("smem pipe: tensor %", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"),
        ("smem pipe: lsu %", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"),
        ("regs", "launch__registers_per_thread"), ("grid", "launch__grid_size"), ("warp instr", "smsp__inst_executed.sum")]

agg = launches(os.path.join(OUT, "launches_final.csv"))
total = sum(t for _, t in agg.values())
with open(os.path.join(PROF, f"{tag}_launch_summary.md"), "w") as o:
    o.write(f"# ncu launch list, {tag}\n\nCommand: `ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 300 --csv python bench.py "
            "--steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-clocks --no-extra` (run after the same command exited 0 without ncu).\n"
            "Per-launch times are cold-cache and serialised: compare shares, not absolutes.\n\n| kernel | launches | total us | mean us | share |\n|---|---|---|---|---|\n")
    for name, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        o.write(f"| `{name}` | {c} | {t:.1f} | {t / c:.1f} | {100 * t / total:.1f} % |\n")
    snip = {k: v for k, v in agg.items() if any(s in k for s in ("k_snip", "k_select", "k_emit", "k_score"))}
    st = sum(t for _, t in snip.values())
    dom = max(snip.items(), key=lambda kv: kv[1][1])
    o.write(f"\nMask-build step only: `{dom[0]}` share = {100 * dom[1][1] / st:.1f} % of {st / dom[1][0]:.1f} us per step (cold, serialised).\n")
os.replace(os.path.join(OUT, "launches_final.csv"), os.path.join(PROF, f"{tag}_launches_bench.csv")) if "--move" in sys.argv else None

rows = []
for rep in ("prof_final_snip.ncu-rep", "prof_final_lost.ncu-rep", "prof_final_select.ncu-rep"):
    p = os.path.join(OUT, rep)
    if os.path.exists(p):
        rows += raw(p)
with open(os.path.join(PROF, f"{tag}_ncu_full_summary.md"), "w") as o:
    o.write(f"# ncu --set full capture, {tag}\n\nCommands (each after the same command exited 0 without ncu): see `tools/gpu_profile_round_r2.sh`.\n\n")
    o.write("| kernel | " + " | ".join(n for n, _ in WANT) + " |\n|---|" + "---|" * len(WANT) + "\n")
    for d in rows:
        name = d["Kernel Name"].split("(")[0].replace("void ", "")
        d["DRAMPCT"] = f(d, "dram__bytes_read.sum.pct_of_peak_sustained_elapsed") + f(d, "dram__bytes_write.sum.pct_of_peak_sustained_elapsed")
        o.write(f"| `{name}` | " + " | ".join(f"{f(d, k):.4g}" for _, k in WANT) + " |\n")

traffic_path = os.path.join(PROF, "roofline_traffic.json")
tr = json.load(open(traffic_path))
def unit_bytes(d, k):
    # raw page reports Mbyte / Gbyte depending on size; normalise with the unit row is not available here: values are Mbyte for these kernels
    return f(d, k) * 1e6
sw = [d for d in rows if "k_snip_score_sweep" in d["Kernel Name"]]
if sw:
    tr["k_snip_score_sweep_traffic_bytes"] = sum(unit_bytes(d, "dram__bytes_read.sum") + unit_bytes(d, "dram__bytes_write.sum") for d in sw) / len(sw)
    tr["k_snip_score_sweep_source"] = f"profiles/{tag}_ncu_full_summary.md (ncu --set full, dram read + write, mean of {len(sw)} launches)"
lg = [d for d in rows if "k_lost_gram_tc2<1, 0" in d["Kernel Name"] or "k_lost_gram_tc2<true, false" in d["Kernel Name"]] or [d for d in rows if "k_lost_gram_tc2" in d["Kernel Name"]]
if lg:
    tr["k_lost_gram_tc2_traffic_bytes"] = sum(unit_bytes(d, "dram__bytes_read.sum") + unit_bytes(d, "dram__bytes_write.sum") for d in lg) / len(lg)
    tr["k_lost_gram_tc2_tensor_pipe_active_pct"] = sum(f(d, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed") for d in lg) / len(lg)
    tr["k_lost_gram_tc2_source"] = f"profiles/{tag}_ncu_full_summary.md (tools/lost_probe2.py 256 2, count-only launches, 256 images of 900 x 384 keys per launch)"
json.dump(tr, open(traffic_path, "w"), indent=1)
print(open(os.path.join(PROF, f"{tag}_ncu_full_summary.md")).read())
print(json.dumps(tr, indent=1))

"""Turn the round's gpurun_out/ captures into the committed summaries under profiles/ (ncu raw-page metrics of one
launch, the per-kernel launch list of the bench command, traffic.json).  usage: make_profiles.py"""
import csv, json, os, subprocess, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out"); P = os.path.join(ROOT, "profiles")
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum", "lts__t_sector_hit_rate.pct", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio"]
def ncu_csv(rep, out, header):
    raw = subprocess.run(["ncu", "-i", os.path.join(G, rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(raw.splitlines())); h, u, v = r[0], r[1], r[2]
    with open(os.path.join(P, out), "w") as f:
        f.write("# " + header + "\n# NOTE: under ncu nanosleep returns at once, so the waiting warps' back-off loops spin: instruction and issue\n"
                "# counts include that polling (it does not run in a normal launch, see DESIGN.md 4a); durations are under the profiler.\n")
        f.write("metric,unit,value\n")
        f.write(f"Kernel Name,,{v[h.index('Kernel Name')]}\n")
        for k in KEYS:
            if k in h: f.write(f"{k},{u[h.index(k)]},{v[h.index(k)]}\n")
    d = {k: v[h.index(k)] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum") if k in h}
    units = {k: u[h.index(k)] for k in d}
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}
    return int(sum(float(d[k]) * scale[units[k]] for k in d))
traffic = {}
traffic["c4"] = ncu_csv("r02_k4_c4_b32.ncu-rep", "r02_ncu_k_score4_c4_b32.csv", "ncu --set full --clock-control none, one k_score4<NC=1> launch; C4: 10M nodes, 3.0e8 mutations, 32 snv40 samples per launch (scripts/one_launch.py c4 0 32 32 2 1)")
traffic["c4_96_per_launch"] = ncu_csv("r02_k4_c4_b96.ncu-rep", "r02_ncu_k_score4_c4_b96.csv", "same, k_score4<NC=3>: 96 samples per launch, three groups share one scan (scripts/one_launch.py c4 0 96 96 2 3)")
traffic["c3"] = ncu_csv("r02_k4_c3_b256.ncu-rep", "r02_ncu_k_score4_c3_b256.csv", "same, C3: 2M-node SARS-CoV-2-shaped MAT (41.6 MB), 256 leaf-derived samples per launch = 3 scan groups of k_score4<NC=3> (scripts/one_launch.py c3 1 256 256 2 0)")
traffic["c4_leaf"] = ncu_csv("r02_k4_c4_leaf.ncu-rep", "r02_ncu_k_score4_c4_leaf.csv", "same, C4 tree, 32 leaf-derived samples (~450 calls each): the dense hit phase (scripts/one_launch.py c4 1 32 32 2 1)")
json.dump(traffic, open(os.path.join(P, "traffic.json"), "w"), indent=1)
# launch list of the bench command
rows = list(csv.reader(l for l in open(os.path.join(G, "r_launches.csv")) if not l.startswith("==")))
h = rows[0]; ik = h.index("Kernel Name"); iv = h.index("Metric Value"); iu = h.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    if len(r) <= iv: continue
    val = float(r[iv].replace(",", "")); val *= {"us": 1e-3, "ns": 1e-6, "ms": 1.0, "usecond": 1e-3, "nsecond": 1e-6, "msecond": 1.0}.get(r[iu], 1e-3)
    a = agg.setdefault(r[ik], [0, 0.0]); a[0] += 1; a[1] += val
tot = sum(a[1] for a in agg.values())
with open(os.path.join(P, "r02_launches_c4_bench.csv"), "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none -c 600: python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra (C4, 32 samples per launch); aggregated per kernel\n")
    f.write("kernel,launches,total_ms,share\n")
    for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]): f.write(f"\"{k}\",{n},{ms:.3f},{ms/tot:.4f}\n")
print(open(os.path.join(P, "r02_launches_c4_bench.csv")).read()); print(traffic)

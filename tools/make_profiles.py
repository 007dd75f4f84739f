"""Turn the ncu outputs in gpurun_out/ into the committed summaries under profiles/ (round tag r01)."""
import csv, io, json, os, shutil, subprocess, sys
from collections import defaultdict
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out"); P = os.path.join(ROOT, "profiles")
shutil.copy(os.path.join(G, "r01_launches_final.csv"), os.path.join(P, "r01_launches.csv"))
bench = json.load(open(os.path.join(P, "r01_bench_n1.json")))
out = ["# Round 1: ncu summaries of the final build (B200, config 2: 3 014 484 reads x 150 bp, automatic block size, primed blocks)", "",
       "Command profiled: `python bench.py --steps 1 --warmup 1 --no-cpu`, each ncu pass run only after the same command had exited 0 without ncu.", "",
       "## Launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`, profiles/r01_launches.csv; cold-cache, serialised)", ""]
rows = [r for r in csv.reader(open(os.path.join(G, "r01_launches_final.csv"))) if len(r) > 10 and r[0].isdigit()]
agg = defaultdict(lambda: [0, 0.0])
for r in rows:
    k = r[4].split("(")[0].replace("void ", ""); agg[k][0] += 1; agg[k][1] += float(r[-1])
tot = sum(v[1] for v in agg.values())
out += ["| kernel | launches | total ms | share |", "|---|---|---|---|"]
out += [f"| {k} | {v[0]} | {v[1] / 1e6:.3f} | {100 * v[1] / tot:.1f}% |" for k, v in sorted(agg.items(), key=lambda x: -x[1][1])]
sm = bench["stage_ms"]
k2 = sm["k2e"] + sm["k2d"]
out += ["", f"bench.py's CUDA-event stage times for the same build (profiles/r01_bench_n1.json): K2 encode {sm['k2e']:.2f} ms, K2 decode {sm['k2d']:.2f} ms, "
        f"K1 {sm['k1']:.2f} ms, K3 {sm['k3']:.2f} ms of a {bench['ms_per_step']:.1f} ms step: the block coder (with its generation merges) is "
        f"{100 * k2 / bench['ms_per_step']:.0f} % of the step there and {100 * sum(v[1] for k, v in agg.items() if 'k2_' in k or 'merge' in k or 'snapshot' in k) / tot:.0f} % of the ncu launch list.", ""]
traffic = {}
def val(r, idx, units, m):
    return float(r[idx[m]]) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[units[idx[m]]]
for f, names in (("r01c_k1k3", {"k1_extract": "k1_extract_kernel", "k3_reconstruct": "k3_reconstruct_kernel"}),
                 ("r01c_k2e", {"k2_coder": "k2_coder_kernel<encode>"}), ("r01c_k2d", {"k2_coder": "k2_coder_kernel<decode>"})):
    rep = os.path.join(G, f + ".ncu-rep")
    out += [f"## `ncu --set full --clock-control none`: gpurun_out/{f}.ncu-rep", "```"]
    out += subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep], capture_output=True, text=True).stdout.rstrip().split("\n")
    out += ["```", ""]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(io.StringIO(raw))); hdr, units = rr[0], rr[1]; idx = {h: i for i, h in enumerate(hdr)}
    for r in rr[2:]:
        for pat, key in names.items():
            if pat in r[idx["Kernel Name"]] and key not in traffic:
                traffic[key] = int(val(r, idx, units, "dram__bytes_read.sum") + val(r, idx, units, "dram__bytes_write.sum"))
traffic["_note"] = ("dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full, config 2; the K2 entries are the "
                    "last-generation launch (2959 of 5007 blocks, 92 % of the reads)")
json.dump(traffic, open(os.path.join(P, "traffic.json"), "w"), indent=1)
out += ["The launch list is of the whole bench command: the device-resident steps (5 coder launches per direction: generations 0-4; K1 in two",
        "launches, the reads of the early generations first, the rest of the batch on a side stream beside them), three one-stream encodes that",
        "time K1 alone for its roofline entry, and the pipelined host-buffer steps (cbcg_encode / cbcg_decode: one K1 launch per chunk, one coder",
        "and one K3 launch per group).", "",
        "Reading: K2 (`k2_coder_kernel<mode, legacy>`) is serial integer work per block: 20 warps per SM (96 registers, no spills), 56 % of issue",
        "slots, 20 % + 11 % of stall samples waiting for instruction fetch (95 KB of SASS against a 32 KB L1.5 instruction cache); DRAM < 2 % of peak.",
        "Its DRAM traffic fell from 1.14 GB to 0.49 GB per last-generation launch with deferred var rows (rows touched once are coded from",
        "the snapshot and never copied). K3 runs at 30 % of the measured HBM peak and is bound by instruction issue (60 % of slots, 58 warp",
        "instructions per read); K1 at 19 %, with a fifth of its stall samples at the barrier behind the look-back over tile edit counts",
        "(the wait for every earlier in-flight tile to have counted its edits) and 97 warp instructions per read."]
open(os.path.join(P, "r01_ncu_summary.md"), "w").write("\n".join(out) + "\n")
print(traffic)

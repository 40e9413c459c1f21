"""Turns an ncu report of the bulk bucket kernel into the tracked summaries under profiles/:
    python tools/prof_summary.py gpurun_out/prof_r01d.ncu-rep r01d <strings in the launch> <steps>
writes profiles/ncu_full_<tag>_bulk_kernel_summary.json, ..._hot_lines.csv (SASS samples of the report joined with
`nvdisasm --print-line-info` of the library that was profiled -- it must be the current build) and ncu_traffic_<tag>.json."""
import collections, csv, io, json, os, re, subprocess, sys, tempfile

rep, tag, strings, steps = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "torch_fdtd_string_b200", "libsfdtd.so")
SRC = os.path.join(ROOT, "torch_fdtd_string_b200", "csrc", "sfdtd.cu")

def ncu(page):
    return list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout)))

raw = ncu("raw")
hdr, units, vals = raw[0], raw[1], raw[2]
d = dict(zip(hdr, vals)); u = dict(zip(hdr, units))
f = lambda k: float(d[k].replace(",", ""))
scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}[u["dram__bytes_read.sum"]]
ss = strings * steps
st = lambda n: f(f"smsp__average_warps_issue_stalled_{n}_per_issue_active.ratio")
summary = {
    "report": os.path.basename(rep), "kernel": d["Kernel Name"], "grid": d["Grid Size"], "block": d["Block Size"],
    "strings": strings, "steps": steps, "string_steps": ss, "duration_ms": f("gpu__time_duration.sum"),
    "registers_per_thread": int(f("launch__registers_per_thread")), "ctas_per_sm_limit_registers": f("launch__occupancy_limit_registers"),
    "warp_instructions": int(f("smsp__inst_executed.sum")), "warp_instructions_per_string_step": f("smsp__inst_executed.sum") / ss,
    "issue_slots_busy_pct": f("sm__inst_issued.avg.pct_of_peak_sustained_active"),
    "fp64_pipe_pct": f("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
    "fma_pipe_pct": (f("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active") if "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active" in d else None),
    "lsu_pipe_pct": f("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
    "alu_pipe_pct": f("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
    "smem_wavefronts_per_string_step": f("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum") / ss,
    "smem_bank_conflict_wavefronts_pct": 100 * f("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum") / f("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
    "dram_bytes_read": f("dram__bytes_read.sum") * scale, "dram_bytes_write": f("dram__bytes_write.sum") * scale,
    "dram_bytes_per_string_step": (f("dram__bytes_read.sum") + f("dram__bytes_write.sum")) * scale / ss,
    "stall_cycles_per_issued_instruction": {k: st(k) for k in ("wait", "short_scoreboard", "not_selected", "no_instruction",
                                                              "math_pipe_throttle", "dispatch_stall", "branch_resolving", "long_scoreboard", "barrier")},
}
# ---- source-line aggregation ----
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", LIB], cwd=tmp, capture_output=True)
cubin = [x for x in os.listdir(tmp) if x.endswith(".cubin")][0]
li = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
ty = {"double": "d", "float": "f"}
BOOL = r"(?:, \(?(?:bool\))?(\d|true|false))?"              # the PF template argument (absent in captures before it existed)
pf = lambda v: "1" if v in ("1", "true") else "0"
if "group_kernel" in d["Kernel Name"]:
    m = re.search(r"group_kernel<(double|float), \(?(?:int\))?(\d)" + BOOL + ">", d["Kernel Name"])
    mangled = "sfdtd_group_kernelI" + ty[m.group(1)] + "Li" + m.group(2) + "ELb" + pf(m.group(3)) + "E"
else:
    kn = re.search(r"<(double|float), \(?(?:int\))?(\d+), \(?(?:int\))?(\d+), \(?(?:int\))?(\d+), \(?(?:int\))?(\d)" + BOOL + ">", d["Kernel Name"]).groups()
    mangled = f"sfdtd_step_kernelI{ty[kn[0]]}Li{kn[1]}ELi{kn[2]}ELi{kn[3]}ELi{kn[4]}ELb{pf(kn[5])}EEE"
start = [i for i, l in enumerate(li) if l.startswith(".text.") and mangled in l][0]
end = next(i for i in range(start + 1, len(li)) if li[i].startswith("//-----"))
cur, lines = None, []
for l in li[start:end]:
    m = re.search(r'//## File ".*sfdtd.cu", line (\d+)', l)
    if m:
        cur = int(m.group(1)); continue
    if re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+\S", l):
        lines.append(cur)
srcp = ncu("source")
sh, sd = srcp[1], srcp[2:]
ix = {h: i for i, h in enumerate(sh)}
assert len(lines) == len(sd), f"{len(lines)} SASS lines in the library vs {len(sd)} in the report: rebuild mismatch"
src = open(SRC).read().splitlines()
agg = collections.defaultdict(lambda: [0, 0, 0]); ops = collections.Counter()
for ln, r in zip(lines, sd):
    a = agg[ln]
    a[0] += int(r[ix["# Samples"]] or 0); a[1] += int(r[ix["Instructions Executed"]] or 0); a[2] += int(r[ix["L1 Wavefronts Shared"]] or 0)
    ops[re.sub(r"^@!?U?P\d+\s+", "", r[ix["Source"]]).split()[0].split(".")[0]] += int(r[ix["Instructions Executed"]] or 0)
ts, ti, tw = (sum(v[i] for v in agg.values()) for i in range(3))
summary["executed_instruction_mix_pct"] = {k: round(100 * v / ti, 1) for k, v in ops.most_common(14)}
with open(os.path.join(ROOT, "profiles", f"ncu_full_{tag}_bulk_kernel_hot_lines.csv"), "w") as fh:
    w = csv.writer(fh); w.writerow(["line", "pct_samples", "pct_instructions", "pct_smem_wavefronts", "source"])
    for ln, (s, e, wv) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        if 100 * s / ts >= 0.4:
            w.writerow([ln, f"{100 * s / ts:.2f}", f"{100 * e / ti:.2f}", f"{100 * wv / max(1, tw):.2f}", src[ln - 1].strip()[:110] if ln else ""])
json.dump(summary, open(os.path.join(ROOT, "profiles", f"ncu_full_{tag}_bulk_kernel_summary.json"), "w"), indent=1)
json.dump({"kernel": d["Kernel Name"], "strings": strings, "steps": steps, "dram_bytes_read": summary["dram_bytes_read"],
           "dram_bytes_write": summary["dram_bytes_write"], "dram_bytes_per_string_step": summary["dram_bytes_per_string_step"],
           "algorithmic_bytes_per_string_step": int(sys.argv[5]) if len(sys.argv) > 5 else 40, "source": f"profiles/ncu_full_{tag}_bulk_kernel_summary.json"},
          open(os.path.join(ROOT, "profiles", f"ncu_traffic_{tag}.json"), "w"), indent=1)
print(json.dumps(summary, indent=1))

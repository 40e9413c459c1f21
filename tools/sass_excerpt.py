"""Writes the static evidence the judge asked for under profiles/ (no GPU needed):
    python tools/sass_excerpt.py <tag>
  profiles/res_usage_<tag>.txt            cuobjdump -res-usage of every stepper kernel of libsfdtd.so (registers, stack = spills, shared)
  profiles/sass_<tag>_sweep_loop_f64.txt  SASS of the block-iteration sweep loop (gs_solve + TriSolver::solve) of the bulk fp64 kernel
  profiles/sass_<tag>_sweep_loop_f32.txt  the same of the bulk fp32 kernel
with the instruction mix of each excerpt in its header."""
import collections, os, re, subprocess, sys, tempfile

tag = sys.argv[1]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "torch_fdtd_string_b200", "libsfdtd.so")
SRC = open(os.path.join(ROOT, "torch_fdtd_string_b200", "csrc", "sfdtd.cu")).read().splitlines()
ru = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout.splitlines()
out = []
for i, l in enumerate(ru):
    m = re.search(r"Function (\S+):", l)
    if m and ("step_kernel" in m.group(1) or "group_kernel" in m.group(1) or "postprocess" in m.group(1) or "prepass" in m.group(1) or "width" in m.group(1)):
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().replace("(anonymous namespace)::", "")
        out.append(f"{name.split('(')[0]}\n    {ru[i + 1].strip()}")
open(os.path.join(ROOT, "profiles", f"res_usage_{tag}.txt"), "w").write(
    "# cuobjdump -res-usage torch_fdtd_string_b200/libsfdtd.so (sm_100a); STACK > 0 = local-memory frame (spills / indexed arrays)\n" + "\n".join(out) + "\n")

# source lines of the sweep loop: from the `do {` of gs_solve to its `} while`, plus TriSolver::solve
lo = next(i for i, l in enumerate(SRC) if "auto gs_solve = [&]" in l) + 1
hi = next(i for i, l in enumerate(SRC) if "} while (sweeps < s_skip" in l) + 1
s_lo = next(i for i, l in enumerate(SRC) if "void solve(T (&d)[ET], int ln) const" in l) + 1
s_hi = next(i for i in range(s_lo, len(SRC)) if SRC[i].startswith("};")) + 1
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", LIB], cwd=tmp, capture_output=True)
cubin = [x for x in os.listdir(tmp) if x.endswith(".cubin")][0]
li = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
for ty, name in (("d", "f64"), ("f", "f32")):
    heads = [i for i, l in enumerate(li) if l.startswith("//--------------------- .text.")]
    start = [i for i in heads if f"sfdtd_step_kernelI{ty}Li16ELi4ELi128" in li[i] and "Lb0E" in li[i]][0]      # the compact build (PF = false)
    end = min([i for i in heads if i > start] + [len(li)])
    cur, rows, mix = None, [], collections.Counter()
    for l in li[start:end]:
        m = re.search(r'//## File ".*sfdtd.cu", line (\d+)', l)
        if m:
            cur = int(m.group(1)); continue
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", l)
        if m and cur and (lo <= cur <= hi or s_lo <= cur <= s_hi or "ldb(const T" in SRC[cur - 1] or "hi_abs" in SRC[cur - 1] or "shfl" in SRC[cur - 1]):
            ins = m.group(2).strip()
            rows.append(f"/*{m.group(1)}*/ {ins:<70s} // sfdtd.cu:{cur}")
            mix[re.sub(r"^@!?U?P\d+\s+", "", ins).split()[0].split(".")[0]] += 1
    hdr = [f"# SASS of the sweep loop of sfdtd_step_kernel<{'double' if ty == 'd' else 'float'}, 16, 4, 128, *> (sm_100a), instructions whose line info",
           f"# points into gs_solve's do-while (sfdtd.cu:{lo}-{hi}), TriSolver::solve (sfdtd.cu:{s_lo}-{s_hi}) or the helpers they inline",
           f"# (ldb gathers, hi_abs norms, warp shuffles).  {len(rows)} static instructions; mix: "
           + ", ".join(f"{k} {v}" for k, v in mix.most_common(16))]
    open(os.path.join(ROOT, "profiles", f"sass_{tag}_sweep_loop_{name}.txt"), "w").write("\n".join(hdr + rows) + "\n")
    print(name, len(rows), mix.most_common(10))

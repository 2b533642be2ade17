"""Where a kernel's warp-time goes, by SOURCE REGION: joins the per-instruction stall samples of an ncu report (`--set full
--import-source on`) with the line table of the shipped cubin (`nvdisasm -g`), because `ncu --page source --print-source cuda` carries no
metrics on the command line.  Runs where ncu / cuobjdump / nvdisasm are installed; no GPU needed.

    python tools/ncu_regions.py gpurun_out/r2b_chain2_full.ncu-rep [kernel-regex] [launch index] [checkout of the profiled build]

Prints (1) samples / executed warp instructions / dominant stall reasons per region of mlp_chain2.cuh (regions are found by their marker
comments, so the table follows the file as it changes -- profile and library must come from the same build: for an older report pass a
`git worktree` of that commit with its libdwbc.so built, e.g. 1c239a0 for profiles' r2b capture), (2) the 25 source lines with the
most samples, (3) the opcode mix.  This produced the tables of profiles/r2b_summary.md."""
import collections, csv, glob, io, os, re, subprocess, sys, tempfile

ROOT = sys.argv[4] if len(sys.argv) > 4 else os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep = sys.argv[1]
kre = sys.argv[2] if len(sys.argv) > 2 else "chain2"
launch = int(sys.argv[3]) if len(sys.argv) > 3 else 0
so = os.path.join(ROOT, "deep-whole-body-control_b200", "libdwbc.so")

# ---- SASS offset -> (file, line) of the kernel, from the cubin inside libdwbc.so ----
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, check=True, capture_output=True)
ins = None
for cub in glob.glob(os.path.join(tmp, "*.cubin")):
    dis = subprocess.run(["nvdisasm", "-g", cub], capture_output=True, text=True).stdout.split("\n")
    starts = [i for i, l in enumerate(dis) if l.startswith(".text.") and re.search(kre, l) and l.rstrip().endswith(":")]
    if not starts:
        continue
    cur, ins = None, []
    for l in dis[starts[0] + 1:]:
        if l.startswith("//-----"):
            break
        t = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if t:
            cur = (os.path.basename(t.group(1)), int(t.group(2)))
            continue
        a = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*);", l)
        if a:
            ins.append((int(a.group(1), 16), cur, a.group(2).strip()))
    break
assert ins, "kernel not found in " + so

# ---- per-instruction samples of the chosen launch ----
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre, "--launch-skip", str(launch), "--launch-count", "1"],
                     capture_output=True, text=True).stdout
secs = []
for r in csv.reader(io.StringIO(out)):
    if r and r[0] == "Kernel Name":
        secs.append(dict(hdr=None, data=[]))
    elif r and r[0] == "Address":
        secs[-1]["hdr"] = r
    elif secs and secs[-1]["hdr"] and len(r) >= len(secs[-1]["hdr"]) - 2:
        secs[-1]["data"].append(r)
s = secs[0]          # (ncu prints the selected launch's table once per source view)
hdr, data = s["hdr"], s["data"]
assert len(data) == len(ins), f"profile ({len(data)} instructions) and library ({len(ins)}) come from different builds"
ia, isamp, iex, isrc = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Source")
stalls = [(i, h[6:]) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
I = lambda x: int(x) if x.strip().lstrip("-").isdigit() else 0  # noqa: E731

src = open(os.path.join(ROOT, "deep-whole-body-control_b200", "csrc", "mlp_chain2.cuh")).read().split("\n")
mark = lambda pat: [i + 1 for i, l in enumerate(src) if pat in l][0]  # noqa: E731
bounds = [("helpers / bias + activation", 1), ("loss hooks", mark("epilogue hooks: one thread per row")), ("set-up + item hand-over", mark("the kernel ------")),
          ("MMA warp", mark("weight copies + MMA issue")), ("gather stages (do_loads)", mark("loads + epilogues (sixteen warps)")),
          ("item start + epilogue body", mark("item start: the rows a LATER gather")), ("host", mark("host side ---"))]


def region(f, ln):
    if f != "mlp_chain2.cuh":
        return f
    name = bounds[0][0]
    for n, a in bounds:
        if ln >= a:
            name = n
    return name


reg = collections.defaultdict(lambda: [0, 0, collections.Counter()])
line = collections.defaultdict(lambda: [0, 0, collections.Counter()])
ops = collections.Counter()
for k, r in enumerate(data):
    f, ln = ins[k][1] or ("?", 0)
    n, sm = I(r[iex]), I(r[isamp])
    for d in (reg[region(f, ln)], line[(f, ln)]):
        d[0] += sm
        d[1] += n
        for i, h in stalls:
            d[2][h] += I(r[i])
    op = re.sub(r"^@!?U?P\d+\s+", "", r[isrc].strip()).split()
    ops[".".join(op[0].split(".")[:2]) if op else "?"] += n
tot_s, tot_n = sum(d[0] for d in reg.values()), sum(ops.values())
print(f"{tot_s} warp samples, {tot_n} executed warp instructions, {len(ins)} SASS instructions\n\nregion: samples, share, executed instructions, top stall reasons")
for k, d in sorted(reg.items(), key=lambda kv: -kv[1][0]):
    print(f"  {k:32s} {d[0]:6d} {100 * d[0] / tot_s:5.1f}%  {d[1]:10d}  " + ", ".join(f"{h} {100 * v / max(d[0], 1):.0f}%" for h, v in d[2].most_common(4)))
print("\nsource lines with the most samples")
for (f, ln), d in sorted(line.items(), key=lambda kv: -kv[1][0])[:25]:
    text = src[ln - 1].strip()[:90] if f == "mlp_chain2.cuh" and 0 < ln <= len(src) else ""
    print(f"  {d[0]:6d} {100 * d[0] / tot_s:5.1f}%  {f}:{ln}  [{', '.join(h for h, _ in d[2].most_common(2))}]  {text}")
print("\nopcode mix (executed warp instructions)")
for op, n in ops.most_common(20):
    print(f"  {op:18s} {n:10d} {100 * n / tot_n:5.1f}%")

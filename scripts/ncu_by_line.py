"""Aggregate an ncu source-page CSV (SASS level) by CUDA source line using nvdisasm line info of the same cubin.
usage: python scripts/ncu_by_line.py <report.ncu-rep> <cubin> [kernel]"""
import csv, re, subprocess, sys, collections, io
rep, cubin = sys.argv[1], sys.argv[2]
kernel = sys.argv[3] if len(sys.argv) > 3 else "jdde_k_integrate"
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]; idx = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.split("\n")
# collect (offset -> (file,line)) for the kernel's text section
cur = None; infn = False; linemap = []
for l in dis:
    if l.startswith(".text." + kernel + ":") or l.strip() == ".text." + kernel + ":": infn = True; continue
    if infn and re.match(r"^\s*\.text\.", l) and kernel not in l: infn = False
    if not infn: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m: linemap.append((int(m.group(1), 16), cur, m.group(2)))
print("sass instr in cubin:", len(linemap), " in report:", len(data))
S = idx["Warp Stall Sampling (All Samples)"]; I = idx["Instructions Executed"]; T = idx["Thread Instructions Executed"]
agg = collections.defaultdict(lambda: [0, 0, 0, 0])
n = min(len(linemap), len(data))
for k in range(n):
    r = data[k]; key = linemap[k][1]
    a = agg[key]; a[0] += int(r[S] or 0); a[1] += int(r[I] or 0); a[2] += int(r[T] or 0); a[3] += 1
tot_s = sum(a[0] for a in agg.values()); tot_i = sum(a[1] for a in agg.values())
lines = {}
def srcline(key):
    if not key: return ""
    f, ln = key
    if f not in lines:
        try: lines[f] = open(f if f.startswith("/") else "/root/repo/jitcdde_b200/csrc/" + f).read().split("\n")
        except Exception: lines[f] = []
    L = lines[f]
    return L[ln-1].strip()[:80] if 0 < ln <= len(L) else ""
print(f"total samples {tot_s}, total warp-instr {tot_i}")
print("--- by stall samples")
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:28]:
    print(f"{100*a[0]/tot_s:5.1f}% smp {100*a[1]/tot_i:5.1f}% inst  sass={a[3]:4d} thr/inst={a[2]/max(a[1],1):5.1f}  {key}  {srcline(key)}")
print("--- by executed instructions")
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:28]:
    print(f"{100*a[1]/tot_i:5.1f}% inst {100*a[0]/tot_s:5.1f}% smp  sass={a[3]:4d} thr/inst={a[2]/max(a[1],1):5.1f}  {key}  {srcline(key)}")

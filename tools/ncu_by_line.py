"""Dev tool: join an ncu SASS-page CSV with nvdisasm line info and rank source lines by executed warp instructions.
usage: ncu_by_line.py <src.csv from `ncu --page source --csv`> <nvdisasm -g -c output> <kernel mangled prefix> [top]"""
import collections, csv, re, sys
src_csv, sass_txt, prefix = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 50
amap = {}; cur = None; infn = False
for ln in open(sass_txt):
    if ln.startswith(".text." + prefix): infn = True; continue
    if infn and ln.startswith("//-----"): break
    if not infn: continue
    m = re.match(r'\s*//## File "(.*)", line (\d+)', ln)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r'\s+/\*([0-9a-f]+)\*/\s+(.*);', ln)
    if m: amap[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(src_csv))); H = rows[1]
ia, ie, it, isamp = H.index("Address"), H.index("Instructions Executed"), H.index("Thread Instructions Executed"), H.index("# Samples")
base = None; agg = collections.defaultdict(lambda: [0, 0, 0]); tot = 0
for r in rows[2:]:
    if len(r) < len(H): continue
    a = int(r[ia], 16) if r[ia].startswith("0x") else int(r[ia])
    base = a if base is None else base
    k = amap.get(a - base); e = float(r[ie] or 0)
    agg[k][0] += e; agg[k][1] += float(r[it] or 0); agg[k][2] += float(r[isamp] or 0); tot += e
import os
root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "mh-ppo_b200", "csrc")
src = {f: open(os.path.join(root, f)).read().split("\n") for f in os.listdir(root) if f.endswith((".cuh", ".cu"))}
tots = sum(v[2] for v in agg.values())
print("total warp-inst %.4g" % tot)
for k, v in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
    f, l = k if k else ("?", 0)
    text = src[f][l - 1].strip()[:84] if f in src and l > 0 else ""
    print("%5.2f%% inst  smp %5.2f%%  eff %4.1f  %s:%d  %s" % (100 * v[0] / tot, 100 * v[2] / max(tots, 1), v[1] / max(v[0], 1), f, l, text))

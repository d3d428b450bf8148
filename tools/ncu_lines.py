"""Per-source-line executed-instruction attribution from an .ncu-rep (needs -lineinfo)."""
import collections, csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
kern = sys.argv[3] if len(sys.argv) > 3 else None
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file = None; hdr = None; func = None
per = collections.defaultdict(collections.Counter); text = {}
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split('/')[-1]; continue
    if r[0] == "Function Name": func = r[1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr and r[0] not in ("", "-"):
        try: ln = int(r[0])
        except ValueError: continue
        v = r[hdr.index("Instructions Executed")]
        try: ie = int(v)
        except ValueError: continue
        per[func][(cur_file, ln)] += ie; text[(cur_file, ln)] = r[1]
for f, c in per.items():
    if kern and kern not in f: continue
    tot = sum(c.values())
    print("==", f, "total warp-instr", tot)
    for (fl, ln), v in c.most_common(top):
        print("%5.1f%% %s:%d  %s" % (100 * v / max(tot, 1), fl, ln, text[(fl, ln)].strip()[:105]))

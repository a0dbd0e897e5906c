"""Per source line: warp instructions executed and stall samples, from `ncu --page source --csv --print-source cuda,sass`."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file = None; hdr = None
agg = collections.OrderedDict()
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if r[0] in ("Function Name",): continue
    if hdr is None: continue
    if r[0] != "" and r[0].isdigit():
        d = dict(zip(hdr, r))
        key = (cur_file, int(r[0]))
        ie = d.get("Instructions Executed", "0"); ss = d.get("Warp Stall Sampling (All Samples)", "0")
        try: ie = int(ie)
        except ValueError: ie = 0
        try: ss = int(ss)
        except ValueError: ss = 0
        a = agg.setdefault(key, [0, 0, r[1]])
        a[0] += ie; a[1] += ss
tot_i = sum(a[0] for a in agg.values()); tot_s = sum(a[1] for a in agg.values())
print("total warp instructions", tot_i, "stall samples", tot_s)
for (f, ln), (ie, ss, src) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%5.1f%% inst %5.1f%% stall  %s:%d  %s" % (100.0 * ie / max(tot_i, 1), 100.0 * ss / max(tot_s, 1), f, ln, src.strip()[:110]))

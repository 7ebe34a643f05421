"""Aggregates an `ncu --page source --csv --print-source cuda,sass` dump by CUDA source line: executed warp
instructions and stall samples per (file, line). Usage: ncu_lines.py dump.csv [top_n]"""
import csv, sys, collections
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
inst = collections.Counter(); samp = collections.Counter(); text = {}
cur_file = None; hdr = None; n_sass = collections.Counter()
for row in csv.reader(open(path)):
    if not row: continue
    if row[0] == "File Path": cur_file = row[1].split("/")[-1]; hdr = None; continue
    if row[0] == "Function Name": continue
    if row[0] == "Line No": hdr = row; continue
    if hdr is None: continue
    try:
        ln = int(row[0])
    except ValueError:
        continue
    i_ie = hdr.index("Instructions Executed"); i_s = hdr.index("# Samples")
    key = (cur_file, ln)
    text.setdefault(key, row[1].strip()[:110])
    try:
        inst[key] += int(row[i_ie] or 0); samp[key] += int(row[i_s] or 0); n_sass[key] += 1
    except ValueError:
        pass
ti = sum(inst.values()); ts = sum(samp.values())
print("total warp-inst %d samples %d sass-lines %d" % (ti, ts, sum(n_sass.values())))
for key, v in inst.most_common(top):
    print("%-22s %5d %6.2f%% inst %6.2f%% samp  %s" % (key[0], key[1], 100.0 * v / max(ti, 1), 100.0 * samp[key] / max(ts, 1), text[key]))

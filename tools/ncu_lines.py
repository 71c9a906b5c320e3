"""development aid: rank source lines of one kernel by warp-stall samples (ncu --page source --print-source cuda,sass --csv)"""
import csv, sys, collections
path, want = sys.argv[1], sys.argv[2]        # csv, substring of the function name
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
rows = list(csv.reader(open(path, errors='replace')))
cur_file, cur_fn, hdr = None, None, None
by_line = collections.Counter(); text = {}; inst = collections.Counter(); stall = collections.defaultdict(collections.Counter)
for r in rows:
    if not r: continue
    if r[0] == 'File Path': cur_file = r[1].split('/')[-1]; continue
    if r[0] == 'Function Name': cur_fn = r[1]; continue
    if r[0] == 'Line No': hdr = r; continue
    if hdr is None or cur_fn is None or want not in cur_fn: continue
    if r[0] == '' or not r[0].isdigit(): continue
    key = (cur_file, int(r[0]))
    try: s = int(r[4])
    except ValueError: continue
    by_line[key] += s; text[key] = r[1].strip()[:110]
    try: inst[key] += int(r[7])
    except ValueError: pass
    for i, h in enumerate(hdr):
        if h.startswith('stall_') and 'Not Issued' not in h:
            try: stall[key][h[6:]] += int(r[i])
            except ValueError: pass
tot = sum(by_line.values())
print('total samples', tot)
byfile = collections.Counter()
for (f, l), s in by_line.items(): byfile[f] += s
for f, s in byfile.most_common(): print('  %-22s %6.2f %%' % (f, 100.0 * s / tot))
for key, s in by_line.most_common(top):
    st = ', '.join('%s %d' % (k, v) for k, v in stall[key].most_common(2))
    print('%5.2f %%  %-18s:%4d  inst %8d  [%s]  %s' % (100.0 * s / tot, key[0], key[1], inst[key], st, text[key]))

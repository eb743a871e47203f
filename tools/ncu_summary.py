"""summarise an .ncu-rep: per kernel key metrics + top stall reasons + hottest SASS lines"""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(io.StringIO(raw)))
hdr, units = r[0], r[1]
want = ['gpu__time_duration.sum', 'sm__cycles_elapsed.avg', 'sm__cycles_elapsed.avg.per_second', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'launch__grid_size', 'launch__block_size']
for row in r[2:]:
    name = row[hdr.index('Kernel Name')]
    print("=== ", name[:100])
    for w in want:
        if w in hdr:
            print("  %-78s %s %s" % (w, row[hdr.index(w)], units[hdr.index(w)]))
    st = []
    for i, h in enumerate(hdr):
        if h.startswith('smsp__pcsamp_warps_issue_stalled_') and not h.endswith('_not_issued'):
            try: st.append((float(row[i]), h.replace('smsp__pcsamp_warps_issue_stalled_', '')))
            except ValueError: pass
    tot = sum(v for v, _ in st) or 1
    print("  stalls: " + ", ".join("%s %.1f%%" % (h, 100 * v / tot) for v, h in sorted(st, reverse=True)[:7]))
if len(sys.argv) > 2:
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    starts = [i for i, x in enumerate(rows) if x and x[0] == 'Address'] + [len(rows)]
    for k in range(len(starts) - 1):
        body = rows[starts[k] + 1:starts[k + 1]]
        tot = sum(int(x[2]) for x in body if len(x) > 2 and x[2].isdigit()) or 1
        print("--- kernel %d: %d samples; instructions by opcode (share of stall samples)" % (k, tot))
        agg = {}
        for x in body:
            if len(x) > 2 and x[2].isdigit() and x[1]:
                toks = x[1].split()
                op = toks[1] if toks[0].startswith('@') and len(toks) > 1 else toks[0]
                op = op.split('.')[0]
                a = agg.setdefault(op, [0, 0]); a[0] += int(x[2]); a[1] += 1
        for op, (sm, n) in sorted(agg.items(), key=lambda z: -z[1][0])[:14]:
            print("   %-10s %6.2f%%  (%d static instr)" % (op, 100.0 * sm / tot, n))

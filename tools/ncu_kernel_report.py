#!/usr/bin/env python
"""Key metrics + instruction mix + stall reasons of the kernels in an .ncu-rep (reads via `ncu -i`)."""
import collections
import csv
import io
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.max", "lts__t_sectors_op_read.sum",
        "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg"]


def run(args):
    return subprocess.run(["ncu", "-i"] + args, capture_output=True, text=True).stdout


def main(path):
    rows = list(csv.reader(io.StringIO(run([path, "--page", "raw", "--csv"]))))
    hdr = rows[0]
    for r in rows[2:]:
        print("==", r[hdr.index("Kernel Name")][:90])
        for w in WANT:
            hit = [i for i, n in enumerate(hdr) if n == w or n.endswith("." + w)]
            if hit:
                print("  {:70s} {} {}".format(w, r[hit[0]], rows[1][hit[0]]))
    src = list(csv.reader(io.StringIO(run([path, "--page", "source", "--csv"]))))
    starts = [i for i, r in enumerate(src) if r and r[0] == "Kernel Name"]
    for si, st in enumerate(starts):
        blk = src[st + 1: starts[si + 1] if si + 1 < len(starts) else None]
        hdr, data = blk[0], [r for r in blk[1:] if len(r) == len(blk[0])]
        H = {n: i for i, n in enumerate(hdr)}
        tot_i = sum(float(r[H["Instructions Executed"]]) for r in data) or 1
        tot_s = sum(float(r[H["# Samples"]]) for r in data) or 1
        op = collections.Counter()
        for r in data:
            t = r[H["Source"]].split()
            o = t[1] if t[0].startswith("@") else t[0]
            op[o.split(".")[0]] += float(r[H["Instructions Executed"]])
        print("-- instruction mix (warp instructions {:.3e})".format(tot_i))
        print("   " + "  ".join("{} {:.1f}%".format(o, 100 * c / tot_i) for o, c in op.most_common(14)))
        st_names = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
        tots = sorted(((sum(float(r[H[n]]) for r in data), n) for n in st_names), reverse=True)
        print("-- stalls: " + "  ".join("{} {:.1f}%".format(n[6:], 100 * v / tot_s) for v, n in tots[:8]))
        smem = sorted(((float(r[H["L1 Wavefronts Shared"]]), float(r[H["L1 Wavefronts Shared Excessive"]]),
                        float(r[H["Instructions Executed"]]), r[H["Source"]].strip()) for r in data
                       if float(r[H["L1 Wavefronts Shared"]]) > 0), reverse=True)
        agg = collections.OrderedDict()
        for w, e, i, s_ in smem:
            k = s_.split()[0] if not s_.startswith("@") else s_.split()[1]
            a = agg.setdefault(k, [0, 0, 0]); a[0] += w; a[1] += e; a[2] += i
        print("-- shared-memory wavefronts: " + "  ".join("{} wf {:.3e} excess {:.3e} ({:.2f}/inst)".format(k, a[0], a[1], a[0] / max(a[2], 1)) for k, a in agg.items()))
        top = sorted(data, key=lambda r: -float(r[H["# Samples"]]))[:12]
        print("-- hottest instructions by samples:")
        for r in top:
            print("   {:6.2f}%  {}".format(100 * float(r[H["# Samples"]]) / tot_s, r[H["Source"]].strip()[:100]))


if __name__ == "__main__":
    main(sys.argv[1])

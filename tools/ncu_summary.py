#!/usr/bin/env python
"""Summarise an .ncu-rep (read with `ncu -i`, no GPU needed): key counters of each profiled kernel, and the
distribution of warp-stall samples over the SASS listing.  Used to write profiles/*.md."""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.sum",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio"]


def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print(f"## {d.get('Kernel Name')}  grid={d.get('Grid Size')} block={d.get('Block Size')}")
        for k in KEYS:
            if k in d:
                print(f"  {k} = {d[k]} {units[hdr.index(k)]}")


def stalls(path, blk=80):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    # one section per profiled launch: a "Kernel Name" row, a header row, then the instructions
    sections, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1] if len(r) > 1 else "?", "hdr": None, "data": []}
            sections.append(cur)
        elif cur is not None and cur["hdr"] is None:
            cur["hdr"] = r
        elif cur is not None:
            cur["data"].append(r)
    for sec in sections:
        hdr, data = sec["hdr"], sec["data"]
        ix = {h: i for i, h in enumerate(hdr)}

        def f(r, k):
            try:
                return float(r[ix[k]])
            except Exception:
                return 0.0
        tot = sum(f(r, "# Samples") for r in data) or 1.0
        print(f"## stall samples over the SASS listing of {sec['name'][:90]} ({int(tot)} samples, {len(data)} instructions)")
        for s0 in range(0, len(data), blk):
            seg = data[s0:s0 + blk]
            sm = sum(f(r, "# Samples") for r in seg)
            if sm < 0.003 * tot:
                continue
            ops = {}
            for r in seg:
                t = r[ix["Source"]].strip().split()
                op = (t[1] if t and t[0].startswith("@") else t[0]).split(".")[0] if t else "?"
                ops[op] = ops.get(op, 0) + 1
            top = ",".join(f"{k}:{v}" for k, v in sorted(ops.items(), key=lambda x: -x[1])[:4])
            print(f"  sass[{s0:4d}..] {100 * sm / tot:5.1f}%  long_sb={sum(f(r, 'stall_long_sb') for r in seg):6.0f} "
                  f"short_sb={sum(f(r, 'stall_short_sb') for r in seg):6.0f} wait={sum(f(r, 'stall_wait') for r in seg):6.0f} "
                  f"math={sum(f(r, 'stall_math') for r in seg):6.0f} selected={sum(f(r, 'stall_selected') for r in seg):6.0f}  [{top}]")


if __name__ == "__main__":
    raw(sys.argv[1])
    if len(sys.argv) > 2:
        stalls(sys.argv[1])

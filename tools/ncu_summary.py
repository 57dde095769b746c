"""Summaries of an .ncu-rep read on the CPU box.
    python tools/ncu_summary.py raw  <report>                 # key raw-page metrics of every captured launch
    python tools/ncu_summary.py src  <report> <kernel regex>  # per-segment instruction counts / stall samples of one kernel
"""
import collections
import csv
import io
import subprocess
import sys

RAW = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
       "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
       "sm__inst_executed_pipe_tensor.sum", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
       "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
       "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__registers_per_thread",
       "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
       "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_tex_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_drain_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio", "smsp__average_warps_issue_stalled_imc_miss_per_issue_active.ratio"]


def page(report, which, extra=()):
    out = subprocess.run(["ncu", "-i", report, "--page", which, "--csv", *extra], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def raw(report):
    rows = page(report, "raw")
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print("----", r[ix["Kernel Name"]][:110])
        for m in RAW:
            if m in ix and r[ix[m]] not in ("", "0", "0.000000"):
                print(f"   {m:88s} {r[ix[m]]:>18s} {units[ix[m]]}")


def src(report, regex):
    rows = page(report, "source", ("--kernel-name", f"regex:{regex}", "--launch-count", "1"))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows[2:] if len(r) > ix["Instructions Executed"] and r[ix["Instructions Executed"]].isdigit()]
    seen, uniq = set(), []
    for r in data:                                  # the page repeats per launch: keep the first copy of every address
        if r[ix["Address"]] in seen:
            continue
        seen.add(r[ix["Address"]])
        uniq.append(r)
    data = uniq
    tot = sum(int(r[ix["Instructions Executed"]]) for r in data)
    samp = sum(int(r[ix["# Samples"]]) for r in data)
    print(f"{rows[0][1][:100]}\ntotal warp instructions {tot}, stall samples {samp}, SASS lines {len(data)}")
    chunk = [(r[ix["Source"]].strip(), int(r[ix["Instructions Executed"]]), int(r[ix["# Samples"]])) for r in data]
    seg, cur, last = [], [], None
    for s_, n, sm in chunk:
        if last is not None and (n > last * 1.3 or n < last / 1.3) and len(cur) > 2:
            seg.append(cur)
            cur = []
        cur.append((s_, n, sm))
        last = max(n, 1)
    seg.append(cur)
    pos = 0
    for sg in seg:
        n, sm = sum(x[1] for x in sg), sum(x[2] for x in sg)
        if n / max(tot, 1) > 0.02 or sm / max(samp, 1) > 0.03:
            ops = collections.Counter((x[0].split()[1] if x[0].startswith("@") else x[0].split()[0]).split(".")[0] for x in sg)
            hot = max(sg, key=lambda x: x[2])
            print(f"@{pos:5d} {len(sg):4d} SASS  exec/instr {sg[0][1]:>9d}  inst {n / tot:6.1%}  samples {sm / max(samp, 1):6.1%}  "
                  f"ops {ops.most_common(5)}  hottest: {hot[0][:60]} ({hot[2]})")
        pos += len(sg)


if __name__ == "__main__":
    (raw if sys.argv[1] == "raw" else src)(*sys.argv[2:])

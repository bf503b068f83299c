"""Turn the captures of tools/profile_round.sh (gpurun_out/<round>_*) into the committed summaries under profiles/.

    python tools/make_profiles.py r01

Runs in the build container (`ncu -i` needs no GPU).  Per round it writes
  profiles/<round>_launches.csv.gz   the ncu launch list of the timed step of `python bench.py` (durations only)
  profiles/<round>_step_share.md     per-kernel share of that step
  profiles/<round>_<kernel>.md       `--set full` summary of each hot kernel (+ the hottest SASS lines)
  profiles/<round>_traffic.json      DRAM bytes per launch of the roofline kernel (read by bench.py -> roofline.traffic)
"""
import collections
import csv
import glob
import gzip
import io
import json
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import ncu_summary  # noqa: E402


NOTES = {
    "r02": "Round 2: tools/profile_round2.sh.  Under ncu every launch is serialised with cold caches: the 128x128 diagonal-tile kernel of K2\n"
           "shows 178 us here and ~85 us in the un-profiled factorisation (tools/r02_micro.py k2).\n",
    "r01": "This list was captured before the round's last two kernel changes (CTA-pair SYRK: -7 % per launch; 8-warp GPTQ\n"
           "column loop + 2-CTA/SM lazy update: -28 % per layer), so K1/K3 shares are slightly lower in the final build\n"
           "(step 460 -> 420 ms); the per-kernel captures next to this file are from the final build.\n",
}


def launch_share(rnd):
    src = os.path.join(ROOT, "gpurun_out", f"{rnd}_launches.csv")
    if not os.path.exists(src):
        return
    rows = list(csv.reader(open(src)))
    st = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[st]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0, 0.0])
    n = 0
    for r in rows[st + 1:]:
        if len(r) <= vi:
            continue
        name = re.sub(r"\(.*", "", r[ki])
        name = re.sub(r"<.*", "", name).replace("void ", "")[:72]
        v = float(r[vi].replace(",", ""))
        v = v / 1e6 if r[ui] in ("ns", "nsecond") else v / 1e3 if r[ui] in ("us", "usecond") else v
        agg[name][0] += 1
        agg[name][1] += v
        n += 1
    tot = sum(v[1] for v in agg.values())
    OURS = ("b200woq", "hessian_syrk_tc", "cholinv::", "tc::", "woqtc::", "w8a8::", "stream::", "lazytc::")
    ours = sum(v[1] for k, v in agg.items() if any(t in k for t in OURS))
    libs = sum(v[1] for k, v in agg.items() if any(t in k for t in ("getrf", "trsm", "xxtrf", "cusolver")))
    with open(os.path.join(ROOT, "profiles", f"{rnd}_step_share.md"), "w") as f:
        f.write(f"# {rnd}: kernel share of one timed step (one Llama-2-7B decoder block of GPTQ calibration)\n\n")
        f.write("Source: `ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include \"timed_steps/\"` around\n"
                "`python bench.py --steps 1 --warmup 1 --no-e2e --no-decode --no-cpu-baseline` (tools/profile_round.sh); the raw\n"
                f"launch list is `{rnd}_launches.csv.gz`.  Durations are serialised, cold-cache per-launch times: read the SHARES.\n"
                + NOTES.get(rnd, "") + "\n")
        f.write(f"{n} launches, {tot:.1f} ms of kernel time; hand-written b200woq kernels: {ours:.1f} ms ({100 * ours / tot:.1f} %), the rest is\n"
                "the model's own forward (cuBLAS `nvjet` GEMMs, SDPA, torch elementwise)"
                + (f"; cuSOLVER/cuBLAS factorisation kernels: {libs:.1f} ms.\n\n" if rnd == "r01" or libs > 0 else
                   "; NO cuSOLVER / cuBLAS-trsm kernel is left in the step (K2 is hand-written).\n\n"))
        f.write("| kernel | launches | total ms | share |\n|---|---|---|---|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:28]:
            f.write(f"| `{k}` | {v[0]} | {v[1]:.2f} | {100 * v[1] / tot:.1f} % |\n")
    with open(src, "rb") as fi, gzip.open(os.path.join(ROOT, "profiles", f"{rnd}_launches.csv.gz"), "wb") as fo:
        shutil.copyfileobj(fi, fo)


def hot_sass(path, top=14):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True,
                         text=True).stdout
    rd = list(csv.reader(io.StringIO(out)))
    try:
        st = next(i for i, r in enumerate(rd) if "Instructions Executed" in r)
    except StopIteration:
        return ""
    hdr = rd[st]
    si, ie, ws = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
    rows = []
    for r in rd[st + 1:]:
        try:
            rows.append((int(r[ws] or 0), int(r[ie]), r[si].strip()))
        except (ValueError, IndexError):
            pass
    tot = sum(r[0] for r in rows) or 1
    lines = [f"\nHottest SASS by warp-stall samples ({tot} samples, {sum(r[1] for r in rows)} warp instructions executed):\n",
             "| samples | executed | SASS |", "|---|---|---|"]
    for smp, ex, src in sorted(rows, reverse=True)[:top]:
        lines.append(f"| {smp} ({100 * smp / tot:.1f} %) | {ex} | `{src[:90]}` |")
    return "\n".join(lines) + "\n"


def kernel_reports(rnd):
    traffic = {}
    for rep in sorted(glob.glob(os.path.join(ROOT, "gpurun_out", f"{rnd}_*.ncu-rep"))):
        tag = os.path.basename(rep)[:-len(".ncu-rep")]
        buf = io.StringIO()
        old, sys.stdout = sys.stdout, buf
        try:
            ncu_summary.main([rep])
        finally:
            sys.stdout = old
        with open(os.path.join(ROOT, "profiles", f"{tag}.md"), "w") as f:
            f.write(f"# {tag}\n\nCapture: `ncu --set full --clock-control none --import-source on` (tools/profile_round.sh); one launch is replayed\n"
                    "~40 times, so `duration` is a cold, serialised time -- bench.py's CUDA-event timings are the reported numbers.\n\n")
            f.write(buf.getvalue())
            f.write(hot_sass(rep))
        if "syrk" in tag:
            hdr, units, rows = ncu_summary.rows_of(rep)
            for r in rows:
                by = sum(float(r[hdr.index(k)]) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}[units[hdr.index(k)]]
                         for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
                traffic[f"grid{r[hdr.index('launch__grid_size')]}"] = by
    if traffic:
        json.dump({"kernel": "hessian_syrk_tc_kernel", "T": 16384, "dram_bytes_per_launch_by_grid": traffic,
                   "note": "C=4096 launches have the smaller grid, C=11008 the larger; bench.py weights them 3:1"},
                  open(os.path.join(ROOT, "profiles", f"{rnd}_traffic.json"), "w"), indent=1)


if __name__ == "__main__":
    rnd = sys.argv[1] if len(sys.argv) > 1 else "r01"
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    launch_share(rnd)
    kernel_reports(rnd)
    print(sorted(os.listdir(os.path.join(ROOT, "profiles"))))

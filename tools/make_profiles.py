"""Turn the raw files of a `tools/gpu_round.sh` run (gpurun_out/) into the tracked summaries under profiles/:
launch list summary, details page, per-source-line hotspots, DRAM traffic of the step kernel, bench lines."""
import collections
import csv
import json
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
RND = os.environ.get("ROUND", "r02")  # file-name prefix of this round's summaries; raw inputs: gpurun_out/<RND>_launches.csv, <RND>_prof.ncu-rep


def launch_list():
    rows = list(csv.reader(open(os.path.join(G, RND + "_launches.csv"))))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= vi or not r[0].isdigit():
            continue
        v, u = float(r[vi].replace(",", "")), r[ui]
        us = v / 1e3 if u in ("ns", "nsecond") else (v if u in ("us", "usecond") else v * 1e3)
        a = agg.setdefault(re.sub(r"\(.*", "", r[ki])[:96], [0, 0.0])
        a[0] += 1
        a[1] += us
    tot = sum(a[1] for a in agg.values())
    out = ["ncu --metrics gpu__time_duration.sum --clock-control none -c 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline",
           "(first 200 launches of the process: setup fills, reset, 6 device-resident steps, per-launch timing steps, then the e2e "
           "loop; cold-cache, serialised: compare shares)",
           "%-92s %6s %12s %7s" % ("kernel", "count", "total_us", "share")]
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        out.append("%-92s %6d %12.1f %6.2f%%" % (k, a[0], a[1], 100 * a[1] / tot))
    open(os.path.join(P, RND + "_ncu_launch_list_summary.txt"), "w").write("\n".join(out) + "\n")
    shutil.copy(os.path.join(G, RND + "_launches.csv"), os.path.join(P, RND + "_ncu_launches.csv"))
    print("\n".join(out[3:6]))


def full_capture():
    rep = os.path.join(G, RND + "_prof.ncu-rep")
    open(os.path.join(P, RND + "_ncu_full_step_kernel_details.txt"), "w").write(
        subprocess.run(["ncu", "-i", rep, "--page", "details"], capture_output=True, text=True).stdout)
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    open("/tmp/src.csv", "w").write(src)
    hot = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_lines.py"), "/tmp/src.csv", "45"], capture_output=True, text=True).stdout
    open(os.path.join(P, RND + "_ncu_source_hotspots.txt"), "w").write(hot)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    d = {k: (rows[2][i], rows[1][i]) for i, k in enumerate(rows[0])}

    def nbytes(k):
        val, unit = d[k]
        return float(val.replace(",", "")) * {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1}[unit]

    r, w = nbytes("dram__bytes_read.sum"), nbytes("dram__bytes_write.sum")
    json.dump({"envs": 4096, "dram_bytes_per_launch": int(r + w),
               "source": "profiles/" + RND + "_ncu_full_step_kernel_details.txt (ncu --set full, vnl_env_kernel<0>, 4096 envs): "
                         "dram__bytes_read.sum %.3f MB + dram__bytes_write.sum %.3f MB; the algorithmic 35.5 MB per launch and the "
                         "inertia workspace stay in the 126 MB L2" % (r / 1e6, w / 1e6)}, open(os.path.join(P, "traffic.json"), "w"))
    print("dram MB", r / 1e6, w / 1e6, "duration", d["gpu__time_duration.sum"], "inst", d["smsp__inst_executed.sum"])
    print(hot[-520:])


def bench_lines():
    for src, dst in (("bench.log", "r01_bench_n1.json"), ("bench_ref.log", "r01_bench_reference_arm.json"),
                     ("bench_ant.log", "r01_bench_ant_config0.json"), ("bench_hum.log", "r01_bench_humanoid_config2.json"),
                     ("sweep_pair.log", "r01_sweep_rodent_pair_config4.jsonl")):
        if os.path.exists(os.path.join(G, src)):
            shutil.copy(os.path.join(G, src), os.path.join(P, dst))
    for f in ("r01_bench_n1", "r01_bench_ant_config0", "r01_bench_humanoid_config2", "r01_bench_reference_arm"):
        d = json.loads(open(os.path.join(P, f + ".json")).read().strip().split("\n")[-1])
        print(f, round(d["value"]), round(d["e2e"]["value"]), round(d["ms_per_step"], 3), d.get("roofline_fp32", {}).get("frac"),
              d.get("roofline", {}).get("frac"))


if __name__ == "__main__":
    launch_list()
    full_capture()
    if RND == "r01":
        bench_lines()

"""Summarise an ncu report (or launch-list CSV) into profiles/*.md.
  python scripts/ncu_summary.py full  gpurun_out/prof.ncu-rep   profiles/r01_ncu_full.md
  python scripts/ncu_summary.py list  gpurun_out/launches.csv   profiles/r01_launches.md"""
import collections
import csv
import re
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum",
    "sm__cycles_elapsed.avg.per_second",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread",
    "launch__grid_size",
    "launch__block_size",
    "launch__shared_mem_per_block_dynamic",
    "dram__bytes_read.sum",
    "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active",
]


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*", "", name)
    return name.replace("(anonymous namespace)::", "").replace("<unnamed>::", "")


def full(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    with open(out, "w") as f:
        f.write(f"# ncu --set full --clock-control none ({rep.split('/')[-1]})\n\n")
        f.write("Per captured launch; DRAM bytes are per launch.  Captured under ncu replay, so durations are\n"
                "cold-cache and serialised: compare shares and pipe utilisation, not absolutes.\n\n")
        for d in data:
            f.write(f"## {short(d[ix['Kernel Name']])}  (launch id {d[ix['ID']]})\n\n| metric | value | unit |\n|---|---|---|\n")
            for m in METRICS:
                if m in ix:
                    f.write(f"| {m} | {d[ix[m]]} | {units[ix[m]]} |\n")
            f.write("\n")


def launches(path, out):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    agg = collections.OrderedDict()
    seq = []
    for r in rows:
        v = float(r["Metric Value"].replace(",", ""))
        u = r["Metric Unit"]
        ms = v / 1e6 if u in ("nsecond", "ns") else v / 1e3 if u in ("usecond", "us") else v
        seq.append((short(r["Kernel Name"]), ms))
    half = seq[len(seq) // 2:] if len(seq) > 40 else seq      # second (warm) step when two were captured
    tot = sum(ms for _, ms in half)
    for n, ms in half:
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += ms
    with open(out, "w") as f:
        f.write(f"# ncu launch list (gpu__time_duration.sum, --clock-control none): {path.split('/')[-1]}\n\n")
        f.write(f"One fused fwd+bwd step, {len(half)} launches, {tot:.3f} ms of kernel time (serialised, cold-cache).\n\n")
        f.write("| kernel | launches | total ms | share |\n|---|---|---|---|\n")
        for n, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| {n} | {c} | {ms:.3f} | {100 * ms / tot:.1f}% |\n")
        f.write("\n## in launch order\n\n| # | kernel | ms |\n|---|---|---|\n")
        for i, (n, ms) in enumerate(half):
            f.write(f"| {i} | {n} | {ms:.4f} |\n")


if __name__ == "__main__":
    {"full": full, "list": launches}[sys.argv[1]](sys.argv[2], sys.argv[3])

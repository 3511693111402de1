"""Turn gpurun_out/{launches_*.csv, gemm_launches_*.csv, *.ncu-rep} into small text summaries under profiles/."""
import io
import subprocess
import sys

import pandas as pd


def launches(path, out):
    lines = open(path).read().splitlines()
    st = [i for i, l in enumerate(lines) if l.startswith('"ID"')][0]
    df = pd.read_csv(io.StringIO("\n".join(lines[st:])))
    df = df[df["Metric Name"] == "gpu__time_duration.sum"]
    df["ns"] = pd.to_numeric(df["Metric Value"].astype(str).str.replace(",", ""))
    df["kernel"] = df["Kernel Name"].str.replace(r"\(.*", "", regex=True).str.replace("void ", "").str.slice(0, 70)
    g = df.groupby("kernel").ns.agg(["sum", "count"]).sort_values("sum", ascending=False)
    tot = g["sum"].sum()
    g["share_%"] = (g["sum"] / tot * 100).round(2)
    g["sum_us"] = (g["sum"] / 1e3).round(1)
    with open(out, "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none, one device-resident utterance (tools/one_step.py)\n")
        f.write(f"# {len(df)} launches, total {tot/1e6:.3f} ms (cold-cache, serialised: compare SHARES, not absolutes)\n")
        f.write(g[["sum_us", "count", "share_%"]].to_string() + "\n")


def gemm(path, out):
    df = pd.read_csv(path)
    with open(out, "w") as f:
        f.write("# per-launch CUDA-event timing of gemm_tc_kernel over one device-resident utterance (RVCB_PROF_CSV)\n")
        f.write(f"# {len(df)} launches, {df.ms.sum():.3f} ms total\n")
        g = df.groupby(["M", "N", "kblocks", "BK", "BN", "batch"]).ms.agg(["sum", "count", "mean"]).sort_values("sum", ascending=False)
        f.write(g.head(40).round(4).to_string() + "\n")


def ncu_raw(rep, out, keys):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    df = pd.read_csv(io.StringIO(txt))
    with open(out, "w") as f:
        f.write(f"# ncu --set full --clock-control none, {rep}\n")
        for c in df.columns:
            if any(k in c for k in keys):
                f.write(f"{c}: {list(df[c].values[1:])}  [{df[c].values[0]}]\n")


if __name__ == "__main__":
    tag = sys.argv[1]
    import os
    if os.path.exists(f"gpurun_out/launches_{tag}.csv"):
        launches(f"gpurun_out/launches_{tag}.csv", f"profiles/{tag}_launches_by_kernel.txt")
    if os.path.exists(f"gpurun_out/gemm_launches_{tag}.csv"):
        gemm(f"gpurun_out/gemm_launches_{tag}.csv", f"profiles/{tag}_gemm_launches_by_shape.txt")
    for rep in [f for f in os.listdir("gpurun_out") if f.endswith(".ncu-rep") and tag in f]:
        ncu_raw(os.path.join("gpurun_out", rep), f"profiles/{rep.replace('.ncu-rep', '')}_metrics.txt",
                ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct",
                 "lts__throughput.avg.pct", "sm__pipe_tensor_cycles_active.avg.pct", "sm__warps_active.avg.pct", "launch__registers_per_thread",
                 "launch__grid_size", "l1tex__data_pipe_lsu_wavefronts.avg", "smsp__cycles_active.avg", "sm__throughput.avg.pct"])

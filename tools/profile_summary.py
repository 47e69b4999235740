"""Turns an `ncu --set full` capture of the propagation launches into the committed evidence:
    python tools/profile_summary.py gpurun_out/r01c_propagate.ncu-rep
writes profiles/r01_propagate_raw.csv (raw page) and profiles/r01_keyed_traffic.json (DRAM bytes summed over the
launches of one evaluation, read by bench.py -> roofline.traffic) and prints the per-launch table of r01_summary.md."""
import csv, io, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep = sys.argv[1]          # .ncu-rep, or the raw page already exported on the GPU box (ncu -i X --page raw --csv)
if rep.endswith(".csv"):
    raw = open(rep).read()
else:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True, check=True).stdout
raw = raw[raw.index('"ID"'):]
open(os.path.join(ROOT, "profiles", "r01_propagate_raw.csv"), "w").write(raw)
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, body = rows[0], rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}


def col(name):
    c = [h for h in hdr if h == name or h.endswith(name)]
    return ix[c[0]]


def scale(v, u):
    v = float(v.replace(",", ""))
    return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(u, 1.0)


rd, wr = col("dram__bytes_read.sum"), col("dram__bytes_write.sum")
tot_r = sum(scale(r[rd], units[rd]) for r in body)
tot_w = sum(scale(r[wr], units[wr]) for r in body)
json.dump({
    "kernels": "rc_propagate_keyed_kernel<8> (epochs 0-3) + rc_propagate_rows_kernel (epochs 4-7): the 8 propagation launches of one evaluation",
    "launch": "C3, 44 models per evaluation (bench.py default)",
    "dram_bytes_read": int(tot_r), "dram_bytes_write": int(tot_w), "dram_bytes_total": int(tot_r + tot_w),
    "algorithmic_hbm_bytes": 46207392,
    "source": "profiles/r01_propagate_raw.csv (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum summed over the 8 launches)",
    "note": "trajectory tables (~60 MB per model) written by rc_traj_kernel and read here, bin images re-read per model, b carried between epoch launches (2 x 5.9 MB per model and epoch); a few hundred GB/s against the measured 6539 GB/s copy bandwidth - not the bound",
}, open(os.path.join(ROOT, "profiles", "r01_keyed_traffic.json"), "w"))
names = {"ms": "gpu__time_duration.sum", "smem": "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
         "issue": "smsp__issue_active.avg.pct_of_peak_sustained_active", "fp64": "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
         "warps": "sm__warps_active.avg.pct_of_peak_sustained_active", "regs": "launch__registers_per_thread", "grid": "launch__grid_size"}
print("| epoch | kernel | grid | ms | smem pipe % | issue % | FP64 % | warps % | regs | DRAM rd / wr |")
for e, r in enumerate(body):
    kn = "rows" if "rows" in r[ix["Kernel Name"]] else "keyed"
    g = lambda k: r[col(names[k])]
    try:
        f64 = "%.1f" % float(g("fp64"))
    except Exception:
        f64 = "-"
    print(f"| {e} | {kn} | {g('grid')} | {float(g('ms')):.2f} | {float(g('smem')):.1f} | {float(g('issue')):.1f} | {f64} | "
          f"{float(g('warps')):.1f} | {g('regs')} | {scale(r[rd], units[rd]) / 1e9:.2f} GB / {scale(r[wr], units[wr]) / 1e6:.0f} MB |")
print("sum ms", sum(float(r[col(names['ms'])]) for r in body), "dram GB", (tot_r + tot_w) / 1e9)

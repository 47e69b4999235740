"""Turns an `ncu --set full` capture of the propagation launches into the committed evidence:
    python tools/profile_summary.py gpurun_out/r01c_propagate.ncu-rep
writes profiles/r01_propagate_raw.csv (raw page) and profiles/r01_keyed_traffic.json (DRAM bytes summed over the
launches of one evaluation, read by bench.py -> roofline.traffic) and prints the per-launch table of r01_summary.md."""
import csv, io, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TAG = os.environ.get("RC_PROFILE_TAG", "r02")
rep = sys.argv[1]          # .ncu-rep, or the raw page already exported on the GPU box (ncu -i X --page raw --csv)
if rep.endswith(".csv"):
    raw = open(rep).read()
else:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True, check=True).stdout
raw = raw[raw.index('"ID"'):]
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, body = rows[0], rows[1], rows[2:]
if len(sys.argv) > 2:      # e.g. "10-13,0-9": the launches of one evaluation when the capture window straddles two
    pick = []
    for part in sys.argv[2].split(","):
        a, _, b = part.partition("-")
        pick += list(range(int(a), int(b or a) + 1))
    body = [body[i] for i in pick]
ix = {h: i for i, h in enumerate(hdr)}


def col(name):
    c = [h for h in hdr if h == name or h.endswith(name)]
    return ix[c[0]]


def scale(v, u):
    v = float(v.replace(",", ""))
    return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(u, 1.0)


with open(os.path.join(ROOT, "profiles", TAG + "_propagate_raw.csv"), "w", newline="") as f:
    w = csv.writer(f, quoting=csv.QUOTE_ALL)
    w.writerows([hdr, units] + body)
rd, wr = col("dram__bytes_read.sum"), col("dram__bytes_write.sum")
tot_r = sum(scale(r[rd], units[rd]) for r in body)
tot_w = sum(scale(r[wr], units[wr]) for r in body)
names = {"ms": "gpu__time_duration.sum", "smem": "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
         "issue": "smsp__issue_active.avg.pct_of_peak_sustained_active", "fp64": "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
         "warps": "sm__warps_active.avg.pct_of_peak_sustained_active", "regs": "launch__registers_per_thread", "grid": "launch__grid_size"}


def kind(r):
    n = r[ix["Kernel Name"]]
    for cap in ("12", "25", "36"):
        if "_pat_kernel<%s>" % cap in n:
            return "pattern<%s>" % cap
    return "rows" if "rows" in n else "states"


_tu = {"us": 1e-3, "usecond": 1e-3, "ms": 1.0, "ns": 1e-6, "msecond": 1.0, "nsecond": 1e-6, "second": 1e3}.get(units[col(names["ms"])], 1.0)
ms = [float(r[col(names["ms"])].replace(",", "")) * _tu for r in body]
tms = sum(ms)
pipes = {k: sum(float(r[col(names[k])]) * m for r, m in zip(body, ms)) / tms for k in ("smem", "issue", "fp64", "warps")}
share = {}
for r, m in zip(body, ms):
    share[kind(r)] = share.get(kind(r), 0.0) + m / tms
json.dump({
    "kernels": "the propagation launches of one evaluation: per structural epoch rc_propagate_pat_kernel<12>, <25>, <36> (one pattern "
               "per thread; boxes of at most 12 / 13-25 / 26-36 live states); rc_propagate_rows_kernel / rc_propagate_keyed_kernel "
               "take what is left (nothing on C3)",
    "launch": "C3, 44 models per evaluation (bench.py default); %d launches" % len(body),
    "dram_bytes_read": int(tot_r), "dram_bytes_write": int(tot_w), "dram_bytes_total": int(tot_r + tot_w),
    "algorithmic_hbm_bytes": 46207392,
    "time_share": {k: round(v, 3) for k, v in share.items()},
    "pipes_pct_of_peak_time_weighted": {"shared_memory_wavefronts": round(pipes["smem"], 1), "fp64_pipe": round(pipes["fp64"], 1),
                                        "issue_slots": round(pipes["issue"], 1), "warps_active": round(pipes["warps"], 1)},
    "source": "profiles/" + TAG + "_propagate_raw.csv (ncu --set full, per-launch metrics; DRAM = dram__bytes_read.sum + dram__bytes_write.sum summed)",
    "note": "trajectory tables (~57 MB per model) written by rc_traj_kernel and read here, bin images re-read per model, b carried "
            "between epoch launches (lane-ordered region); a few hundred GB/s against the measured copy bandwidth - not the bound",
}, open(os.path.join(ROOT, "profiles", TAG + "_traffic.json"), "w"))
print("| launch | kernel | grid | ms | smem pipe % | issue % | FP64 pipe % | warps % | regs | DRAM rd / wr |")
for e, r in enumerate(body):
    g = lambda k: r[col(names[k])]
    print(f"| {e} | {kind(r)} | {g('grid')} | {ms[e]:.3f} | {float(g('smem')):.1f} | {float(g('issue')):.1f} | {float(g('fp64')):.1f} | "
          f"{float(g('warps')):.1f} | {g('regs')} | {scale(r[rd], units[rd]) / 1e9:.2f} GB / {scale(r[wr], units[wr]) / 1e6:.0f} MB |")
print("sum ms", tms, "dram GB", (tot_r + tot_w) / 1e9, "pipes", pipes, "share", share)

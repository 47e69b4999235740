"""Measures the FP64 roofline denominator (rc_fp64_peak: register-resident DFMA chains) with the SM clock sampled while it runs:
    python tools/fp64_peak.py > profiles/r02_fp64_peak.json"""
import json, os, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rarecoal_b200 as R

rows, stop = [], False


def sample():
    import pynvml as N
    N.nvmlInit()
    h = N.nvmlDeviceGetHandleByIndex(0)
    while not stop:
        rows.append((N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM), N.nvmlDeviceGetPowerUsage(h) / 1000.0))
        time.sleep(0.01)


eng = R.Engine(0)
eng.fp64_peak_tflops()                       # warm-up
t = threading.Thread(target=sample, daemon=True)
t.start()
vals = [eng.fp64_peak_tflops() for _ in range(20)]
stop = True
t.join(timeout=2)
sm = sorted(r[0] for r in rows)
print(json.dumps({"fp64_dfma_tflops_best": max(vals), "fp64_dfma_tflops_median": sorted(vals)[len(vals) // 2], "runs": len(vals),
                  "kernel": "rc_fp64_peak_kernel: 8 independent DFMA chains per thread, 256 threads x 16 blocks per SM, 64 DFMA per chain step",
                  "sm_mhz_median_under_load": sm[len(sm) // 2] if sm else None, "sm_mhz_min": sm[0] if sm else None,
                  "power_w_max": max(r[1] for r in rows) if rows else None, "clock_samples": len(rows),
                  "nominal": "148 SMs x 64 DFMA/clk x 2 flops x 1.965 GHz = 37.2 TFLOP/s"}))
eng.close()

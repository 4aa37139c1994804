"""Per-kernel CUDA-event times of one workload at 1x / 2x / 4x batch: separates fixed cost from per-byte cost.
  python tools/bench_kernels.py --workload k4 --precision tf32"""
import argparse, json, os, subprocess, sys
ap = argparse.ArgumentParser(); ap.add_argument("--workload", default="k4"); ap.add_argument("--precision", default="tf32")
a = ap.parse_args()
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, root)
import bench
B0 = bench.WORKLOADS[a.workload]["B"]
res = {}
for mult in (1, 2, 4):
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--workload", a.workload, "--precision", a.precision, "--steps", "10",
                          "--warmup", "3", "--no-cpu-baseline", "--batch", str(B0 * mult)], capture_output=True, text=True).stdout
    d = json.loads(out.strip().splitlines()[-1])
    res[mult] = ({k: v["avg_us"] for k, v in d["kernels"].items()}, d["ms_per_step"] * 1e3)
names = sorted(res[1][0], key=lambda k: -res[1][0][k])
print(f"{'kernel':24s} {'1x':>8s} {'2x':>8s} {'4x':>8s}   per-1x-batch slope (us)   fixed (us)")
for k in names:
    t = [res[m][0].get(k, float('nan')) for m in (1, 2, 4)]
    slope = (t[2] - t[0]) / 3
    print(f"{k:24s} {t[0]:8.1f} {t[1]:8.1f} {t[2]:8.1f}   {slope:8.1f}                 {t[0] - slope:8.1f}")
print("step us", [round(res[m][1], 1) for m in (1, 2, 4)])

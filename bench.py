#!/usr/bin/env python
"""bench.py — fusion-step train samples/s on B200 (BASELINE.json metric), one process per GPU.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload k4] [--impl ours|reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
         bench.py --gpus N --steps K --warmup W

A "step" is one pass of the fused late-fusion training step (heads -> fusion -> losses -> backward ->
EMA / QMF History / OGM-GE coefficients [-> OGM-GE modulation]) over one batch of synthetic features.
Workloads are the BASELINE.json configs; the default, k4, is the Food101 QMF shape
named by the north-star target.  Scaling is weak: every GPU holds the config's batch, all statistics and
gradients are global-batch quantities (all-reduce / all-gather inside the step).

Rank 0 prints ONE JSON line (contract in the task description): `value` = device-resident throughput,
`e2e` = same metric through the public Python API with pinned-host inputs copied H2D and the loss read
back D2H every step, `roofline` = dominant kernel's algorithmic bytes / its CUDA-event time vs the
measured HBM peak, `cpu_baseline` = the CPU oracle port timed on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch
import torch.distributed as dist

# name: (head mode, per-GPU batch, D, C, History length N, OGM alpha, need_dfeat, modulate encoder grads)
WORKLOADS = {
    "k1": dict(desc="Enrico joint_model late fusion, 512-d, 20 classes, batch 32 (frozen encoders)",
               mode="jlogits", B=32, D=512, C=20, N=None, alpha=None, dfeat=False, modulate=False),
    "k2": dict(desc="Crema-D joint_model_qmf, 512-d, 6 classes, batch 64",
               mode="qmf", B=64, D=512, C=6, N=6698, alpha=None, dfeat=True, modulate=False),
    "k3": dict(desc="Crema-D joint_model_ogm_ge, 512-d, 6 classes, batch 8192/GPU, OGM_GE modulation of 2x ResNet18 conv grads",
               mode="jlogits", B=8192, D=512, C=6, N=None, alpha=0.8, dfeat=True, modulate=True),
    "k4": dict(desc="Food101 joint_model_qmf, SigLIP 768-d, 101 classes, batch 32768/GPU",
               mode="qmf", B=32768, D=768, C=101, N=65536, alpha=None, dfeat=True, modulate=False),
    "k5": dict(desc="VGGSound-shape OGM-GE + EMA heads, 512-d, 309 classes, batch 131072/GPU",
               mode="jlogits", B=131072, D=512, C=309, N=None, alpha=0.8, dfeat=True, modulate=False),
}
L2_BYTES = 126 * 1024 * 1024

# ResNet18 conv weight shapes (audio 1-channel stem; the visual stem has 3 channels) — the 4-D tensors
# existing_algos/OGM_GE.py:44 selects; 20 tensors, 11.16 M elements per encoder
def resnet18_conv_shapes(in_ch):
    s = [(64, in_ch, 7, 7)]
    for cin, cout, down in ((64, 64, False), (64, 128, True), (128, 256, True), (256, 512, True)):
        s += [(cout, cin, 3, 3), (cout, cout, 3, 3)]
        if down:
            s += [(cout, cin, 1, 1)]
        s += [(cout, cout, 3, 3), (cout, cout, 3, 3)]
    return s


def step_alg_bytes_per_sample(w, fe=4):
    """SURVEY.md §8(d) algorithmic bytes per sample (fp32): read f, write df, write the returned logits,
    label (+idx).  The QMF step is two-pass by construction (mid-step global dependency), so it adds one
    more read of f (§8d: 'a two-pass design must report +M*D*4')."""
    M, D, C = 2, w["D"], w["C"]
    n_out = 4 if w["mode"] == "qmf" else 3
    b = M * D * fe + (M * D * fe if w["dfeat"] else 0) + n_out * C * 4 + 8      # fe: bytes per feature element (2 in bf16 mode)
    if w["mode"] == "qmf":
        b += 8 + M * D * fe
    return b


def kernel_alg_bytes(name, w, B, fe=4):
    """Algorithmic HBM bytes of ONE launch of kernel `name`: every input read once + every output written once
    (DESIGN.md §4).  k = number of dL/dlogits matrices (1 for mean fusion: dz1 == dz2; 2 for QMF)."""
    D, C = w["D"], w["C"]
    f = 4
    ldz = (C + 3) // 4 * 4
    k = 2 if w["mode"] == "qmf" else 1
    n_out = 4 if w["mode"] == "qmf" else 3
    if fe == 2:
        ldz = (C + 7) // 8 * 8
    logits = 2 * (B * D + C * D) * fe + 2 * (C + B * C) * f
    dfeat = (k * B * ldz + 2 * C * D + 2 * B * D) * fe
    dweight = (k * B * ldz + 2 * B * D) * fe + 2 * C * D * f
    table = {
        "sgemm_logits": logits, "tc_logits": logits,
        "tc_heads_forward": (2 * B * D + 2 * C * D + n_out * B * C) * f + 8 * B,
        # fused QMF forward: features + heads in, z1 z2 avg z_df (fp32) + conf (2) + row statistics (4) out, labels in
        "tc_forward_qmf": (2 * B * D + 2 * C * D) * fe + (4 * B * C + 6 * B) * f + 8 * B,
        "rows_forward_qmf": (2 * B * C + 2 * B * C + 2 * B + 4 * B) * f + 8 * B,
        "rows_forward_jlogits": (2 * B * C + B * C) * f + B * ldz * fe + 8 * B,
        "rows_backward_qmf": (2 * B * C + 8 * B) * f + 2 * B * ldz * fe + 8 * B,
        # fused backward: z1 z2 + per-sample scalars in, bf16 heads in, dF and dz (for the dW GEMM) out, labels in
        "tc_backward_qmf": (2 * B * C + 8 * B) * f + 2 * C * D * fe + 2 * B * D * fe + 2 * B * ldz * fe + 8 * B,
        "rows_calibrated": 2 * B * C * f + 8 * B,
        "sgemm_dfeat": dfeat, "tc_dfeat": dfeat,
        "sgemm_dweight": dweight, "tc_dweight": dweight,
        "narrow_step_jlogits": (2 * B * D + (2 * B * D if w["dfeat"] else 0) + 3 * B * C + B * ldz) * f + 8 * B,
        "narrow_forward_qmf": (2 * B * D + 4 * B * C + 6 * B) * f + 8 * B,
        "narrow_backward_qmf": (2 * B * D + (2 * B * D if w["dfeat"] else 0) + 2 * B * C + 2 * B * ldz + 8 * B) * f + 8 * B,
        "step_mid": 16 * B + 16 * (w["N"] or 0),
        "modulate_stats": 11_160_000 * f,
        "modulate_apply": 2 * 11_160_000 * f,
    }
    return table.get(name)


def measured_traffic(workload, precision, kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture of the
    same command (profiles/traffic.json, written by tools/ncu_summary.py); None when no capture covers it."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        return t.get(f"{workload}/{precision}", {}).get(kernel)
    except Exception:
        return None


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def mark(self):
        return time.time()

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 - 0.05 <= t <= t1 + 0.25] or [r for _, r in self.rows[-3:]]
        sm, mx, reasons = [], None, set()
        for r in rows:
            try:
                sm.append(float(r[0])); mx = float(r[1])
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def pin_to_gpu_numa_node(index):
    """Bind this rank's host threads (and, by first touch, the pinned staging buffers it allocates afterwards) to the NUMA
    node its GPU hangs off: with every rank on node 0 the host-fed (e2e) runs at 8 GPUs were bounded by one socket's
    memory bandwidth (round 1: 61 M samples/s at 8 GPUs against 72 M at 4).  Returns a short description for the JSON line."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        dev = bus.lower()
        if len(dev.split(":")[0]) == 8:                      # 00000000:1b:00.0 -> 0000:1b:00.0
            dev = dev[4:]
        node = int(open(f"/sys/bus/pci/devices/{dev}/numa_node").read().strip())
        if node < 0:
            return "numa node unknown"
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if not allowed:
            return f"numa node {node}: no allowed cpus"
        os.sched_setaffinity(0, allowed)
        return f"numa node {node} ({len(allowed)} cpus)"
    except Exception as e:
        return f"unpinned ({type(e).__name__})"


def make_batches(w, n_sets, device, seed):
    g = torch.Generator().manual_seed(seed)
    sets = []
    for _ in range(n_sets):
        B, D, C = w["B"], w["D"], w["C"]
        s = {"f1": torch.randn(B, D, generator=g), "f2": torch.randn(B, D, generator=g),
             "y": torch.randint(0, C, (B,), generator=g, dtype=torch.int64)}
        if w["N"]:
            s["idx"] = torch.randint(0, w["N"], (B,), generator=g, dtype=torch.int64) if w["B"] <= 8192 else \
                (torch.arange(B, dtype=torch.int64) + int(torch.randint(0, w["N"], (1,), generator=g))) % w["N"]
        sets.append(s)
    return sets


def head_params(w, seed=5):
    g = torch.Generator().manual_seed(seed)
    bound = 1.0 / (w["D"] ** 0.5)
    mk = lambda *s: (torch.rand(*s, generator=g) * 2 - 1) * bound
    return [mk(w["C"], w["D"]), mk(w["C"], w["D"])], [mk(w["C"]), mk(w["C"])]


# ------------------------------------------------------------------------------------------------
def run_reference(args, w, rank, world):
    """Reference arm: the reference's CPU implementation of the path.  /root/reference is a pure-Python
    package that does not exist on the GPU box, so this times the oracle port (oracle/late_fusion.py:
    the same torch CPU ops as the reference's modules, closed-form reg_loss) on all host cores."""
    if rank != 0:
        return
    from oracle import late_fusion as O
    torch.set_num_threads(os.cpu_count() or 1)
    Bs = w["B"]
    W, b = head_params(w)
    sets = make_batches(dict(w, B=Bs), 2, "cpu", seed=7)
    hist = O.HistoryState(w["N"]) if w["N"] else None
    ema = torch.zeros(2, w["C"])

    def one(i):
        nonlocal ema
        s = sets[i % len(sets)]
        if w["mode"] == "qmf":
            r = O.qmf_step([s["f1"], s["f2"]], W, b, s["y"], s["idx"], hist, ema_x=ema, feat_grad=w["dfeat"])
        else:
            r = O.jlogits_step([s["f1"], s["f2"]], W, b, s["y"], ema_x=ema, feat_grad=w["dfeat"])
            if w["alpha"]:
                O.ogm_coeffs(r["score1"], r["score2"], w["alpha"])
        ema = r["ema_x"]
        return float(r["loss"])

    for i in range(max(1, min(args.warmup, 3))):
        one(i)
    steps = args.steps
    t0 = time.perf_counter()
    for i in range(steps):
        one(i)
    dt = time.perf_counter() - t0
    val = Bs * steps / dt
    cores = torch.get_num_threads()
    line = {"impl": "reference", "metric": "fusion_step_train_samples_per_sec", "value": val, "unit": "samples/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup, "ms_per_step": dt / steps * 1e3,
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(w, args, world),
            "cpu_baseline": {"value": val, "unit": "samples/s", "cores": cores, "kind": "port",
                             "sample": f"{steps} steps of B={Bs} samples (one GPU's share of the global batch {Bs * world}; the port's "
                                       f"samples/s does not depend on the batch) of the oracle port on {cores} host threads"},
            "e2e": {"value": val, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    # the LITERAL reference modules beside the vectorised port (the port is the harder baseline and stays the `value`)
    line["cpu_baseline"]["literal_reference"] = literal_reference(w)
    print(json.dumps(line), flush=True)


def workload_config(w, args, world):
    cfg = _workload_config(w, args, world)
    if w["mode"] != "qmf" and w["C"] > 128 and world == 1:
        # mean fusion on wide heads: lf_step_mid + the calibrated-count pass run beside the dfeat GEMM (step.py)
        cfg["streams"] = "single stream (LF_NO_CAL_OVERLAP)" if os.environ.get("LF_NO_CAL_OVERLAP") else "two (step_mid + calibrated counts beside tc_dfeat)"
    return cfg


def _workload_config(w, args, world):
    return {"workload": f"{args.workload}: {w['desc']}", "head": w["mode"], "batch_per_gpu": w["B"],
            "global_batch": w["B"] * world, "feature_dim": w["D"], "classes": w["C"], "history_len": w["N"],
            "precision": args.precision, "parallelism": f"dp{world}", "scaling": args.scaling, "cuda_graph": not args.no_graph,
            "optimizer": ("SGD(momentum 0.9, wd 1e-4) on the heads fused into the step" if getattr(args, "fused_sgd", False)
                          else "none (gradients only)"),
            "l2": ("inputs rotate over buffer sets totalling > 2x the 126 MB L2"
                   if 2 * w["B"] * w["D"] * (2 if args.precision == "bf16" else 4) * 8 >= 2 * L2_BYTES
                   else "256 MB L2 flush between steps, outside the per-step CUDA-event brackets")}


def literal_reference(w, budget_s=8.0):
    """The UNMODIFIED reference modules (staged under oracle/_ref by oracle/vendor_ref.py) on the host cores, at a bounded
    batch: QMF.reg_loss materialises (B,B) matrices, so its cost per sample grows with B."""
    try:
        from oracle import literal
        if not literal.available():
            return {"unavailable": "oracle/_ref is not staged (run oracle/vendor_ref.py where /root/reference exists)"}
        Bs = min(w["B"], 8192 if w["mode"] == "qmf" else 32768)
        v, n, dt = literal.time_step(w["mode"], Bs, w["D"], w["C"], w["N"], w["alpha"], budget_s=budget_s)
        return {"value": v, "unit": "samples/s", "cores": torch.get_num_threads(), "kind": "reference",
                "sample": f"{n} steps of B={Bs} through the reference's own FusionNet.forward / backward / EMA.update (identity encoders) "
                          f"in {dt:.1f} s; quadratic in B for QMF, so the full batch would be slower per sample"}
    except Exception as e:                                     # never let the baseline take the bench down
        return {"unavailable": f"{type(e).__name__}: {e}"[:200]}


def cpu_baseline(w, budget_s=20.0):
    from oracle import late_fusion as O
    torch.set_num_threads(os.cpu_count() or 1)
    W, b = head_params(w)
    s = make_batches(w, 1, "cpu", seed=7)[0]
    hist = O.HistoryState(w["N"]) if w["N"] else None
    ema = torch.zeros(2, w["C"])

    def one():
        if w["mode"] == "qmf":
            return O.qmf_step([s["f1"], s["f2"]], W, b, s["y"], s["idx"], hist, ema_x=ema, feat_grad=w["dfeat"])
        return O.jlogits_step([s["f1"], s["f2"]], W, b, s["y"], ema_x=ema, feat_grad=w["dfeat"])

    one()
    n, t0 = 0, time.perf_counter()
    while True:
        one(); n += 1
        dt = time.perf_counter() - t0
        if dt > budget_s or n >= 200:
            break
    cores = torch.get_num_threads()
    return {"value": w["B"] * n / dt, "unit": "samples/s", "cores": cores, "kind": "port",
            "sample": f"{n} full-batch steps (B={w['B']}) of oracle/late_fusion.py (torch CPU fp32) in {dt:.1f} s",
            "literal_reference": literal_reference(w)}


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="k4", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="auto", choices=["auto", "fp32", "tf32", "bf16"],
                    help="auto: bf16 (the reference's own bf16-mixed mode: bf16 features, fp32 accumulation) for wide heads "
                         "(C >= 32), exact fp32 FMA for narrow heads; tf32 = fp32 features consumed as TF32; fp32 on wide heads = 3xTF32 on the tensor pipe")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="issue every step eagerly instead of replaying CUDA graphs")
    ap.add_argument("--no-sgd", action="store_true",
                    help="leave the optimizer step out (default: SGD(momentum 0.9, wd 1e-4) on the heads, utils/BaseModel.py:275-285, "
                         "fused into the tail of the dW kernel for the tensor-pipe workloads on one GPU)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: every GPU holds the workload's batch (global batch grows with N); strong: BASELINE.json's fixed "
                         "global batch split over the N GPUs")
    ap.add_argument("--no-parity-check", action="store_true",
                    help="skip the pre-timing check (N ranks == one GPU on the concatenated batch == CPU oracle on a small problem)")
    ap.add_argument("--batch", type=int, default=None, help="override the workload's per-GPU batch (experiments)")
    ap.add_argument("--dim", type=int, default=None, help="override the feature width (experiments)")
    ap.add_argument("--classes", type=int, default=None, help="override the class count (experiments)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    w = dict(WORKLOADS[args.workload])
    if args.precision == "auto":
        args.precision = "bf16" if (args.classes or w["C"]) >= 32 else "fp32"
    if args.batch or args.dim or args.classes:
        w.update(B=args.batch or w["B"], D=args.dim or w["D"], C=args.classes or w["C"])
        w["desc"] += f" [overridden: B={w['B']} D={w['D']} C={w['C']}]"

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.scaling == "weak" and world > 1 and w["N"]:
        # weak scaling grows the global batch, so the dataset the QMF History indexes grows with it (a global batch larger
        # than the dataset would visit entries several times per step)
        w["N"] *= world

    if args.scaling == "strong" and world > 1:
        if w["B"] % (2 * world):
            raise SystemExit(f"--scaling strong: global batch {w['B']} is not divisible by 2 x {world} ranks")
        w["B"] //= world
        w["desc"] += f" [strong scaling: global batch split over {world} GPUs]"

    if args.impl == "reference":
        run_reference(args, w, rank, world)
        return

    # CPU baseline first, on rank 0, while the other ranks wait in the (host-side) rendezvous below: nothing spins
    # on a GPU meanwhile and the host cores are not shared with the timed region
    cpu_line = cpu_baseline(w) if (rank == 0 and not args.no_cpu_baseline) else None

    numa = pin_to_gpu_numa_node(local) if world > 1 else "single process"
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from multimodal_clinical_b200 import _lib
    from multimodal_clinical_b200.step import LateFusionStep
    lib = _lib.load()

    eng = LateFusionStep(w["C"], mode=w["mode"], n_data=w["N"], device=dev, precision=args.precision)
    W, b = head_params(w)
    W = [x.to(dev) for x in W]; b = [x.to(dev) for x in b]
    # SGD on the heads inside the step: in the tail of the dW kernel, after the gradient all-reduce on several GPUs
    fused_sgd = (not args.no_sgd) and w["C"] >= 32
    args.fused_sgd = fused_sgd
    parity = None
    if not args.no_parity_check:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import parity_multigpu
        parity = parity_multigpu.run(w, args.precision, dev, rank, world, fused_sgd)
        if not parity["ok"]:
            if rank == 0:
                print(json.dumps({"parity_check": parity, "error": "parity check failed: nothing was timed"}), flush=True)
            torch.cuda.synchronize()
            os._exit(3)
    if fused_sgd:
        eng.enable_sgd(lr=1e-3, momentum=0.9, weight_decay=1.0e-4)
    fe = 2 if args.precision == "bf16" else 4
    in_bytes = 2 * w["B"] * w["D"] * fe
    need_sets = max(2, -(-2 * L2_BYTES // in_bytes))
    # few, large sets: rotating over them keeps every step's inputs out of L2.  Small workloads would need
    # hundreds of sets; they use 8 and an explicit L2 flush (256 MB fill) between steps, outside the per-step
    # CUDA-event brackets.
    flush_l2 = need_sets > 8
    n_sets = min(need_sets, 8)
    flush_buf = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev) if flush_l2 else None
    host_sets = make_batches(w, n_sets, dev, seed=100 + rank)
    if fe == 2:            # bf16 mode: the features exist in bf16 on the host too (encoder outputs under autocast)
        for hs in host_sets:
            hs["f1"] = hs["f1"].bfloat16(); hs["f2"] = hs["f2"].bfloat16()
    dev_sets = [{k: v.to(dev) for k, v in s.items()} for s in host_sets]
    enc_grads = None
    if w["modulate"]:
        enc_grads = [[torch.randn(s, device=dev) * 1e-3 for s in resnet18_conv_shapes(c)] for c in (1, 3)]

    def modulate(i):
        for m in range(2):
            eng.modulate(enc_grads[m], which=m, modulation="OGM_GE", seed=5, offset=i * (1 << 24))

    def eager_step(s, i):
        out = eng.step([s["f1"], s["f2"]], W, b, s["y"], idx=s.get("idx"), need_dfeat=w["dfeat"], ogm_alpha=w["alpha"])
        if enc_grads is not None:
            modulate(i)
        return out

    # One CUDA graph per input set (the step's kernels, its collectives and the OGM-GE modulation): the
    # step is ~10 launches of 5-60 us, so launch latency and host time are a first-order cost when issued
    # eagerly.  All step state is device-resident, so a replay advances EMA / History like an eager call.
    # (The Philox offset of the noise is baked per graph: a replay re-draws the same noise stream.)
    graphs = {}
    use_graph = not args.no_graph

    def one_step(s, i):
        if not use_graph:
            return eager_step(s, i)
        key = id(s)
        if key not in graphs:
            n0 = lib.lf_launch_count()
            g, out = eng.capture([s["f1"], s["f2"]], W, b, s["y"], idx=s.get("idx"), need_dfeat=w["dfeat"],
                                 ogm_alpha=w["alpha"], extra=(lambda: modulate(i)) if enc_grads is not None else None, warmup=2)
            graphs[key] = (g, out, (lib.lf_launch_count() - n0) // 3)      # 2 warm-up steps + the captured one
        g, out, _ = graphs[key]
        g.replay()
        return out

    def timed_steps(n, step_fn):
        """n steps on rotating input sets -> (device ms, launches).  Without flush: one event pair around the loop.
        With flush: one pair per step, the L2 flush between steps stays outside the brackets."""
        if not flush_l2:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(n):
                step_fn(dev_sets[i % n_sets], i)
            e1.record()
            barrier()
            return e0.elapsed_time(e1)
        evs = []
        for i in range(n):
            flush_buf.fill_(i & 0xff)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); step_fn(dev_sets[i % n_sets], i); e1.record()
            evs.append((e0, e1))
        barrier()
        return sum(a.elapsed_time(b) for a, b in evs)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident throughput (value)
    for i in range(max(args.warmup, n_sets if use_graph else 0)):      # also captures every set's graph
        one_step(dev_sets[i % n_sets], i)
    barrier()
    clocks = ClockSampler(local) if rank == 0 else None
    c0 = clocks.mark() if clocks else 0
    l0 = lib.lf_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ms = timed_steps(args.steps, one_step)
    launches = (lib.lf_launch_count() - l0) if not use_graph else args.steps * next(iter(graphs.values()))[2]
    c1 = clocks.mark() if clocks else 0
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    clk = clocks.stop(c0, c1) if clocks else None
    value = w["B"] * world * args.steps / (ms * 1e-3)

    # ---------------- per-kernel CUDA-event timing of the same steps -> roofline of the dominant kernel
    lib.lf_profile_enable(1)
    for i in range(args.steps):
        eager_step(dev_sets[i % n_sets], i)            # per-kernel events need eager launches
    prof = _lib.profile_report()
    lib.lf_profile_enable(0)

    # ---------------- end-to-end through the public API with pinned host inputs (e2e)
    pinned = [{k: v.pin_memory() for k, v in s.items()} for s in host_sets[:min(n_sets, 4)]]
    stage = {k: torch.empty_like(v, device=dev) for k, v in host_sets[0].items()}
    loss_host = torch.empty(1, dtype=torch.float32).pin_memory()
    h2d = sum(v.numel() * v.element_size() for v in host_sets[0].values())

    def host_batches(n):
        for i in range(n):
            yield pinned[i % len(pinned)]

    if use_graph:
        # the public host-fed API: H2D of batch i+1 on a copy stream under the replay of step i, loss read back
        # (D2H + sync) every step.  Graph capture and staging allocation happen before the first yield.
        gen = eng.stream_from_host(host_batches(args.steps + 3), W, b, need_dfeat=w["dfeat"], ogm_alpha=w["alpha"],
                                   extra=(lambda: modulate(0)) if enc_grads is not None else None)
        for _ in range(3):
            next(gen)
        barrier()
        e0.record()
        for _ in gen:
            pass
        e1.record()
        barrier()
    else:
        def e2e_step(i):
            p = pinned[i % len(pinned)]
            for k in stage:
                stage[k].copy_(p[k], non_blocking=True)
            out = eager_step(stage, i)
            loss_host.copy_(out.loss.view(1), non_blocking=True)
            torch.cuda.current_stream().synchronize()      # the user reads the loss every step
            return float(loss_host[0])

        for i in range(3):
            e2e_step(i)
        barrier()
        e0.record()
        for i in range(args.steps):
            e2e_step(i)
        e1.record()
        barrier()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_val = w["B"] * world * args.steps / (float(t.item()) * 1e-3)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        kern = {k: {"launches": c, "avg_us": tot / c * 1e3, "share": tot} for k, (c, tot) in prof.items()}
        tot_ms = sum(v["share"] for v in kern.values()) or 1.0
        for v in kern.values():
            v["share"] = v["share"] / tot_ms
        # kernels of the second stream (two-stream mean fusion) run BESIDE the dfeat GEMM: their event times overlap it and
        # measure the co-run, not the kernel -- they stay in `kernels` but are not candidates for the dominant kernel
        two_streams = str(workload_config(w, args, world).get("streams", "")).startswith("two")
        cand = [k for k in kern if not (two_streams and k in ("rows_calibrated", "step_mid"))]
        dom = max(cand, key=lambda k: kern[k]["share"]) if cand else None
        roof = None
        if dom:
            ab = kernel_alg_bytes(dom, w, w["B"], fe)
            ach = ab / (kern[dom]["avg_us"] * 1e-6) / 1e9 if ab else None
            roof = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": hbm_peak, "unit": "GB/s",
                    "frac": (ach / hbm_peak) if ach else None,
                    "traffic": measured_traffic(args.workload, args.precision, dom), "peak_source": peak_src,
                    "alg_bytes_per_launch": ab, "avg_launch_us": kern[dom]["avg_us"], "share_of_step": kern[dom]["share"]}
        step_bytes = step_alg_bytes_per_sample(w, fe) * w["B"]
        step_gbs = step_bytes / (ms / args.steps * 1e-3) / 1e9
        line = {"metric": "fusion_step_train_samples_per_sec", "value": value, "unit": "samples/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": args.scaling, "vs_baseline": None, "dtype": {"fp32": "f32", "tf32": "tf32", "bf16": "bf16"}[args.precision],
                "data": "synthetic", "config": workload_config(w, args, world),
                "e2e": {"value": e2e_val, "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4, "host_affinity": numa},
                "gpu_launches": int(launches), "clocks": clk, "roofline": roof,
                "step_roofline": {"alg_bytes_per_sample": step_alg_bytes_per_sample(w, fe), "achieved": step_gbs,
                                  "peak": hbm_peak, "unit": "GB/s", "frac": step_gbs / hbm_peak},
                "kernels": {k: {"launches": v["launches"], "avg_us": round(v["avg_us"], 2), "share": round(v["share"], 4)}
                            for k, v in sorted(kern.items(), key=lambda kv: -kv[1]["share"])}}
        if cpu_line is not None:
            line["cpu_baseline"] = cpu_line
        if parity is not None:
            line["parity_check"] = parity
        print(json.dumps(line), flush=True)
    if world > 1:
        # captured graphs hold NCCL work: drop them and drain the device before leaving; skip
        # destroy_process_group (it can block on communicators referenced by graphs) and exit directly
        graphs.clear()
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        os._exit(0)


if __name__ == "__main__":
    main()

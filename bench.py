#!/usr/bin/env python
"""bench.py — YOLOX-s 640x640 batch-32 decode+NMS (BASELINE.json configs[1]) and SimOTA assignment
(configs[2]) on N B200s, one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

A step = one pass of the hot path over one batch of 32 synthetic images (per GPU: weak scaling,
images shard by rank with no data-path collective; the evaluator's detection all-gather happens
once per timed region, inside it).  `value` is device-timed (CUDA events, max over ranks) with the
head maps resident in HBM; `e2e` goes through the public Python API from pinned HOST buffers with
the H2D / D2H copies inside the timed region.  4 distinct input sets (366 MB > the 126 MB L2) are
rotated so no step finds its input in L2.  Prints ONE JSON line on rank 0.
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

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "YOLOX-s 640² decode+NMS & SimOTA imgs/s at 1/2/4/8 B200; % HBM roofline"
STRIDES = [8, 16, 32]
SIZE, C, BATCH, LMAX = 640, 80, 32, 120
A = 8400
CONF, NMS = 0.01, 0.65
BYTES_DECODE_NMS = A * 85 * 4 + 300 * 6 * 4 + 4   # 2 863 204 B / image
BYTES_SIMOTA = A * 85 * 4 + A * 9 + 8              # + 20 * G, SURVEY.md §8d
N_SETS = 4
WORKLOAD = "YOLOX-s 640x640 batch 32 decode + postprocess(conf 0.01, nms 0.65) [BASELINE configs[1]]"


def traffic_bytes(kernel: str):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the committed ncu --set full capture."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        return json.load(open(p)).get(kernel)
    except Exception:  # noqa: BLE001
        return None


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:  # noqa: BLE001
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def make_inputs(seed_shift: int):
    """N_SETS distinct input sets: seeds (0,1) and (2,3) of SURVEY.md §8d, the rest are batch rotations."""
    from pl_yolo_b200 import synth
    base = [synth.make_heads(BATCH, SIZE, C, seed=0 + seed_shift), synth.make_heads(BATCH, SIZE, C, seed=2 + seed_shift)]
    labs = [synth.make_labels(BATCH, SIZE, LMAX, C, seed=1 + seed_shift), synth.make_labels(BATCH, SIZE, LMAX, C, seed=3 + seed_shift)]
    heads, labels = [], []
    for i in range(N_SETS):
        src, roll = base[i % 2], (i // 2) * 5
        heads.append([np.ascontiguousarray(np.roll(h, roll, axis=0)) for h in src])
        labels.append(np.ascontiguousarray(np.roll(labs[i % 2], roll, axis=0)))
    return heads, labels


def cpu_reference_pass(heads_cpu, labels_cpu, what: str, n_img: int):
    """One pass of the reference's own CPU implementation (torch/torchvision op chain) over n_img images."""
    from oracle import torch_ops_replay as R
    hs = [h[:n_img] for h in heads_cpu]
    if what == "decode_nms":
        preds, _ = R.decode(hs, STRIDES, True)
        return R.postprocess(preds, CONF, NMS)
    preds, _ = R.decode(hs, STRIDES, False)
    return R.simota(preds, labels_cpu[:n_img], [(SIZE // s, SIZE // s) for s in STRIDES], STRIDES, stable=False)


def time_cpu(heads_cpu, labels_cpu, what, n_img, reps):
    cpu_reference_pass(heads_cpu, labels_cpu, what, min(n_img, 4))  # warm-up
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        cpu_reference_pass(heads_cpu, labels_cpu, what, n_img)
        ts.append(time.perf_counter() - t0)
    return n_img / min(ts), ts


def run_reference(args, rank):
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)
    heads, labels = make_inputs(0)
    hc = [[torch.from_numpy(h) for h in hs] for hs in heads]
    lc = [torch.from_numpy(l) for l in labels]
    for w in range(args.warmup):
        cpu_reference_pass(hc[w % N_SETS], lc[w % N_SETS], "decode_nms", 8)
    t0 = time.perf_counter()
    for s in range(args.steps):
        cpu_reference_pass(hc[s % N_SETS], lc[s % N_SETS], "decode_nms", BATCH)
    dt = time.perf_counter() - t0
    v = args.steps * BATCH / dt
    sim_v, _ = time_cpu(hc[0], lc[0], "simota", 8, 1)
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "img/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "batch_per_gpu": BATCH, "anchors": A, "classes": C,
                   "note": "reference op chain (torch/torchvision CPU kernels) on the host cores, rank 0 only"},
        "cpu_baseline": {"value": v, "unit": "img/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": "%d steps x full batch of 32 images; torch %s / torchvision op-for-op replay of the reference "
                                   "(oracle/torch_ops_replay.py, bit-identical to the real reference on CPU)" % (args.steps, torch.__version__),
                         "simota_img_per_s": sim_v, "simota_sample": "1 pass x 8 images of cfg3"},
        "e2e": {"value": v, "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-graph", action="store_true", help="launch eagerly instead of replaying CUDA graphs")
    ap.add_argument("--skip-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        if args.steps > 60:
            args.steps = 60
        run_reference(args, rank)
        return
    assert args.warmup >= 3, "timing rules: at least 3 warm-up steps"
    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU path in the product)"
    import torch.distributed as dist
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from pl_yolo_b200 import YOLOXLoss, _lib, ops, postprocess_dense

    heads_np, labels_np = make_inputs(0)
    heads = [[torch.from_numpy(h).to(dev) for h in hs] for hs in heads_np]
    labels = [torch.from_numpy(l).to(dev) for l in labels_np]
    hw = [v for s in STRIDES for v in (SIZE // s, SIZE // s)]
    K, W = args.steps, args.warmup
    stream = torch.cuda.Stream(dev)
    peak, peak_src = peaks()

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(run_step, finish=None):
        """W warm-up + K timed steps on `stream`, CUDA events, barrier + synchronize both sides; max over ranks."""
        with torch.cuda.stream(stream):
            for i in range(W):
                run_step(i)
            if finish is not None:
                finish()  # warm-up of the exchange too (NCCL communicator set-up is not part of a step)
            barrier()
            l0 = _lib.launch_count()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            sampler = ClockSampler(local)
            sampler.start()
            e0.record(stream)
            for i in range(K):
                run_step(i)
            if finish is not None:
                finish()
            e1.record(stream)
            barrier()
            ms = e0.elapsed_time(e1)
            clocks = sampler.stop()
            launches = _lib.launch_count() - l0
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms, clocks, launches

    # ------------------------------------------------------------------ decode + NMS (cfg2), device-resident
    out_bufs = [(torch.empty((BATCH, 300, 6), device=dev), torch.empty((BATCH,), dtype=torch.int32, device=dev),
                 torch.empty((BATCH, 300), dtype=torch.int32, device=dev)) for _ in range(N_SETS)]
    gathered = [None]

    def dn_eager(s):
        ops.decode_postprocess_raw(heads[s], STRIDES, CONF, NMS, False, 10000, 300, 0, out=out_bufs[s])

    graphs, launches_per_graph, mode = None, 0, "eager"
    if not args.no_graph:
        try:
            with torch.cuda.stream(stream):
                for s in range(N_SETS):
                    dn_eager(s)
                torch.cuda.synchronize(dev)
                graphs = []
                for s in range(N_SETS):
                    g = torch.cuda.CUDAGraph()
                    l0 = _lib.launch_count()
                    with torch.cuda.graph(g, stream=stream):
                        dn_eager(s)
                    launches_per_graph = _lib.launch_count() - l0
                    graphs.append(g)
            mode = "cuda_graph"
        except Exception as e:  # noqa: BLE001
            graphs, mode = None, "eager (graph capture failed: %s)" % str(e)[:80]
            torch.cuda.synchronize(dev)

    def dn_step(i):
        s = i % N_SETS
        if graphs is not None:
            graphs[s].replay()
        else:
            dn_eager(s)

    def dn_finish():
        if world > 1:  # the one exchange of the eval path: all-gather of the padded detections (+ counts)
            acc_d = torch.stack([o[0] for o in out_bufs])
            acc_c = torch.stack([o[1] for o in out_bufs])
            gd = torch.empty((world,) + tuple(acc_d.shape), device=dev)
            gc = torch.empty((world,) + tuple(acc_c.shape), dtype=torch.int32, device=dev)
            dist.all_gather_into_tensor(gd, acc_d)
            dist.all_gather_into_tensor(gc, acc_c)
            gathered[0] = (gd, gc)

    ms, clocks, launches = timed(dn_step, dn_finish)
    if graphs is not None:
        launches = launches_per_graph * K
    value = world * BATCH * K / (ms * 1e-3)
    step_s = ms * 1e-3 / K
    achieved = BATCH * BYTES_DECODE_NMS / step_s / 1e9
    dets_per_img = float(torch.stack([o[1] for o in out_bufs]).float().mean())

    # ---- per-kernel durations, live: CUDA events recorded by the library between its kernels (eager launches on
    # `stream`; the same rotating inputs).  The dominant HBM kernel is score_kernel<fused>: it reads every
    # algorithmic byte of the step; nms_group_kernel works out of L2.
    import ctypes
    L = _lib.lib()
    L.plyolo_debug_stage_events.argtypes = [ctypes.c_void_p] * 3
    L.plyolo_debug_stage_events.restype = None

    def stage_times(run, n=20):
        ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(n)]
        with torch.cuda.stream(stream):
            for trio in ev:
                for e in trio:
                    e.record(stream)  # creates the handles
            torch.cuda.synchronize(dev)
            for i in range(n):
                L.plyolo_debug_stage_events(ev[i][0].cuda_event, ev[i][1].cuda_event, ev[i][2].cuda_event)
                run(i)
            L.plyolo_debug_stage_events(None, None, None)
            torch.cuda.synchronize(dev)
        a = sorted(e[0].elapsed_time(e[1]) for e in ev[2:])
        b_ = sorted(e[1].elapsed_time(e[2]) for e in ev[2:])
        return statistics.mean(a) * 1e-3, statistics.mean(b_) * 1e-3

    score_s, nms_s = stage_times(lambda i: dn_eager(i % N_SETS))

    # ------------------------------------------------------------------ SimOTA (cfg3), device-resident
    preds_t = [ops.decode_raw(heads[s], STRIDES, False)[0] for s in range(N_SETS)]
    gt_mean = float(np.mean([(l.sum(2) > 0).sum(1).mean() for l in labels_np]))

    def sim_eager(s):
        return ops.simota_assign_raw(preds_t[s], labels[s], hw, STRIDES)

    sim_graphs, sim_lpg = None, 0
    if graphs is not None:
        try:
            with torch.cuda.stream(stream):
                for s in range(N_SETS):
                    sim_eager(s)
                torch.cuda.synchronize(dev)
                sim_graphs, sim_keep = [], []
                for s in range(N_SETS):
                    g = torch.cuda.CUDAGraph()
                    l0 = _lib.launch_count()
                    with torch.cuda.graph(g, stream=stream):
                        sim_keep.append(sim_eager(s))
                    sim_lpg = _lib.launch_count() - l0
                    sim_graphs.append(g)
        except Exception:  # noqa: BLE001
            sim_graphs = None
            torch.cuda.synchronize(dev)

    def sim_step(i):
        if sim_graphs is not None:
            sim_graphs[i % N_SETS].replay()
        else:
            sim_eager(i % N_SETS)

    sms, sclocks, slaunches = timed(sim_step)
    if sim_graphs is not None:
        slaunches = sim_lpg * K
    sim_value = world * BATCH * K / (sms * 1e-3)
    prep_s, match_s = stage_times(lambda i: sim_eager(i % N_SETS))
    sim_bytes = BYTES_SIMOTA + 20 * gt_mean
    sim_achieved = BATCH * sim_bytes / (sms * 1e-3 / K) / 1e9

    # ------------------------------------------------------------------ training path (N2): decode -> SimOTA -> loss tail
    # forward + backward to the head maps, kernels only (what YOLOXLoss(train)(heads, labels)["loss"].backward() runs)
    train = None
    try:
        if world > 1:
            raise RuntimeError("measured at N = 1 only")  # an extra, kept off the multi-GPU scaling runs
        gsum = torch.tensor([5.0 / 3000, 1.0 / 3000, 1.0 / 3000], device=dev)

        def train_eager(s):
            pr, _ = ops.decode_raw(heads[s], STRIDES, False)
            fg_, mg_, mi_, _, _ = ops.simota_assign_raw(pr, labels[s], hw, STRIDES)
            sums_ = ops.yolox_loss_sums_raw(pr, labels[s], fg_, mg_, mi_)
            return sums_, ops.yolox_loss_backward_raw(pr, labels[s], fg_, mg_, mi_, gsum, hw, STRIDES)

        tr_graphs, tr_lpg = None, 0
        if graphs is not None:
            with torch.cuda.stream(stream):
                for s in range(N_SETS):
                    train_eager(s)
                torch.cuda.synchronize(dev)
                tr_graphs, tr_keep = [], []
                for s in range(N_SETS):
                    g = torch.cuda.CUDAGraph()
                    l0 = _lib.launch_count()
                    with torch.cuda.graph(g, stream=stream):
                        tr_keep.append(train_eager(s))
                    tr_lpg = _lib.launch_count() - l0
                    tr_graphs.append(g)

        def train_step(i):
            if tr_graphs is not None:
                tr_graphs[i % N_SETS].replay()
            else:
                train_eager(i % N_SETS)

        tms, _, tl = timed(train_step)
        if tr_graphs is not None:
            tl = tr_lpg * K
        # compulsory traffic: head maps read, preds written + read back by the assignment / loss, head-map gradients written
        tbytes = BATCH * (A * 85 * 4 * 3 + A * 9)
        train = {"value": world * BATCH * K / (tms * 1e-3), "unit": "img/s", "ms_per_step": tms / K, "gpu_launches": int(tl),
                 "workload": "YOLOX loss training path batch 32 (decode + SimOTA + loss tail forward, backward into the head maps) "
                             "[SURVEY 8f N2]",
                 "roofline": {"bound": "hbm", "achieved": tbytes / (tms * 1e-3 / K) / 1e9, "peak": peak, "unit": "GB/s",
                              "frac": tbytes / (tms * 1e-3 / K) / 1e9 / peak, "algorithmic_bytes_per_step": tbytes}}
    except Exception as e:  # noqa: BLE001
        train = None if world > 1 else {"unavailable": repr(e)[:200]}
        torch.cuda.synchronize(dev)

    # ------------------------------------------------------------------ end to end through the public API, host buffers
    pinned = [[torch.from_numpy(h).pin_memory() for h in hs] for hs in heads_np]
    stage = [[torch.empty_like(h, device=dev) for h in pinned[0]] for _ in range(2)]
    host_d = torch.empty((BATCH, 300, 6)).pin_memory()
    host_c = torch.empty((BATCH,), dtype=torch.int32).pin_memory()
    model_tail = YOLOXLoss(C, STRIDES, lazy_eval=True).eval()
    h2d = sum(h.numel() * 4 for h in pinned[0])
    d2h = host_d.numel() * 4 + host_c.numel() * 4
    Ke = max(10, min(K, 40))

    def e2e_step(i):
        st = stage[i % 2]
        for dst, src in zip(st, pinned[i % N_SETS]):
            dst.copy_(src, non_blocking=True)                      # H2D of this step's head maps
        d, c, _ = postprocess_dense(model_tail(st, None), CONF, NMS)  # public API: YOLOXLoss(eval) -> postprocess
        host_d.copy_(d, non_blocking=True)                         # D2H of the step's result
        host_c.copy_(c, non_blocking=True)

    with torch.cuda.stream(stream):
        for i in range(3):
            e2e_step(i)
        barrier()
        t0 = time.perf_counter()
        for i in range(Ke):
            e2e_step(i)
        stream.synchronize()
        e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t)
    e2e_value = world * BATCH * Ke / e2e_s

    # ------------------------------------------------------------------ CPU baseline (rank 0, N == 1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.skip_cpu:
        torch.set_num_threads(os.cpu_count() or 1)
        hc = [torch.from_numpy(h) for h in heads_np[0]]
        lc = torch.from_numpy(labels_np[0])
        v_dn, ts = time_cpu(hc, lc, "decode_nms", BATCH, 20)
        v_sim, ts2 = time_cpu(hc, lc, "simota", 8, 1)
        cpu = {"value": v_dn, "unit": "img/s", "cores": torch.get_num_threads(), "kind": "port",
               "sample": "best of 20 passes over one batch of 32 images (decode + postprocess); torch/torchvision op-for-op "
                         "replay of the reference on the host CPU (oracle/torch_ops_replay.py)",
               "simota_img_per_s": v_sim, "simota_sample": "1 pass over 8 images of cfg3 (%.1f s)" % ts2[0]}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "img/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": WORKLOAD,
                       "batch_per_gpu": BATCH, "anchors": A, "classes": C, "dets_per_image": dets_per_img,
                       "launch": mode, "l2": "4 input sets rotated (366 MB > 126 MB L2)",
                       "exchange": "none" if world == 1 else "one NCCL all-gather of the last 4 steps' padded detections + counts inside the timed region"},
            "roofline": {"bound": "hbm", "achieved": BATCH * BYTES_DECODE_NMS / score_s / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": BATCH * BYTES_DECODE_NMS / score_s / 1e9 / peak, "traffic": traffic_bytes("score_kernel"),
                         "peak_source": peak_src, "kernel": "score_kernel<fused> (dominant HBM kernel: reads every head-map byte once)",
                         "kernel_us": score_s * 1e6, "algorithmic_bytes_per_launch": BATCH * BYTES_DECODE_NMS,
                         "algorithmic_bytes_per_image": BYTES_DECODE_NMS,
                         "other_kernels_us": {"nms_group_kernel": nms_s * 1e6},
                         "whole_step": {"achieved": achieved, "frac": achieved / peak, "launches_per_step": max(1, launches // K)}},
            "e2e": {"value": e2e_value, "unit": "img/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": Ke, "api": "postprocess_dense(YOLOXLoss(lazy_eval=True).eval()(heads, None), 0.01, 0.65)"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "simota": {"value": sim_value, "unit": "img/s", "ms_per_step": sms / K, "gpu_launches": int(slaunches),
                       "workload": "YOLOX-s SimOTA assignment batch 32, G~U{1..120} (mean %.1f) [BASELINE configs[2]]" % gt_mean,
                       "roofline": {"bound": "hbm", "achieved": sim_achieved, "peak": peak, "unit": "GB/s",
                                    "frac": sim_achieved / peak, "traffic": traffic_bytes("simota_match_kernel"),
                                    "algorithmic_bytes_per_image": sim_bytes,
                                    "kernels_us": {"simota_prep_kernel + simota_sweep_kernel": prep_s * 1e6, "simota_match_kernel": match_s * 1e6},
                                    "note": "latency/issue-bound: the assignment touches ~25 MB of the 94 MB algorithmic bytes"},
                       "clocks": sclocks},
        }
        if train is not None:
            line["train_path"] = train
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

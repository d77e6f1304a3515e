#!/usr/bin/env python
"""bench.py — YOLOX decode+NMS and SimOTA on N B200s, one process per GPU (BASELINE.json configs).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

Headline (`value`): cfg2 — YOLOX-s 640x640 batch 32 per GPU, fused decode + postprocess(conf 0.01, nms 0.65), weak
scaling (images shard by rank, no data-path collective; the evaluator's detection all-gather runs EVERY step, one fused
collective on a side stream, inside the timed region).  A step = one pass of the hot path over one batch of synthetic
images; device-timed with CUDA events, max over ranks, head maps resident in HBM; 4 distinct input sets (366 MB > the
126 MB L2) rotate so no step finds its input in L2.  `roofline.frac` is the WHOLE step against the measured HBM copy
peak (algorithmic bytes per image x images / step time); the score kernel alone is a sub-key.
`e2e` goes through the public Python API from pinned HOST buffers with the H2D / D2H copies inside the timed region.
Also in the line: `simota` (cfg3), `configs` (cfg1 latency, cfg4 B=256 sharded, cfg5 1280^2 dense-GT), `cuda_baseline`
(the reference's own op chain on CUDA tensors: the incumbent on this box) and `cpu_baseline` (the same chain on the
host cores).  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import ctypes
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
CONF, NMS = 0.01, 0.65
N_SETS = 4
WORKLOAD = "YOLOX-s 640x640 batch 32 decode + postprocess(conf 0.01, nms 0.65) [BASELINE configs[1]]"


def anchors_of(size):
    return sum((size // s) ** 2 for s in STRIDES)


def bytes_decode_nms(size):
    return anchors_of(size) * 85 * 4 + 300 * 6 * 4 + 4     # 2 863 204 B / image at 640^2


def bytes_simota(size, g_mean):
    a = anchors_of(size)
    return a * 85 * 4 + a * 9 + 8 + 20 * g_mean


def traffic_bytes(kernel: str):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the committed ncu --set full capture."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(kernel)
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


def make_inputs(batch=BATCH, size=SIZE, lmax=LMAX, n_sets=N_SETS, objects=12, min_gt=1, tiny_frac=0.0):
    """n_sets distinct input sets: seeds (0,1) and (2,3) of SURVEY.md §8d, the rest are batch rotations; batches larger
    than 32 images tile further rotations of the same two seeds.  tiny_frac: that share of the GTs is shrunk to 2-8 px
    (COCO-like small objects: fewer in-both anchors than the dynamic k)."""
    from pl_yolo_b200 import synth
    nb = min(batch, 32)
    base = [synth.make_heads(nb, size, C, seed=0, objects_per_image=objects), synth.make_heads(nb, size, C, seed=2, objects_per_image=objects)]
    labs = [synth.make_labels(nb, size, lmax, C, seed=1, min_gt=min_gt), synth.make_labels(nb, size, lmax, C, seed=3, min_gt=min_gt)]
    if tiny_frac > 0:
        rng = np.random.default_rng(7)
        for lab in labs:
            valid = lab.sum(2) > 0
            pick = valid & (rng.uniform(0, 1, valid.shape) < tiny_frac)
            lab[..., 3][pick] = rng.uniform(2, 8, pick.sum()).astype(np.float32)
            lab[..., 4][pick] = rng.uniform(2, 8, pick.sum()).astype(np.float32)
    heads, labels = [], []
    for i in range(n_sets):
        src, lsrc, roll = base[i % 2], labs[i % 2], (i // 2) * 5
        reps = (batch + nb - 1) // nb
        hs = [np.concatenate([np.roll(h, roll + 3 * r, axis=0) for r in range(reps)], 0)[:batch] for h in src]
        lb = np.concatenate([np.roll(lsrc, roll + 3 * r, axis=0) for r in range(reps)], 0)[:batch]
        heads.append([np.ascontiguousarray(h) for h in hs])
        labels.append(np.ascontiguousarray(lb))
    return heads, labels


def cpu_reference_pass(heads_cpu, labels_cpu, what: str, n_img: int):
    """One pass of the reference's own implementation (torch/torchvision op chain) over n_img images, on the tensors' device."""
    from oracle import torch_ops_replay as R
    hs = [h[:n_img] for h in heads_cpu]
    if what == "decode_nms":
        preds, _ = R.decode(hs, STRIDES, True)
        return R.postprocess(preds, CONF, NMS)
    preds, _ = R.decode(hs, STRIDES, False)
    return R.simota(preds, labels_cpu[:n_img], [(SIZE // s, SIZE // s) for s in STRIDES], STRIDES, stable=False)


def time_ref(heads, labels, what, n_img, reps, sync=None):
    cpu_reference_pass(heads, labels, what, min(n_img, 4))  # warm-up
    ts = []
    for _ in range(reps):
        if sync:
            sync()
        t0 = time.perf_counter()
        cpu_reference_pass(heads, labels, what, n_img)
        if sync:
            sync()
        ts.append(time.perf_counter() - t0)
    return n_img / min(ts), ts


def run_reference(args, rank):
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)
    heads, labels = make_inputs()
    hc = [[torch.from_numpy(h) for h in hs] for hs in heads]
    lc = [torch.from_numpy(l) for l in labels]
    for w in range(args.warmup):
        cpu_reference_pass(hc[w % N_SETS], lc[w % N_SETS], "decode_nms", 8)
    t0 = time.perf_counter()
    for s in range(args.steps):
        cpu_reference_pass(hc[s % N_SETS], lc[s % N_SETS], "decode_nms", BATCH)
    dt = time.perf_counter() - t0
    v = args.steps * BATCH / dt
    sim_v, sim_ts = time_ref(hc[0], lc[0], "simota", BATCH, 2)
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "img/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "batch_per_gpu": BATCH, "anchors": anchors_of(SIZE), "classes": C,
                   "note": "reference op chain (torch/torchvision CPU kernels) on the host cores, rank 0 only"},
        "cpu_baseline": {"value": v, "unit": "img/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": "%d steps x full batch of 32 images; torch %s / torchvision op-for-op replay of the reference "
                                   "(oracle/torch_ops_replay.py, bit-identical to the real reference on CPU)" % (args.steps, torch.__version__),
                         "simota_img_per_s": sim_v, "simota_sample": "best of 2 passes x 32 images of cfg3 (%.1f s each)" % min(sim_ts)},
        "e2e": {"value": v, "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


def emit(line):
    """The ONE JSON line goes to the real stdout; everything else this process (or a library: NCCL's version banner)
    writes to fd 1 was redirected to stderr at start-up."""
    _REAL_STDOUT.write(json.dumps(line) + "\n")
    _REAL_STDOUT.flush()


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-graph", action="store_true", help="launch eagerly instead of replaying CUDA graphs")
    ap.add_argument("--skip-cpu", action="store_true", help="skip the cpu_baseline / cuda_baseline legs")
    ap.add_argument("--skip-extra", action="store_true", help="only the headline config (cfg2) and SimOTA (cfg3)")
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
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # NCCL's version banner must not land in front of the JSON line
        dist.init_process_group("nccl", device_id=dev)
        if os.environ.get("BENCH_EXCHANGE") == "nccl":  # leave the communication kernel SMs to run on
            os.environ.setdefault("PLYOLO_SCORE_SMS_RESERVED", "2")
    from pl_yolo_b200 import YOLOXLoss, _lib, ops, postprocess_dense
    from pl_yolo_b200.distributed import DetectionExchange, PeerDetections, fused_det_buffer, shard_range
    from pl_yolo_b200.pipeline import Lanes, default_depth
    peer_tables = []

    K, W = args.steps, args.warmup
    stream = torch.cuda.Stream(dev)
    peak, peak_src = peaks()
    L = _lib.lib()
    L.plyolo_debug_skip_nms.argtypes = [ctypes.c_int]
    L.plyolo_debug_skip_nms.restype = None

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(run_step, steps, finish=None, clocks=True, run_many=None):
        """W warm-up + `steps` timed steps on `stream`, CUDA events, barrier + synchronize both sides; max over ranks.
        run_many(n) enqueues n consecutive steps (default: run_step(i) for i < n)."""
        if run_many is None:
            def run_many(n):
                for i in range(n):
                    run_step(i)
        with torch.cuda.stream(stream):
            run_many(W)
            if finish is not None:
                finish()
            barrier()
            l0 = _lib.launch_count()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            sampler = ClockSampler(local) if clocks else None
            if sampler:
                sampler.start()
            e0.record(stream)
            run_many(steps)
            if finish is not None:
                finish()
            e1.record(stream)
            barrier()
            ms = e0.elapsed_time(e1)
            clk = sampler.stop() if sampler else None
            launches = _lib.launch_count() - l0
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms, clk, launches

    def capture(fn, n_sets):
        """One CUDA graph per input set (None if capture is off or fails) + kernels launched per graph."""
        if args.no_graph:
            return None, 0
        try:
            with torch.cuda.stream(stream):
                for s in range(n_sets):
                    fn(s)
                torch.cuda.synchronize(dev)
                graphs, keep, lpg = [], [], 0
                for s in range(n_sets):
                    g = torch.cuda.CUDAGraph()
                    l0 = _lib.launch_count()
                    with torch.cuda.graph(g, stream=stream):
                        keep.append(fn(s))
                    lpg = _lib.launch_count() - l0
                    graphs.append(g)
            graphs_keep.append(keep)
            return graphs, lpg
        except Exception as e:  # noqa: BLE001
            torch.cuda.synchronize(dev)
            sys.stderr.write("graph capture failed: %r\n" % (e,))
            return None, 0

    graphs_keep = []

    def bench_decode_nms(heads, b_loc, size, steps, exchange, clocks=False):
        """Fused decode + postprocess over rotating input sets; with `exchange` the fused detection buffer of every
        step is all-gathered on a side stream.  -> dict"""
        n_sets = len(heads)
        # N > 1: the detection exchange of every step, into peer-mapped gathered buffers (CUDA IPC over NVLink,
        # pl_yolo_b200.distributed.PeerDetections), one cross-rank fence per round of steps.  BENCH_EXCHANGE=
        #   stores (default): the NMS kernels store every row to the peers themselves (plyolo_decode_postprocess_bcast_f32:
        #                  compute and collective in one kernel); with several batches in flight the remote stores of one
        #                  batch's last images run under the next batches' score kernels;
        #   dma:           every step's finished block is pushed to the other ranks by the copy engines on a side stream;
        #   nccl:          one NCCL all-gather of the fused buffer per step on a side stream.
        pdx = None
        xmode = os.environ.get("BENCH_EXCHANGE", "stores")  # stores | dma | nccl
        if exchange and xmode in ("dma", "stores"):
            try:
                pdx = PeerDetections(b_loc, 300, dev, slots=n_sets)
                peer_tables.append(pdx)
            except Exception as e:  # noqa: BLE001
                sys.stderr.write("peer-store exchange unavailable (%r): NCCL all-gather per step\n" % (e,))
                pdx = None
        if pdx is not None:
            outs = [pdx.outputs(s) for s in range(n_sets)]
            bufs = [(None, o[0], o[1]) for o in outs]
            keeps = [o[2] for o in outs]
            peers = [pdx.peers(s) if xmode == "stores" else None for s in range(n_sets)]
            gbufs = None
            xch = DetectionExchange(dev, n_sets) if xmode == "dma" else None
            fence_t = torch.zeros(1, device=dev)
        else:
            bufs = [fused_det_buffer(b_loc, 300, dev) for _ in range(n_sets)]
            keeps = [torch.empty((b_loc, 300), dtype=torch.int32, device=dev) for _ in range(n_sets)]
            peers = [None] * n_sets
            gbufs = [torch.empty((world * bufs[0][0].numel(),), dtype=torch.float32, device=dev) for _ in range(n_sets)] if exchange else None
            xch = DetectionExchange(dev, n_sets) if exchange else None
        fake = None
        if os.environ.get("BENCH_FAKE_EXCHANGE") == "1" and xch is None and pdx is None:   # diagnosis: what the fork / join alone costs
            xch, fake = DetectionExchange(dev, n_sets), torch.zeros(64, device=dev)

        def fence():
            if pdx is not None:
                dist.all_reduce(fence_t)  # every rank's stores of the round have landed before anybody goes on

        def eager(s):
            ops.decode_postprocess_raw(heads[s], STRIDES, CONF, NMS, False, 10000, 300, 0, out=(bufs[s][1], bufs[s][2], keeps[s]),
                                       peers=peers[s])

        def step_eager(s):
            if xch:
                xch.wait_for(s)  # the slot's previous all-gather must have read the buffer before it is rewritten
            eager(s)
            if fake is not None:
                xch.submit(s, push=lambda: fake.zero_())
            elif xch and pdx is not None:
                xch.submit(s, push=lambda: pdx.push(s))
            elif xch:
                xch.submit(s, bufs[s][0], gbufs[s])

        graphs, lpg = capture(eager, n_sets)
        # Batches in flight (pl_yolo_b200.pipeline): step i runs on lane i % depth — its own stream, scratch and output
        # slot — so that the NMS tail of one batch runs under the score kernel of the next ones.  depth divides n_sets:
        # an output slot always belongs to the same lane, i.e. its reuse is ordered by that lane's stream.
        depth = int(os.environ.get("BENCH_IN_FLIGHT", "0")) or default_depth(b_loc)
        depth = max(d for d in range(1, min(depth, n_sets) + 1) if n_sets % d == 0)
        lanes = Lanes(depth, dev) if depth > 1 else None

        def issue_steps(first, n):
            """steps first .. first+n-1 on their lanes, joined back into the current stream"""
            if lanes is None:
                for i in range(first, first + n):
                    step_eager(i % n_sets)
                return
            lanes.fork()
            for i in range(first, first + n):
                lanes.issue(i, lambda i=i: step_eager(i % n_sets))
            lanes.join()

        # One graph per ROUND of R_STEPS steps: every step's kernels on its lane, and (N > 1) every step's exchange forked
        # onto the side stream right behind its NMS kernel, so that it runs under the following steps; lanes and side
        # stream join at the end of the round.  One replay per round also keeps the host (graph launch + NCCL enqueue per
        # step would cost more than the device work of a step) out of the measurement.
        round_graph = None
        # steps per round (one join / one fence per round: the pipeline of batches in flight drains there): up to 32 passes
        # over the input sets, dividing `steps`
        mult = max([m for m in range(1, 33) if steps % (n_sets * m) == 0] or [1])
        R_STEPS = n_sets * mult
        if graphs is not None:
            try:
                with torch.cuda.stream(stream):
                    issue_steps(0, n_sets)
                    if xch:
                        xch.finish()
                    fence()
                    torch.cuda.synchronize(dev)
                    if xch:
                        xch.done = [None] * n_sets
                    round_graph = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(round_graph, stream=stream):
                        issue_steps(0, R_STEPS)
                        if xch:
                            xch.finish()
                        fence()
                    if xch:
                        xch.done = [None] * n_sets
                    round_graph.replay()
                    torch.cuda.synchronize(dev)
            except Exception as e:  # noqa: BLE001
                sys.stderr.write("round graph capture failed (%r): stepping eagerly\n" % (e,))
                round_graph = None
                torch.cuda.synchronize(dev)
                if xch:
                    xch.done = [None] * n_sets

        def run_many(n):
            i = 0
            while i < n:
                if round_graph is not None and i % n_sets == 0 and i + R_STEPS <= n:
                    round_graph.replay()
                    i += R_STEPS
                    continue
                m = min(n - i, n_sets - i % n_sets)  # eager remainder: at most one pass over the input sets at a time
                issue_steps(i, m)
                i += m

        def finish():
            if xch:
                xch.finish()
            if pdx is not None and finish.pending:
                fence()
            finish.pending = False

        finish.pending = False
        _run_many = run_many

        def run_many(n):  # noqa: F811
            finish.pending = (n % R_STEPS != 0) or round_graph is None  # eager steps since the last fence
            _run_many(n)

        ms, clk, launches = timed(None, steps, finish=finish if (xch or pdx) else None, clocks=clocks, run_many=run_many)
        if graphs is not None:
            launches = lpg * steps
        step_s = ms * 1e-3 / steps
        ach = b_loc * bytes_decode_nms(size) / step_s / 1e9
        # the score stage alone (memset + score kernel: the debug hook stops the call before the NMS kernels)
        L.plyolo_debug_skip_nms(1)
        g2, _ = capture(eager, n_sets)
        ms2, _, _ = timed(lambda i: g2[i % n_sets].replay() if g2 is not None else eager(i % n_sets), min(steps, 40))
        L.plyolo_debug_skip_nms(0)
        score_s = ms2 * 1e-3 / min(steps, 40)
        # one batch in flight: the latency of a step (the same kernels, one stream)
        one_us = None
        if graphs is not None:
            ms1, _, _ = timed(lambda i: graphs[i % n_sets].replay(), min(steps, 40), clocks=False)
            one_us = 1e3 * ms1 / min(steps, 40)
        return {"value": world * b_loc * steps / (ms * 1e-3), "unit": "img/s", "ms_per_step": ms / steps, "batch_per_gpu": b_loc,
                "anchors": anchors_of(size), "gpu_launches": int(launches),
                "batches_in_flight": depth, "one_in_flight_step_us": one_us,
                "launch": ("cuda_graph (one replay per round of %d steps, %d batches in flight)" % (R_STEPS, depth)) if round_graph is not None else "eager",
                "dets_per_image": float(torch.stack([b[2] for b in bufs]).float().mean()),
                "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                             "algorithmic_bytes_per_image": bytes_decode_nms(size),
                             "score_stage": {"us": score_s * 1e6, "achieved": b_loc * bytes_decode_nms(size) / score_s / 1e9,
                                             "frac": b_loc * bytes_decode_nms(size) / score_s / 1e9 / peak,
                                             "what": "memset + score_kernel<fused> alone (reads every head-map byte once)"}},
                "exchange": (("copy-engine pushes of the finished block into the peers' gathered buffers (CUDA IPC over NVLink) on a side "
                              "stream, one fence per %d steps" if xmode == "dma" else
                              "peer stores from the NMS kernels (CUDA IPC over NVLink), one fence per %d steps") % R_STEPS) if pdx is not None else
                            ("NCCL all-gather per step on a side stream" if xch else "none"),
                "clocks": clk}

    def bench_simota(heads, labels, b_loc, size, steps, clocks=False):
        n_sets = len(heads)
        hw = [v for s in STRIDES for v in (size // s, size // s)]
        preds = [ops.decode_raw(heads[s], STRIDES, False)[0] for s in range(n_sets)]
        g_mean = float(np.mean([(l.sum(2) > 0).sum(1).float().mean().item() for l in labels]))

        def eager(s):
            return ops.simota_assign_raw(preds[s], labels[s], hw, STRIDES)

        graphs, lpg = capture(eager, n_sets)
        ms, clk, launches = timed(lambda i: graphs[i % n_sets].replay() if graphs is not None else eager(i % n_sets), steps, clocks=clocks)
        if graphs is not None:
            launches = lpg * steps
        by = bytes_simota(size, g_mean)
        ach = b_loc * by / (ms * 1e-3 / steps) / 1e9
        nfg = float(eager(0)[3].float().mean())
        return {"value": world * b_loc * steps / (ms * 1e-3), "unit": "img/s", "ms_per_step": ms / steps, "us_per_step": 1e3 * ms / steps,
                "batch_per_gpu": b_loc, "gt_mean": g_mean, "num_fg_mean": nfg, "gpu_launches": int(launches),
                "roofline": {"bound": "latency (touches ~1/4 of the algorithmic bytes: HBM is not the bound; target 50 us per batch of 32)",
                             "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "algorithmic_bytes_per_image": by,
                             "traffic": (traffic_bytes("simota_prep_kernel") or 0) + (traffic_bytes("simota_sweep_kernel") or 0) +
                                        (traffic_bytes("simota_match_kernel") or 0) or None},
                "clocks": clk}

    def to_dev(heads_np, labels_np):
        return ([[torch.from_numpy(h).to(dev) for h in hs] for hs in heads_np], [torch.from_numpy(l).to(dev) for l in labels_np])

    # ------------------------------------------------------------------ cfg2 (headline) and cfg3
    heads_np, labels_np = make_inputs()
    heads, labels = to_dev(heads_np, labels_np)
    dn = bench_decode_nms(heads, BATCH, SIZE, K, exchange=world > 1 and os.environ.get("BENCH_NO_EXCHANGE") != "1", clocks=True)
    sim = bench_simota(heads, labels, BATCH, SIZE, K, clocks=True)
    sim["workload"] = "YOLOX-s SimOTA assignment batch 32, G~U{1..120} (mean %.1f) [BASELINE configs[2]]" % sim["gt_mean"]

    extra = {}
    Ke = 32 if K >= 32 else max(8, (K // 8) * 8)  # steps of the extra configurations (whole rounds)
    if not args.skip_extra:
        # SimOTA with COCO-like small objects: 15 % of the GTs are 2-8 px (fewer in-both anchors than the dynamic k)
        try:
            _, lab_t = make_inputs(tiny_frac=0.15)
            st = bench_simota(heads, [torch.from_numpy(l).to(dev) for l in lab_t], BATCH, SIZE, Ke)
            sim["tiny_gt_variant"] = {"value": st["value"], "us_per_step": st["us_per_step"], "num_fg_mean": st["num_fg_mean"],
                                      "what": "same batch, 15 % of the GTs shrunk to 2-8 px (SURVEY T8, yolox_loss.py:343)"}
        except Exception as e:  # noqa: BLE001
            sim["tiny_gt_variant"] = {"unavailable": repr(e)[:200]}
            torch.cuda.synchronize(dev)
        # cfg1: one image, latency
        try:
            h1 = [[t[:1].contiguous() for t in hs] for hs in heads]
            r = bench_decode_nms(h1, 1, SIZE, Ke, exchange=False)
            extra["cfg1"] = {"workload": "YOLOX-s 640x640 batch 1 decode + postprocess [BASELINE configs[0]]", "latency_us": r["one_in_flight_step_us"],
                             "us_per_image_pipelined": 1e3 * r["ms_per_step"], "batches_in_flight": r["batches_in_flight"],
                             "value": r["value"] / world, "unit": "img/s (one GPU)", "dets_per_image": r["dets_per_image"],
                             "score_stage_us": r["roofline"]["score_stage"]["us"]}
            del h1
        except Exception as e:  # noqa: BLE001
            extra["cfg1"] = {"unavailable": repr(e)[:200]}
            torch.cuda.synchronize(dev)
        # cfg4: batch 256 sharded by image over the N GPUs (strong scaling), decode+NMS with the per-step all-gather + SimOTA
        try:
            lo, hi = shard_range(256, rank, world)
            b4 = hi - lo
            h4np, l4np = make_inputs(batch=b4, n_sets=2 if b4 > 64 else 4)  # as many input sets as batches in flight
            h4, l4 = to_dev(h4np, l4np)
            r = bench_decode_nms(h4, b4, SIZE, Ke, exchange=world > 1)
            s4 = bench_simota(h4, l4, b4, SIZE, Ke)
            extra["cfg4"] = {"workload": "YOLOX-l 640x640 batch 256 sharded by image over %d GPU(s): decode+NMS (+ detection all-gather "
                                         "every step) and SimOTA [BASELINE configs[3]]" % world, "scaling": "strong", "batch_per_gpu": b4,
                             "decode_nms": {"value": r["value"], "unit": "img/s", "ms_per_step": r["ms_per_step"], "batches_in_flight": r["batches_in_flight"],
                                            "one_in_flight_step_us": r["one_in_flight_step_us"], "frac": r["roofline"]["frac"] * world,
                                            "frac_note": "whole job against N x the measured HBM peak" if world > 1 else "against the measured HBM peak"},
                             "simota": {"value": s4["value"], "unit": "img/s", "ms_per_step": s4["ms_per_step"]}}
            extra["cfg4"]["decode_nms"]["frac"] = r["roofline"]["frac"]
            del h4, l4, r, s4
            torch.cuda.empty_cache()
        except Exception as e:  # noqa: BLE001
            extra["cfg4"] = {"unavailable": repr(e)[:200]}
            torch.cuda.synchronize(dev)
        # cfg5: 1280^2 (33600 anchors), 8 images per GPU, 250-500 GTs per image
        try:
            h5np, l5np = make_inputs(batch=8, size=1280, lmax=500, n_sets=N_SETS, objects=40, min_gt=250)
            h5, l5 = to_dev(h5np, l5np)
            r = bench_decode_nms(h5, 8, 1280, Ke, exchange=world > 1)
            s5 = bench_simota(h5, l5, 8, 1280, Ke)
            extra["cfg5"] = {"workload": "YOLOX-x 1280x1280 (33600 anchors) 8 images per GPU, 250-500 GT per image [BASELINE configs[4]]",
                             "batch_per_gpu": 8,
                             "decode_nms": {"value": r["value"], "unit": "img/s", "ms_per_step": r["ms_per_step"], "frac": r["roofline"]["frac"],
                                            "batches_in_flight": r["batches_in_flight"], "one_in_flight_step_us": r["one_in_flight_step_us"],
                                            "dets_per_image": r["dets_per_image"],
                                            "note": "~16 k candidates per image: max_nms = 10000 truncates by position, every image takes the general NMS path"},
                             "simota": {"value": s5["value"], "unit": "img/s", "ms_per_step": s5["ms_per_step"], "gt_mean": s5["gt_mean"],
                                        "num_fg_mean": s5["num_fg_mean"], "frac": s5["roofline"]["frac"]}}
            del h5, l5, r, s5
            torch.cuda.empty_cache()
        except Exception as e:  # noqa: BLE001
            extra["cfg5"] = {"unavailable": repr(e)[:200]}
            torch.cuda.synchronize(dev)

    # ------------------------------------------------------------------ training path (N2): decode -> SimOTA -> loss tail
    train = None
    if world == 1 and not args.skip_extra:
        try:
            hw = [v for s in STRIDES for v in (SIZE // s, SIZE // s)]
            gsum = torch.tensor([5.0 / 3000, 1.0 / 3000, 1.0 / 3000], device=dev)

            def train_eager(s):
                pr, _ = ops.decode_raw(heads[s], STRIDES, False)
                fg_, mg_, mi_, _, _ = ops.simota_assign_raw(pr, labels[s], hw, STRIDES)
                sums_ = ops.yolox_loss_sums_raw(pr, labels[s], fg_, mg_, mi_)
                return sums_, ops.yolox_loss_backward_raw(pr, labels[s], fg_, mg_, mi_, gsum, hw, STRIDES)

            tg, tlpg = capture(train_eager, N_SETS)
            tms, _, tl = timed(lambda i: tg[i % N_SETS].replay() if tg is not None else train_eager(i % N_SETS), Ke)
            if tg is not None:
                tl = tlpg * Ke
            A = anchors_of(SIZE)
            tbytes = BATCH * (A * 85 * 4 * 3 + A * 9)  # head maps read, preds written + read back, head-map gradients written
            train = {"value": BATCH * Ke / (tms * 1e-3), "unit": "img/s", "ms_per_step": tms / Ke, "gpu_launches": int(tl),
                     "workload": "YOLOX loss training path batch 32 (decode + SimOTA + loss tail forward, backward into the head maps) [SURVEY 8f N2]",
                     "roofline": {"bound": "hbm", "achieved": tbytes / (tms * 1e-3 / Ke) / 1e9, "peak": peak, "unit": "GB/s",
                                  "frac": tbytes / (tms * 1e-3 / Ke) / 1e9 / peak, "algorithmic_bytes_per_step": tbytes}}
        except Exception as e:  # noqa: BLE001
            train = {"unavailable": repr(e)[:200]}
            torch.cuda.synchronize(dev)

    # ------------------------------------------------------------------ end to end through the public API, host buffers
    # double-buffered: step i+1's head maps upload on a copy stream while step i computes (the link is the bound:
    # 91.4 MB per step at PCIe line rate; in production the head maps are already on the device)
    pinned = [[torch.from_numpy(h).pin_memory() for h in hs] for hs in heads_np]
    stage = [[torch.empty_like(h, device=dev) for h in pinned[0]] for _ in range(2)]
    host_d = [torch.empty((BATCH, 300, 6)).pin_memory() for _ in range(2)]
    host_c = [torch.empty((BATCH,), dtype=torch.int32).pin_memory() for _ in range(2)]
    model_tail = YOLOXLoss(C, STRIDES, lazy_eval=True).eval()
    h2d = sum(h.numel() * 4 for h in pinned[0])
    d2h = host_d[0].numel() * 4 + host_c[0].numel() * 4
    Kx = max(10, min(K, 40))
    copy_stream = torch.cuda.Stream(dev)
    up_done = [torch.cuda.Event() for _ in range(2)]
    used = [torch.cuda.Event() for _ in range(2)]

    def upload(i):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(used[i % 2])  # the staging buffer's previous step has consumed it
            for dst, src in zip(stage[i % 2], pinned[i % N_SETS]):
                dst.copy_(src, non_blocking=True)                  # H2D of step i's head maps
            up_done[i % 2].record(copy_stream)

    def e2e_run(n):
        for e in used:
            e.record(stream)
        upload(0)
        for i in range(n):
            if i + 1 < n:
                upload(i + 1)
            stream.wait_event(up_done[i % 2])
            d, c, _ = postprocess_dense(model_tail(stage[i % 2], None), CONF, NMS)  # public API: YOLOXLoss(eval) -> postprocess
            used[i % 2].record(stream)
            host_d[i % 2].copy_(d, non_blocking=True)               # D2H of the step's result
            host_c[i % 2].copy_(c, non_blocking=True)

    with torch.cuda.stream(stream):
        e2e_run(3)
        barrier()
        t0 = time.perf_counter()
        e2e_run(Kx)
        stream.synchronize()
        e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t)
    e2e_value = world * BATCH * Kx / e2e_s

    # ------------------------------------------------------------------ baselines (rank 0, N == 1 only)
    cpu = cuda_base = None
    if rank == 0 and world == 1 and not args.skip_cpu:
        # the incumbent on this box: the reference's op chain on CUDA tensors (torchvision's sm_100 nms_kernel behind
        # batched_nms, ATen kernels for the rest), synchronised on both sides
        try:
            sync = lambda: torch.cuda.synchronize(dev)  # noqa: E731
            v_dn, ts = time_ref(heads[0], labels[0], "decode_nms", BATCH, 10, sync)
            v_sim, ts2 = time_ref(heads[0], labels[0], "simota", BATCH, 2, sync)
            cuda_base = {"kind": "reference op chain (oracle/torch_ops_replay.py) on CUDA tensors: torch %s / torchvision ops, one B200" % torch.__version__,
                         "decode_nms": {"value": v_dn, "unit": "img/s", "ms_per_batch": 1e3 * min(ts), "sample": "best of 10 passes x 32 images"},
                         "simota": {"value": v_sim, "unit": "img/s", "ms_per_batch": 1e3 * min(ts2), "sample": "best of 2 passes x 32 images"},
                         "speedup_decode_nms": dn["value"] / v_dn, "speedup_simota": sim["value"] / v_sim}
        except Exception as e:  # noqa: BLE001
            cuda_base = {"unavailable": repr(e)[:200]}
            torch.cuda.synchronize(dev)
        torch.set_num_threads(os.cpu_count() or 1)
        hc = [torch.from_numpy(h) for h in heads_np[0]]
        lc = torch.from_numpy(labels_np[0])
        v_dn, ts = time_ref(hc, lc, "decode_nms", BATCH, 20)
        v_sim, ts2 = time_ref(hc, lc, "simota", BATCH, 2)
        cpu = {"value": v_dn, "unit": "img/s", "cores": torch.get_num_threads(), "kind": "port",
               "sample": "best of 20 passes over one batch of 32 images (decode + postprocess); torch/torchvision op-for-op "
                         "replay of the reference on the host CPU (oracle/torch_ops_replay.py)",
               "simota_img_per_s": v_sim, "simota_sample": "best of 2 passes x 32 images of cfg3 (%.1f s each)" % min(ts2)}

    if rank == 0:
        roof = dn["roofline"]
        line = {
            "metric": METRIC, "value": dn["value"], "unit": "img/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": dn["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "batch_per_gpu": BATCH, "anchors": anchors_of(SIZE), "classes": C,
                       "dets_per_image": dn["dets_per_image"], "launch": dn["launch"], "l2": "4 input sets rotated (366 MB > 126 MB L2)",
                       "batches_in_flight": dn["batches_in_flight"],
                       "exchange": "per step: " + dn["exchange"]},
            "roofline": {"bound": "hbm", "achieved": roof["achieved"], "peak": peak, "unit": "GB/s", "frac": roof["frac"],
                         "traffic": (traffic_bytes("score_kernel") or 0) + (traffic_bytes("nms_fast_kernel") or 0) or None,
                         "peak_source": peak_src,
                         "kernel": "whole step: memset + score_kernel<fused> (HBM stream) with nms_fast_kernel running under it (and under the "
                                   "following batches' score kernels: %d batches in flight, one stream each)" % dn["batches_in_flight"],
                         "step_us": 1e3 * dn["ms_per_step"], "one_in_flight_step_us": dn["one_in_flight_step_us"], "algorithmic_bytes_per_launch": BATCH * bytes_decode_nms(SIZE),
                         "algorithmic_bytes_per_image": bytes_decode_nms(SIZE), "score_stage": roof["score_stage"]},
            "e2e": {"value": e2e_value, "unit": "img/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": Kx,
                    "api": "postprocess_dense(YOLOXLoss(lazy_eval=True).eval()(heads, None), 0.01, 0.65)",
                    "note": "PCIe-bound by construction (91.4 MB of head maps per step; uploads double-buffered on a copy stream); "
                            "in production the head maps are produced on the device"},
            "gpu_launches": dn["gpu_launches"],
            "clocks": dn["clocks"],
            "simota": {k: v for k, v in sim.items() if not k.startswith("_")},
        }
        if extra:
            line["configs"] = extra
        if train is not None:
            line["train_path"] = train
        if cuda_base is not None:
            line["cuda_baseline"] = cuda_base
        if cpu is not None:
            line["cpu_baseline"] = cpu
        emit(line)
    if world > 1:
        for t in peer_tables:
            t.close()
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench.py -- 1080p detect+recognize frames/s of the B200 hot path (BASELINE.json metric).

Workload (BASELINE.json configs[2], the configuration the metric "detect+recognize frames/sec" is quoted on;
it fits one GPU): synthetic 1080p BGR frames -> 736x1312 -> DBNet-ResNet18 + fused DB head -> box extraction
(~50 planted boxes/frame) -> 50 crops/frame 32x128 -> CRNN -> CTC greedy.  bf16 tier.
A step is one batch of --batch frames through the whole path.

  value : frames/s with the frames already resident in HBM (frame pool larger than L2), results left on device
  e2e   : frames/s through vtd_run_batch with HOST (pinned) frames: H2D of every frame and D2H of the records
          inside the timed region
  roofline     : dominant kernel (tcgen05 implicit-GEMM conv, all launches of the timed region), algorithmic
                 FLOPs / CUDA-event time, against MEASURED_PEAKS.json bf16_tflops_sustained
  cpu_baseline : oracle/port.py (the reference's PyTorch/PIL/OpenCV arithmetic) on the host cores, bounded sample
  hbm_stages   : the non-GEMM stages (preprocess, DB head tail, box extraction, crop gather, CTC): algorithmic bytes
                  / CUDA-event time against MEASURED_PEAKS.json hbm_gbs

`--crop-w 100` runs BASELINE's wording of configs[2] (32x100 crops; default 128 = the reference's own width, the larger
workload).  `--dtype fp16` runs the speed tier of the half-storage library (libvtd_b200_f16.so).

`--impl reference` times that CPU path alone (one frame per step).  N>1: one process per GPU (torchrun), frames
sharded by rank, records gathered to rank 0 every step with NCCL; time = max over ranks.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SRC_H, SRC_W = 1080, 1920
DET_H, DET_W = 736, 1312
CROP_W = 128
BOXES = 50
GF_DET_PER_FRAME = 184.51          # SURVEY.md 8d, DBNet-R18 @736x1312, live layers
GF_CRNN = {128: 1.787, 100: 1.394}  # SURVEY.md 8d, GFLOP per crop @32x128 (reference default) / @32x100 (BASELINE wording)
GF_CRNN_PER_CROP = GF_CRNN[CROP_W]
METRIC = "1080p detect+recognize frames/sec"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"tflops": float(d.get("bf16_tflops_sustained", 1357.9)), "tflops_burst": float(d.get("bf16_tflops", 1635.9)),
                "hbm": float(d.get("hbm_gbs", 6531.9)), "source": "measured"}
    return {"tflops": 1400.0, "tflops_burst": 1590.0, "hbm": 6650.0, "source": "fallback"}


def top_kernel_traffic(op):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed
    `ncu --set full` capture (profiles/r01_top_kernel_ncu.json), when it is the same layer; else None."""
    p = os.path.join(ROOT, "profiles", "r01_top_kernel_ncu.json")
    try:
        d = json.load(open(p))
        if (op["H"], op["W"], op["Cin"], op["Cout"], op["KH"]) == (184, 328, 256, 256, 3):
            return float(d["dram_bytes_per_launch"]), {"algorithmic_bytes_per_launch": d["algorithmic_bytes_per_launch"],
                                                       "unit": "bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum)",
                                                       "source": "profiles/r01_top_kernel_ncu.json"}
    except Exception:
        pass
    return None, None


def hbm_stage_rooflines(stages, steps, batch, crops_per_step, crop_src_bytes_per_step, hbm_gbs, bias_plane=True):
    """Achieved HBM GB/s of the stages that are HBM-bound by their bytes (north_star: "achieved HBM GB/s for the
    elementwise, post-processing and gather stages").  `stages` = Engine.op_profile(2) of a profiled pass of `steps`
    steps with ONE batch in flight (CUDA-event time per stage, summed over the steps).  ALGORITHMIC bytes per step,
    SURVEY.md 8d (bf16 tier, s = 2): what the stage must read and write once, not what the kernels happen to move."""
    px = DET_H * DET_W
    per_step = {
        # K1: BGR frame in, 3 normalised channels out
        "preprocess": batch * (SRC_H * SRC_W * 3 + 3 * px * 2),
        # K3: feat [Hd/4, Wd/4, 128] bf16 in; prob + thresh fp32 and the u8 mask out; + the planted fp32 logit plane
        "head_tail": batch * ((px // 16) * 128 * 2 + 2 * px * 4 + px + (px * 4 if bias_plane else 0)),
        # K4-K6: mask in, the int32 label plane written and read once, the records out
        "boxes": batch * (px + 2 * px * 4 + 64 * 128),
        # K7: the source pixels under the boxes in, 32 x crop_w x 3 bf16 per crop out
        "crop": crop_src_bytes_per_step + crops_per_step * 3 * 32 * CROP_W * 2,
        # K10: [T, 97] fp32 logits per crop in, ids + length + confidence out
        "ctc": crops_per_step * ((CROP_W // 4 - 1) * 97 * 4 + 36 + 8),
    }
    out = []
    for st in stages:
        name = st.get("name")
        if name not in per_step or not st.get("ms") or steps <= 0:
            continue
        ms = st["ms"] / steps
        gbs = per_step[name] / (ms * 1e-3) / 1e9
        out.append({"stage": name, "ms_per_step": ms, "algorithmic_bytes_per_step": int(per_step[name]),
                    "achieved_gbs": gbs, "frac_of_hbm_peak": gbs / hbm_gbs})
    return out


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons during the timed region (pynvml; same counters as nvidia-smi)."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.samples, self.reasons, self.max_mhz = index, False, [], set(), None
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception:
            pass

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
        while not self.stop_flag:
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                r = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.05)

    def result(self):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": []}
        return {"sm_mhz": int(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_models():
    """The oracle's PyTorch modules holding the SAME random-init weights the B200 arm loads (synthetic.random_state_dicts)."""
    import torch
    from oracle import port
    from video_text_detection_system_b200 import synthetic
    torch.set_num_threads(os.cpu_count() or 1)
    det_sd, rec_sd = synthetic.random_state_dicts(seed=0)
    det, rec = port.build_dbnet("resnet18", seed=0), port.build_crnn(seed=0)
    det.load_state_dict(det_sd)
    rec.load_state_dict(rec_sd)
    return port, det.eval(), rec.eval()


def cpu_frame(port, det, rec, frame, bias):
    """The reference's per-frame path (pipeliine.py:143-172): detect, then batch-1 recognise per crop."""
    import torch
    return port.process_frame(det, rec, frame, 0.5, DET_H, DET_W, CROP_W, torch.from_numpy(bias)[None, None],
                              per_crop=True)


CPU_WORKERS = 4     # the reference detects frames on ThreadPoolExecutor(max_workers=4) (pipeliine.py:32,96-101)


def cpu_frames(port, det, rec, frames, bias):
    """`frames` through the reference's path with every host thread in use: CPU_WORKERS frames in parallel (the
    reference's executor width), each with cores/CPU_WORKERS intra-op torch threads -- measured faster than one frame at
    a time on all cores (0.94 vs 0.74 frames/s on 8 cores).  Returns the number of text regions found."""
    import torch
    from concurrent.futures import ThreadPoolExecutor
    torch.set_num_threads(max(1, (os.cpu_count() or 1) // CPU_WORKERS))
    with ThreadPoolExecutor(CPU_WORKERS) as ex:
        return sum(ex.map(lambda f: len(cpu_frame(port, det, rec, f, bias)), frames))


def cpu_sample_text(n_frames, boxes):
    cores = os.cpu_count() or 1
    return ("%d frames of the same workload, %d frames in parallel (the reference's ThreadPoolExecutor(4)) x %d torch "
            "threads each = %d cores, per-frame detect + per-crop recognise as pipeliine.py:117-125, %.1f boxes/frame"
            % (n_frames, CPU_WORKERS, max(1, cores // CPU_WORKERS), cores, boxes / max(n_frames, 1)))


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    port, det, rec = cpu_models()
    n = args.warmup + args.steps
    frames = port.synthetic_frames(min(n, 4), SRC_H, SRC_W, seed=0)
    bias = port.planted_logit_bias(1, DET_H, DET_W, seed=0, boxes=BOXES)[0]
    if args.warmup:
        cpu_frames(port, det, rec, [frames[i % len(frames)] for i in range(args.warmup)], bias)
    t0 = time.perf_counter()
    nb = cpu_frames(port, det, rec, [frames[(args.warmup + i) % len(frames)] for i in range(args.steps)], bias)
    dt = time.perf_counter() - t0
    fps = args.steps / dt
    cores = os.cpu_count() or 1
    line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.batch, 1),
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
                             "sample": "1 frame per step; " + cpu_sample_text(args.steps, nb)},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_config(batch, world, inflight=1):
    return {"workload": "configs[2]: full pipeline DBNet-ResNet18 detect (1080p -> 736x1312, fused DB head, box "
                        "extraction) + CRNN recognise (32x%d crops, CTC greedy), ~50 planted boxes/frame" % CROP_W,
            "frame": [SRC_H, SRC_W], "det": [DET_H, DET_W], "crop": [32, CROP_W], "boxes_per_frame": BOXES,
            "frames_per_step_per_gpu": batch, "batches_in_flight": inflight, "parallelism": "frame-sharded dp%d" % world,
            "l2": "frame pool of 32 distinct 1080p frames (199 MB) + per-step activations exceed the 126 MB L2"}


# ------------------------------------------------------------------------------------------------ GPU arm
def main():
    global CROP_W, GF_CRNN_PER_CROP
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp16", "fp32"],
                    help="bf16 = the shipped speed tier; fp16 = the same kernels over IEEE half (libvtd_b200_f16.so); "
                         "fp32 = the CUDA-core parity tier")
    ap.add_argument("--inflight", type=int, default=3, help="batches in flight (contexts/streams/host threads)")
    ap.add_argument("--cpu-frames", type=int, default=24, help="frames of the bounded CPU-baseline sample (~10 s of host time)")
    ap.add_argument("--crop-w", type=int, default=CROP_W, choices=sorted(GF_CRNN),
                    help="recogniser crop width: 128 = the reference's text_recognizer.py:118 (default, the larger "
                         "workload), 100 = BASELINE.json configs[2] as worded")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-out", default=None, help="write the per-op device-time table (JSON) here")
    args = ap.parse_args()
    CROP_W, GF_CRNN_PER_CROP = args.crop_w, GF_CRNN[args.crop_w]
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from video_text_detection_system_b200 import _lib, parallel, synthetic     # the oracle is not imported on this arm

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: no CUDA device visible (there is no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B = args.batch
    POOL = 32
    NW = max(1, args.inflight)                    # batches in flight: one context + stream + host thread each
    det_sd, rec_sd = synthetic.random_state_dicts(seed=0)     # random-init weights of the reference architecture
    engines = []
    for _ in range(NW):
        e = _lib.Engine(device=local_rank, backbone=18, dtype=args.dtype, det_h=DET_H, det_w=DET_W, crop_w=CROP_W,
                        max_batch=B, max_boxes=64, max_src_h=SRC_H, max_src_w=SRC_W)
        e.load_detector(det_sd)
        e.load_recognizer(rec_sd)
        engines.append(e)
    eng = engines[0]
    streams = [torch.cuda.ExternalStream(e.stream(), device=dev) for e in engines]
    main_stream = torch.cuda.current_stream()

    # synthetic inputs: every rank owns its shard of a global pool (rank-strided), seeded
    rng = np.random.default_rng(1000 + rank)
    host_pool = torch.from_numpy(rng.integers(0, 256, (POOL, SRC_H, SRC_W, 3), dtype=np.uint8)).pin_memory()
    dev_pool = host_pool.to(dev)
    bias = torch.from_numpy(synthetic.planted_logit_bias(B, DET_H, DET_W, seed=7 + rank, boxes=BOXES)).to(dev)
    frame_bytes = SRC_H * SRC_W * 3
    import ctypes as C
    import queue

    def ptrs_of(base_ptr, step):
        arr = (C.c_void_p * B)()
        for i in range(B):
            arr[i] = base_ptr + ((step * B + i) % POOL) * frame_bytes
        return arr

    rec_t, cnt_t, host_rec, host_cnt = [], [], [], []
    for e in engines:
        rp, cp = e.device_records()
        rec_t.append(parallel.device_bytes_as_tensor(rp, B * 64 * 128, dev).view(B, 64 * 128))
        cnt_t.append(parallel.device_bytes_as_tensor(cp, B * 4, dev).view(torch.int32))
        host_rec.append(torch.empty((B, 64 * 128), dtype=torch.uint8).pin_memory())
        host_cnt.append(torch.empty((B,), dtype=torch.int32).pin_memory())

    def step_resident(w, i):
        engines[w].run_batch_raw(ptrs_of(dev_pool.data_ptr(), i), B, SRC_H, SRC_W, SRC_W * 3, True, 0.5, True,
                                 bias.data_ptr())

    def step_e2e(w, i):
        engines[w].run_batch_raw(ptrs_of(host_pool.data_ptr(), i), B, SRC_H, SRC_W, SRC_W * 3, False, 0.5, True,
                                 bias.data_ptr(), host_rec[w].data_ptr(), host_cnt[w].data_ptr())

    # worker threads: ctypes releases the GIL inside the library, so NW batches really are in flight.  With N>1 GPUs
    # every step ends with the gather of its records to rank 0 (NCCL); collectives must be issued in the same order
    # on every rank, so the workers take turns in step order (a condition variable), which still lets the next
    # batches run while a gather is in flight.
    gather_cv = threading.Condition()
    gather_next = [0]

    def gather_in_order(w, i):
        with gather_cv:
            while gather_next[0] != i:
                gather_cv.wait()
            main_stream.wait_stream(streams[w])
            with torch.cuda.stream(main_stream):
                parallel.gather_records(rec_t[w], cnt_t[w], 0)
            streams[w].wait_stream(main_stream)
            gather_next[0] = i + 1
            gather_cv.notify_all()

    class Worker(threading.Thread):
        def __init__(self, w):
            super().__init__(daemon=True)
            self.w, self.q, self.done = w, queue.Queue(), queue.Queue()

        def run(self):
            torch.cuda.set_device(local_rank)
            while True:
                job = self.q.get()
                if job is None:
                    return
                fn, steps = job
                try:
                    for i in steps:
                        fn(self.w, i)
                        if world > 1:
                            gather_in_order(self.w, i)
                    self.done.put(None)
                except Exception as ex:      # surface failures in the main thread
                    self.done.put(ex)

    workers = [Worker(w) for w in range(NW)]
    for wk in workers:
        wk.start()

    def run_steps(fn, first, count, nw=None):
        """`count` steps starting at index `first`, dealt round-robin to the in-flight contexts."""
        nw = nw or NW
        gather_next[0] = first
        for w in range(nw):
            workers[w].q.put((fn, list(range(first + w, first + count, nw))))
        for w in range(nw):
            r = workers[w].done.get()
            if r is not None:
                raise r

    def timed(fn, steps, warmup, profile=False, nw=None):
        run_steps(fn, 0, warmup, nw)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        if profile:
            eng.set_profiling(True)
        l0 = sum(e.launch_count() for e in engines)
        e0 = torch.cuda.Event(enable_timing=True)
        ends = [torch.cuda.Event(enable_timing=True) for _ in range(NW + 1)]
        e0.record(main_stream)
        for st in streams:
            st.wait_event(e0)                  # nothing of the timed region starts before e0
        run_steps(fn, warmup, steps, nw)
        for st, ev in zip(streams, ends):
            ev.record(st)
        ends[NW].record(main_stream)
        torch.cuda.synchronize()
        ms = max(e0.elapsed_time(ev) for ev in ends)
        launches = sum(e.launch_count() for e in engines) - l0
        prof = None
        if profile:
            eng.set_profiling(False)
            prof = {"detector": eng.op_profile(0), "recogniser": eng.op_profile(1), "stages": eng.op_profile(2)}
        if world > 1:
            dist.barrier()
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, launches, prof

    sampler = ClockSampler(local_rank)
    sampler.start()
    ms, launches, _ = timed(step_resident, args.steps, args.warmup)
    sampler.stop_flag = True
    sampler.join(timeout=2)
    ms_e2e, _, _ = timed(step_e2e, args.steps, args.warmup)
    # per-kernel CUDA-event times for the roofline: same steps, ONE batch in flight so that a kernel's bracket
    # holds only that kernel (with several streams the bracket also counts time spent queued behind the other stream)
    ms_1, _, prof = timed(step_resident, max(3, args.steps // 2), 3, profile=True, nw=1)
    for wk in workers:
        wk.q.put(None)
    cnt_t = cnt_t[0]

    # sanity: the path really produced ~50 boxes per frame with text
    torch.cuda.synchronize()
    counts = cnt_t.cpu().numpy()
    frames_total = args.steps * B * world
    value = frames_total / (ms / 1e3)
    e2e = frames_total / (ms_e2e / 1e3)

    pk = peaks()
    # roofline of the dominant kernel family: every tcgen05 conv launch of the timed region
    tc_flops = tc_ms = 0.0
    top = None
    for which, per_unit in (("detector", B), ("recogniser", int(counts.sum()))):
        for op in prof[which]:
            if op["kind"] != 0 or op["launches"] == 0:
                continue
            units = per_unit
            fl = 2.0 * units * op["Ho"] * op["Wo"] * op["Cout"] * op["Cin"] * op["KH"] * op["KW"] * op["launches"]
            op["gflop"] = fl / 1e9
            op["tflops"] = fl / (op["ms"] * 1e-3) / 1e12 if op["ms"] > 0 else None
            if op["tensor_core"]:
                tc_flops += fl
                tc_ms += op["ms"]
                if top is None or op["ms"] > top["ms"]:
                    top = op
    if top is not None:
        per_launch_flops = top["gflop"] * 1e9 / top["launches"]
        per_launch_ms = top["ms"] / top["launches"]
        ach = per_launch_flops / (per_launch_ms * 1e-3) / 1e12
        roof = {"bound": "tensor", "kernel": "tcgen05 implicit-GEMM conv (CTA pairs, cta_group::2) %dx%d %d->%d k%d" %
                (top["H"], top["W"], top["Cin"], top["Cout"], top["KH"]),
                "achieved": ach, "peak": pk["tflops"], "unit": "TFLOP/s", "frac": ach / pk["tflops"],
                "traffic": top_kernel_traffic(top)[0], "traffic_detail": top_kernel_traffic(top)[1],
                "peak_source": pk["source"] + " bf16_tflops_sustained (kernel timed inside a long step)",
                "frac_of_burst_peak": ach / pk["tflops_burst"],
                "ms_per_launch": per_launch_ms, "gflop_per_launch": per_launch_flops / 1e9,
                "all_tc_convs": {"tflops": tc_flops / (tc_ms * 1e-3) / 1e12 if tc_ms > 0 else None,
                                 "share_of_step": tc_ms / ms_1 if ms_1 > 0 else None,
                                 "note": "per-kernel times from a pass with one batch in flight"}}
    else:
        fl = 2.0 * 0
        roof = {"bound": "tensor", "achieved": None, "peak": pk["tflops"], "unit": "TFLOP/s", "frac": None, "traffic": None}
    alg_gf = GF_DET_PER_FRAME + GF_CRNN_PER_CROP * float(counts.mean())
    hbm_stages = None
    try:    # reporting only: never let it cost the bench line
        recs = rec_t[0].cpu().numpy().reshape(B, 64, 128).view(_lib.RECORD_DTYPE).reshape(B, 64)
        bb = np.concatenate([recs[i]["bbox"][:int(counts[i])] for i in range(B)]).astype(np.int64)
        crop_src = int(((bb[:, 2] - bb[:, 0]) * (bb[:, 3] - bb[:, 1])).sum()) * 3
        hbm_stages = hbm_stage_rooflines(prof["stages"], max(3, args.steps // 2), B, int(counts.sum()), crop_src, pk["hbm"])
    except Exception as ex:
        print("hbm stage table unavailable: %r" % (ex,), file=sys.stderr)
    if args.profile_out and rank == 0:
        os.makedirs(os.path.dirname(os.path.abspath(args.profile_out)), exist_ok=True)
        json.dump({"ms_total": ms_1, "steps": max(3, args.steps // 2), "batch": B, "ops": prof}, open(args.profile_out, "w"), indent=1)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        p2, det, rec = cpu_models()
        fr = host_pool[:max(1, args.cpu_frames)].numpy()
        b0 = bias[0].cpu().numpy()
        cpu_frames(p2, det, rec, list(fr[:CPU_WORKERS]), b0)       # warm-up
        t0 = time.perf_counter()
        nb = cpu_frames(p2, det, rec, list(fr), b0)
        dt = time.perf_counter() - t0
        cpu = {"value": len(fr) / dt, "unit": "frames/s", "cores": os.cpu_count() or 1, "kind": "port",
               "sample": cpu_sample_text(len(fr), nb)}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": args.dtype, "data": "synthetic", "config": workload_config(B, world, NW),
                "e2e": {"value": e2e, "unit": "frames/s", "h2d_bytes_per_step": B * frame_bytes,
                        "d2h_bytes_per_step": B * 64 * 128 + B * 4, "ms_per_step": ms_e2e / args.steps},
                "gpu_launches": int(launches), "clocks": sampler.result(), "roofline": roof,
                "cpu_baseline": cpu,
                "hbm_stages": hbm_stages,
                "boxes_per_frame": float(counts.mean()),
                "alg_gflop_per_frame": alg_gf,
                "e2e_tensor_frac": (alg_gf * 1e9 * value / world) / (pk["tflops"] * 1e12)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

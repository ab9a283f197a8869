#!/usr/bin/env python
"""bench.py -- 1080p detect+recognize frames/s of the B200 hot path (BASELINE.json metric).

Workload (default, BASELINE.json configs[2], the configuration the metric "detect+recognize frames/sec" is quoted on;
it fits one GPU): synthetic 1080p BGR frames -> 736x1312 -> DBNet-ResNet18 + fused DB head -> box extraction
(~50 planted boxes/frame) -> 50 crops/frame 32x128 -> CRNN -> CTC greedy, in the shipped tcgen05 speed
tier (16-bit storage = IEEE half, fp32 accumulate; --dtype).  A step is one batch of --batch frames through the whole path.
`--config 5` runs BASELINE configs[4] instead: 2160x3840 frames -> 2176x3840, DBNet-ResNet50, CRNN recogniser.

  value : frames/s with the frames already resident in HBM (frame pool larger than L2), results left on device
  e2e   : frames/s through vtd_run_batch (the C-ABI call) with HOST (pinned) frames: H2D of every frame and D2H of the
          records inside the timed region
  e2e_api      : the same through the reference-facing Python surface, VideoTextPipeline.detect_and_recognize: NumPy
                 frames in (pageable, and from pinned buffers), result dictionaries out
  sustained    : `value` measured again over BASELINE configs[3]'s 3000 frames (>= 1 s of back-to-back steps), clocks recorded
  roofline     : dominant kernel (tcgen05 implicit-GEMM conv), algorithmic FLOPs / CUDA-event time of that kernel timed
                 alone (one batch in flight), against MEASURED_PEAKS.json's burst bf16 figure
  cpu_baseline : oracle/port.py (the reference's PyTorch/PIL/OpenCV arithmetic) on the host cores, bounded sample, the
                 reference's own control flow; cpu_baseline_batched: the generous variant (one batched forward)
  hbm_stages   : the non-GEMM stages (preprocess, DB head tail, box extraction, crop gather, CTC): algorithmic bytes
                  / CUDA-event time against MEASURED_PEAKS.json hbm_gbs

`--crop-w 100` runs BASELINE's wording of configs[2] (32x100 crops; default 128 = the reference's own width, the larger
workload).  `--dtype bf16` runs the bfloat16-storage build of the speed tier, `--dtype fp32` the CUDA-core parity tier.

`--impl reference` times the CPU path alone (one frame per step).  N>1: one process per GPU (torchrun), frames
sharded by rank; every step ends with ONE collective that gathers the step's records + counts to rank 0 (NCCL, on a
dedicated communication stream), rank 0 copies the gathered block to the host inside the timed region and checks the
detection count; time = max over ranks.
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

# BASELINE.json configs (1-based, as SURVEY.md 8d numbers them): algorithmic GFLOP per frame from SURVEY.md 8d
WORKLOADS = {
    3: {"label": "configs[2]", "src": (1080, 1920), "det": (736, 1312), "backbone": 18, "gf_det": 184.51, "batch": 16,
        "inflight": 3, "pool": 32, "metric": "1080p detect+recognize frames/sec", "plant": 8.0,
        "text": "full pipeline DBNet-ResNet18 detect (1080p -> 736x1312, fused DB head, box extraction)"},
    5: {"label": "configs[4]", "src": (2160, 3840), "det": (2176, 3840), "backbone": 50, "gf_det": 2450.07, "batch": 4,
        "inflight": 2, "pool": 8, "metric": "4K (3840x2160) detect+recognize frames/sec",
        # the planted logits must exceed the net's own: random-init ResNet50 (randomised BN) reaches +-150, ResNet18 stays within +-8
        "plant": 1000.0,
        "text": "full pipeline DBNet-ResNet50 detect (2160x3840 -> 2176x3840, fused DB head, box extraction)"},
}
WL = WORKLOADS[3]
SRC_H, SRC_W = WL["src"]
DET_H, DET_W = WL["det"]
CROP_W = 128
BOXES = 50
KMAX = 64                           # record slots per frame (max_boxes of the contexts)
GF_DET_PER_FRAME = WL["gf_det"]     # SURVEY.md 8d, live layers
GF_CRNN = {128: 1.787, 100: 1.394}  # SURVEY.md 8d, GFLOP per crop @32x128 (reference default) / @32x100 (BASELINE wording)
GF_CRNN_PER_CROP = GF_CRNN[CROP_W]
METRIC = WL["metric"]


def select_workload(config: int, crop_w: int):
    global WL, SRC_H, SRC_W, DET_H, DET_W, GF_DET_PER_FRAME, METRIC, CROP_W, GF_CRNN_PER_CROP
    WL = WORKLOADS[config]
    SRC_H, SRC_W = WL["src"]
    DET_H, DET_W = WL["det"]
    GF_DET_PER_FRAME, METRIC = WL["gf_det"], WL["metric"]
    CROP_W, GF_CRNN_PER_CROP = crop_w, GF_CRNN[crop_w]


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"tflops": float(d.get("bf16_tflops_sustained", 1357.9)), "tflops_burst": float(d.get("bf16_tflops", 1635.9)),
                "hbm": float(d.get("hbm_gbs", 6531.9)), "source": "measured (MEASURED_PEAKS.json)"}
    return {"tflops": 1400.0, "tflops_burst": 1590.0, "hbm": 6650.0, "source": "fallback (B200_PROFILING.md)"}


def top_kernel_traffic(op):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed
    `ncu --set full` capture, when it is the same layer; else None."""
    for name in ("r02_top_kernel_ncu.json", "r01_top_kernel_ncu.json"):
        p = os.path.join(ROOT, "profiles", name)
        try:
            d = json.load(open(p))
            if (op["H"], op["W"], op["Cin"], op["Cout"], op["KH"]) == (184, 328, 256, 256, 3):
                return float(d["dram_bytes_per_launch"]), {"algorithmic_bytes_per_launch": d["algorithmic_bytes_per_launch"],
                                                           "unit": "bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum)",
                                                           "source": "profiles/" + name}
        except Exception:
            pass
    return None, None


def hbm_stage_rooflines(stages, steps, batch, crops_per_step, crop_src_bytes_per_step, hbm_gbs, bias_plane=True):
    """Achieved HBM GB/s of the stages that are HBM-bound by their bytes (north_star: "achieved HBM GB/s for the
    elementwise, post-processing and gather stages").  `stages` = Engine.op_profile(2) of a profiled pass of `steps`
    steps with ONE batch in flight (CUDA-event time per stage, summed over the steps).  ALGORITHMIC bytes per step,
    SURVEY.md 8d (16-bit tier, s = 2): what the stage must read and write once, not what the kernels happen to move."""
    px = DET_H * DET_W
    per_step = {
        # K1: BGR frame in, 3 normalised channels out
        "preprocess": batch * (SRC_H * SRC_W * 3 + 3 * px * 2),
        # K3 (when the tail runs as its own kernel): feat [Hd/4, Wd/4, 128] 16-bit in; prob + thresh fp32 and the u8 mask
        # out; + the planted fp32 logit plane
        "head_tail": batch * ((px // 16) * 128 * 2 + 2 * px * 4 + px + (px * 4 if bias_plane else 0)),
        # K4-K6: mask in, the int32 label plane written and read once, the records out
        "boxes": batch * (px + 2 * px * 4 + 64 * 128),
        # K7: the source pixels under the boxes in, 32 x crop_w x 3 16-bit values per crop out
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
            time.sleep(0.02)

    def finish(self):
        self.stop_flag = True
        self.join(timeout=2)
        return self.result()

    def result(self):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": []}
        return {"sm_mhz": int(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def bind_to_gpu_numa_node(index: int):
    """Keep this rank's threads (and with them its pinned pool, first-touched below) on the CPU cores local to its GPU."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {w * 64 + b for w, v in enumerate(words) for b in range(64) if (int(v) >> b) & 1 and w * 64 + b < ncpu}
        allowed = cpus & set(os.sched_getaffinity(0))
        if allowed:
            os.sched_setaffinity(0, allowed)
            return len(allowed)
    except Exception:
        pass
    return None


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_models():
    """The oracle's PyTorch modules holding the SAME random-init weights the B200 arm loads (synthetic.random_state_dicts)."""
    import torch
    from oracle import port
    from video_text_detection_system_b200 import synthetic
    torch.set_num_threads(os.cpu_count() or 1)
    bb = "resnet50" if WL["backbone"] == 50 else "resnet18"
    det_sd, rec_sd = synthetic.random_state_dicts(seed=0, backbone=bb)
    det, rec = port.build_dbnet(bb, seed=0), port.build_crnn(seed=0)
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


def cpu_frames_batched(port, det, rec, frames, bias):
    """The generous CPU bound: what the reference could do with its own modules if it batched --
    ONE batched DBNet forward over the frames and ONE recognize_batch over all crops, every core as intra-op threads."""
    import torch
    torch.set_num_threads(os.cpu_count() or 1)
    x = torch.cat([port.preprocess(f, DET_H, DET_W) for f in frames])
    with torch.no_grad():
        prob = port.dbnet_forward(det, x, torch.from_numpy(bias)[None, None].expand(len(frames), 1, DET_H, DET_W))["probability"].numpy()
    crops = []
    for f, p in zip(frames, prob):
        for d in port.post_process(p[0], f.shape[1], f.shape[0], 0.5, DET_H, DET_W):
            x1, y1, x2, y2 = d["bbox"]
            c = f[y1:y2, x1:x2]
            if c.size:
                crops.append(c)
    res = port.recognize_batch(rec, crops, CROP_W) if crops else []
    return len(res)


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
    bias = port.planted_logit_bias(1, DET_H, DET_W, seed=0, boxes=BOXES, inside=WL["plant"], outside=-WL["plant"])[0]
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
            "config": workload_config(args.batch, world),
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
                             "sample": "1 frame per step; " + cpu_sample_text(args.steps, nb)},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_config(batch, world):
    """Names the workload only (identical for the B200 arm and the reference arm; how many batches the B200 arm keeps in
    flight is an execution detail and is reported beside it, as `inflight`)."""
    pool_mb = WL["pool"] * SRC_H * SRC_W * 3 / 1e6
    return {"workload": "%s: %s + CRNN recognise (32x%d crops, CTC greedy), ~50 planted boxes/frame"
                        % (WL["label"], WL["text"], CROP_W),
            "frame": [SRC_H, SRC_W], "det": [DET_H, DET_W], "crop": [32, CROP_W], "boxes_per_frame": BOXES,
            "frames_per_step_per_gpu": batch, "parallelism": "frame-sharded dp%d" % world,
            "l2": "frame pool of %d distinct frames (%.0f MB) + per-step activations exceed the 126 MB L2" % (WL["pool"], pool_mb)}


# ------------------------------------------------------------------------------------------------ GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", type=int, default=3, choices=sorted(WORKLOADS),
                    help="BASELINE.json config, 1-based as SURVEY.md 8d numbers them: 3 = configs[2] (1080p, ResNet18; the "
                         "metric's configuration), 5 = configs[4] (4K, ResNet50)")
    ap.add_argument("--batch", type=int, default=None)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--dtype", default="fp16", choices=["fp16", "bf16", "fp32"],
                    help="fp16 = the shipped speed tier (tcgen05, IEEE half storage); bf16 = the same kernels over bfloat16 "
                         "(libvtd_b200_bf16.so); fp32 = the CUDA-core parity tier")
    ap.add_argument("--inflight", type=int, default=None, help="batches in flight (contexts/streams/host threads)")
    ap.add_argument("--cpu-frames", type=int, default=None, help="frames of the bounded CPU-baseline sample (~10 s of host time)")
    ap.add_argument("--crop-w", type=int, default=CROP_W, choices=sorted(GF_CRNN),
                    help="recogniser crop width: 128 = the reference's text_recognizer.py:118 (default, the larger "
                         "workload), 100 = BASELINE.json configs[2] as worded")
    ap.add_argument("--idle-ms", type=int, default=500, help="idle time between timed legs (0: legs back to back)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip e2e_api, e2e_nv12 and the sustained run")
    ap.add_argument("--sustained-frames", type=int, default=None,
                    help="frames of the sustained run (default: BASELINE configs[3]'s 3000 at 1080p, 600 at 4K)")
    ap.add_argument("--trocr", action="store_true",
                    help="also time the transformer recogniser (TrOCR branch, base configuration, random-init) on a chunk of "
                         "crops; on by default with --config 5 (BASELINE configs[4]: 'CRNN+Transformer recognizer')")
    ap.add_argument("--profile-out", default=None, help="write the per-op device-time table (JSON) here")
    args = ap.parse_args()
    select_workload(args.config, args.crop_w)
    args.batch = args.batch or WL["batch"]
    args.inflight = args.inflight or WL["inflight"]
    if args.cpu_frames is None:
        args.cpu_frames = 24 if args.config == 3 else 4
    if args.sustained_frames is None:
        args.sustained_frames = 3000 if args.config == 3 else 600
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
    numa_cpus = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B = args.batch
    POOL = WL["pool"]
    NW = max(1, args.inflight)                    # batches in flight: one context + stream + host thread each
    bb = "resnet50" if WL["backbone"] == 50 else "resnet18"
    det_sd, rec_sd = synthetic.random_state_dicts(seed=0, backbone=bb)     # random-init weights of the reference architecture
    engines = []
    for _ in range(NW):
        e = _lib.Engine(device=local_rank, backbone=WL["backbone"], dtype=args.dtype, det_h=DET_H, det_w=DET_W, crop_w=CROP_W,
                        max_batch=B, max_boxes=KMAX, max_src_h=SRC_H, max_src_w=SRC_W)
        e.load_detector(det_sd)
        e.load_recognizer(rec_sd)
        engines.append(e)
    eng = engines[0]
    streams = [torch.cuda.ExternalStream(e.stream(), device=dev) for e in engines]
    main_stream = torch.cuda.current_stream()
    comm_stream = torch.cuda.Stream(device=dev)   # the per-step gather runs here: worker streams never wait on each other

    # synthetic inputs: every rank owns its shard of a global pool (rank-strided), seeded
    rng = np.random.default_rng(1000 + rank)
    host_pool = torch.from_numpy(rng.integers(0, 256, (POOL, SRC_H, SRC_W, 3), dtype=np.uint8)).pin_memory()
    dev_pool = host_pool.to(dev)
    bias = torch.from_numpy(synthetic.planted_logit_bias(B, DET_H, DET_W, seed=7 + rank, boxes=BOXES, inside=WL["plant"],
                                                         outside=-WL["plant"])).to(dev)
    frame_bytes = SRC_H * SRC_W * 3
    nv12_bytes = SRC_H * SRC_W * 3 // 2
    import ctypes as C
    import queue

    def ptrs_of(base_ptr, step, fbytes=frame_bytes):
        arr = (C.c_void_p * B)()
        for i in range(B):
            arr[i] = base_ptr + ((step * B + i) % POOL) * fbytes
        return arr

    # results: per context ONE device block (records, then counts: vtd_get_records) and its pinned host mirror
    blk_bytes = parallel.packed_bytes(B, KMAX)
    blk_t, host_blk, gath_t, host_gath = [], [], [], []
    for e in engines:
        rp, cp = e.device_records()
        assert cp == rp + B * KMAX * 128
        blk_t.append(parallel.device_bytes_as_tensor(rp, blk_bytes, dev))
        host_blk.append(torch.empty((blk_bytes,), dtype=torch.uint8).pin_memory())
        if world > 1:
            gath_t.append(torch.empty((world, blk_bytes), dtype=torch.uint8, device=dev))
            host_gath.append(torch.empty((world, blk_bytes), dtype=torch.uint8).pin_memory() if rank == 0 else None)

    def step_resident(w, i):
        engines[w].run_batch_raw(ptrs_of(dev_pool.data_ptr(), i), B, SRC_H, SRC_W, SRC_W * 3, True, 0.5, True,
                                 bias.data_ptr())

    def step_e2e(w, i):
        hb = host_blk[w].data_ptr()
        engines[w].run_batch_raw(ptrs_of(host_pool.data_ptr(), i), B, SRC_H, SRC_W, SRC_W * 3, False, 0.5, True,
                                 bias.data_ptr(), hb, hb + B * KMAX * 128)

    nv12_pool = [None]

    def step_e2e_nv12(w, i):
        hb = host_blk[w].data_ptr()
        engines[w].run_batch_raw(ptrs_of(nv12_pool[0].data_ptr(), i, nv12_bytes), B, SRC_H, SRC_W, SRC_W, False, 0.5, True,
                                 bias.data_ptr(), hb, hb + B * KMAX * 128, pixfmt=_lib.VTD_PIX_NV12)

    # worker threads: ctypes releases the GIL inside the library, so NW batches really are in flight.  With N>1 GPUs
    # every step ends with the gather of its result block to rank 0: ONE collective per step, issued in step order (NCCL
    # needs the same order on every rank) on the communication stream, which waits for that step's batch only; the
    # context's stream waits for ITS gather before the next batch overwrites the block.  Rank 0 copies the gathered
    # blocks to pinned host memory inside the timed region.
    gather_cv = threading.Condition()
    gather_next = [0]
    gather_events = []            # (start, end) CUDA events of every gather on the communication stream
    gather_hosts = []             # rank 0: host tensors of the last gathers, for the count check

    def gather_in_order(w, i):
        with gather_cv:
            while gather_next[0] != i:
                gather_cv.wait()
            done = torch.cuda.Event()
            done.record(streams[w])
            comm_stream.wait_event(done)
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(comm_stream):
                g0.record(comm_stream)
                parallel.gather_packed(blk_t[w], 0, out=gath_t[w])
                g1.record(comm_stream)
                if rank == 0:
                    host_gath[w].copy_(gath_t[w], non_blocking=True)
                fin = torch.cuda.Event()
                fin.record(comm_stream)
            streams[w].wait_event(fin)
            gather_events.append((g0, g1))
            if rank == 0:
                gather_hosts.append((i, w))
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
                fn, steps, gather = job
                try:
                    for i in steps:
                        fn(self.w, i)
                        if gather:
                            gather_in_order(self.w, i)
                    self.done.put(None)
                except Exception as ex:      # surface failures in the main thread
                    self.done.put(ex)

    workers = [Worker(w) for w in range(NW)]
    for wk in workers:
        wk.start()

    def run_steps(fn, first, count, nw=None, gather=None):
        """`count` steps starting at index `first`, dealt round-robin to the in-flight contexts."""
        nw = nw or NW
        gather = (world > 1) if gather is None else gather
        gather_next[0] = first
        for w in range(nw):
            workers[w].q.put((fn, list(range(first + w, first + count, nw)), gather))
        for w in range(nw):
            r = workers[w].done.get()
            if r is not None:
                raise r

    legs_done = [0]

    def timed(fn, steps, warmup, profile=False, nw=None, sample_clocks=False):
        # every leg is its own burst: a short idle lets the board's power-cap average recover, so that a leg is not timed
        # at the clocks its predecessor left behind (`sustained` is the leg that measures the capped state)
        if args.idle_ms > 0 and legs_done[0]:
            torch.cuda.synchronize()
            time.sleep(args.idle_ms / 1e3)
        legs_done[0] += 1
        run_steps(fn, 0, warmup, nw)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        del gather_events[:], gather_hosts[:]
        if profile:
            eng.set_profiling(True)
        sampler = ClockSampler(local_rank) if sample_clocks else None
        if sampler:
            sampler.start()
        l0 = sum(e.launch_count() for e in engines)
        e0 = torch.cuda.Event(enable_timing=True)
        ends = [torch.cuda.Event(enable_timing=True) for _ in range(NW + 2)]
        e0.record(main_stream)
        for st in streams + [comm_stream]:
            st.wait_event(e0)                  # nothing of the timed region starts before e0
        run_steps(fn, warmup, steps, nw)
        for st, ev in zip(streams + [comm_stream, main_stream], ends):
            ev.record(st)
        torch.cuda.synchronize()
        ms = max(e0.elapsed_time(ev) for ev in ends)
        clocks = sampler.finish() if sampler else None
        launches = sum(e.launch_count() for e in engines) - l0
        prof = None
        if profile:
            eng.set_profiling(False)
            prof = {"detector": eng.op_profile(0), "recogniser": eng.op_profile(1), "stages": eng.op_profile(2)}
        info = {"launches": launches, "clocks": clocks}
        if world > 1:
            info["gather_ms"] = float(np.mean([a.elapsed_time(b) for a, b in gather_events])) if gather_events else None
            if rank == 0 and gather_hosts:
                # the gathered blocks reached rank 0's HOST inside the timed region: every rank's every frame holds its boxes
                tot = 0
                for _, w in gather_hosts[-NW:]:
                    _, cnts = parallel.split_packed(host_gath[w], B, KMAX)
                    tot += int(cnts.sum())
                info["gather_checked"] = {"blocks": len(gather_hosts[-NW:]), "detections": tot,
                                          "expected": len(gather_hosts[-NW:]) * world * B * BOXES,
                                          "ok": tot == len(gather_hosts[-NW:]) * world * B * BOXES}
            dist.barrier()
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, info, prof

    ms, info, _ = timed(step_resident, args.steps, args.warmup, sample_clocks=True)
    launches, clocks = info["launches"], info["clocks"]
    ms_e2e, info_e2e, _ = timed(step_e2e, args.steps, args.warmup, sample_clocks=True)
    frames_total = args.steps * B * world
    extras = {}
    if not args.no_extras:
        # sustained: BASELINE configs[3]'s frame count through the resident path, back to back (>= 1 s at 1080p)
        sus_steps = max(args.steps, (args.sustained_frames // world + B - 1) // B)
        ms_s, info_s, _ = timed(step_resident, sus_steps, 3, sample_clocks=True)
        extras["sustained"] = {"frames": sus_steps * B * world, "seconds": ms_s / 1e3, "value": sus_steps * B * world / (ms_s / 1e3),
                               "unit": "frames/s", "clocks": info_s["clocks"]}
        # decoder-surface ingest: NV12 host frames (half the H2D bytes of BGR), otherwise the same call
        nv12_pool[0] = torch.from_numpy(rng.integers(0, 256, (POOL, SRC_H * 3 // 2, SRC_W), dtype=np.uint8)).pin_memory()
        ms_n, _, _ = timed(step_e2e_nv12, args.steps, args.warmup)
        extras["e2e_nv12"] = {"value": frames_total / (ms_n / 1e3), "unit": "frames/s", "h2d_bytes_per_step": B * nv12_bytes,
                              "d2h_bytes_per_step": blk_bytes, "ms_per_step": ms_n / args.steps,
                              "note": "same vtd_run_batch call on NV12 host surfaces (what a hardware decoder delivers)"}
    # per-kernel CUDA-event times for the roofline: same steps, ONE batch in flight so that a kernel's bracket
    # holds only that kernel (with several streams the bracket also counts time spent queued behind the other stream)
    ms_1, _, prof = timed(step_resident, max(3, args.steps // 2), 3, profile=True, nw=1)
    for wk in workers:
        wk.q.put(None)

    # sanity: the path really produced ~50 boxes per frame with text
    torch.cuda.synchronize()
    blk0 = blk_t[0].cpu().numpy()
    recs0, counts = parallel.split_packed(blk0[None], B, KMAX)
    counts = counts[0]
    value = frames_total / (ms / 1e3)
    e2e = frames_total / (ms_e2e / 1e3)

    # ---- the reference-facing Python surface: NumPy frames in, result dictionaries out (single GPU leg)
    if not args.no_extras and world == 1:
        try:
            extras.update(api_leg(args, _lib, synthetic, det_sd, rec_sd, host_pool, bias, B, NW, local_rank, bb))
        except Exception as ex:
            print("e2e_api unavailable: %r" % (ex,), file=sys.stderr)

    # ---- the sink after the path: annotated frames drawn on the device (row N3), frames resident, one call per batch
    if not args.no_extras and world == 1:
        try:
            extras["overlay"] = overlay_leg(_lib, engines[0], recs0, counts, dev_pool, B, SRC_H, SRC_W)
        except Exception as ex:
            print("overlay leg unavailable: %r" % (ex,), file=sys.stderr)

    if (args.trocr or args.config == 5) and world == 1 and not args.no_extras:
        try:
            extras["trocr"] = trocr_leg(args, _lib, synthetic, local_rank, recs0, counts, host_pool, B)
        except Exception as ex:
            print("trocr leg unavailable: %r" % (ex,), file=sys.stderr)

    pk = peaks()
    # roofline of the dominant kernel family: every tcgen05 conv launch of the timed region
    tc_flops = tc_ms = 0.0
    top = None
    for which, per_unit in (("detector", B), ("recogniser", int(counts.sum()))):
        for op in prof[which]:
            if op["kind"] != 0 or op["launches"] == 0:
                continue
            cin = 3 if op["Cin"] == 4 else op["Cin"]          # the two stems run zero-padded to 4 channels: count the 3 real ones
            fl = 2.0 * per_unit * op["Ho"] * op["Wo"] * op["Cout"] * cin * op["KH"] * op["KW"] * op["launches"]
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
        roof = {"bound": "tensor", "kernel": "tcgen05 implicit-GEMM conv %dx%d %d->%d k%d" %
                (top["H"], top["W"], top["Cin"], top["Cout"], top["KH"]),
                "achieved": ach, "peak": pk["tflops_burst"], "unit": "TFLOP/s", "frac": ach / pk["tflops_burst"],
                "traffic": top_kernel_traffic(top)[0], "traffic_detail": top_kernel_traffic(top)[1],
                "peak_source": pk["source"] + ": bf16_tflops (burst; the kernel is timed alone, bracketed by CUDA events, "
                                              "one batch in flight)",
                "frac_of_sustained_peak": ach / pk["tflops"],
                "ms_per_launch": per_launch_ms, "gflop_per_launch": per_launch_flops / 1e9,
                "all_tc_convs": {"tflops": tc_flops / (tc_ms * 1e-3) / 1e12 if tc_ms > 0 else None,
                                 "share_of_step": tc_ms / ms_1 if ms_1 > 0 else None,
                                 "note": "per-kernel times from a pass with one batch in flight"}}
    else:
        roof = {"bound": "tensor", "achieved": None, "peak": pk["tflops_burst"], "unit": "TFLOP/s", "frac": None, "traffic": None}
    alg_gf = GF_DET_PER_FRAME + GF_CRNN_PER_CROP * float(counts.mean())
    hbm_stages = None
    try:    # reporting only: never let it cost the bench line
        recs = recs0[0].reshape(B, KMAX, 128).view(_lib.RECORD_DTYPE).reshape(B, KMAX)
        bbx = np.concatenate([recs[i]["bbox"][:int(counts[i])] for i in range(B)]).astype(np.int64)
        crop_src = int(((bbx[:, 2] - bbx[:, 0]) * (bbx[:, 3] - bbx[:, 1])).sum()) * 3
        hbm_stages = hbm_stage_rooflines(prof["stages"], max(3, args.steps // 2), B, int(counts.sum()), crop_src, pk["hbm"])
    except Exception as ex:
        print("hbm stage table unavailable: %r" % (ex,), file=sys.stderr)
    if args.profile_out and rank == 0:
        os.makedirs(os.path.dirname(os.path.abspath(args.profile_out)), exist_ok=True)
        json.dump({"ms_total": ms_1, "steps": max(3, args.steps // 2), "batch": B, "ops": prof}, open(args.profile_out, "w"), indent=1)

    cpu = cpu_b = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        p2, det, rec = cpu_models()
        fr = host_pool[:max(1, min(args.cpu_frames, POOL))].numpy()
        b0 = bias[0].cpu().numpy()
        cpu_frames(p2, det, rec, list(fr[:CPU_WORKERS]), b0)       # warm-up
        t0 = time.perf_counter()
        nb = cpu_frames(p2, det, rec, list(fr), b0)
        dt = time.perf_counter() - t0
        cpu = {"value": len(fr) / dt, "unit": "frames/s", "cores": os.cpu_count() or 1, "kind": "port",
               "sample": cpu_sample_text(len(fr), nb)}
        nbt = min(8, len(fr))
        cpu_frames_batched(p2, det, rec, list(fr[:2]), b0)          # warm-up
        t0 = time.perf_counter()
        nb2 = cpu_frames_batched(p2, det, rec, list(fr[:nbt]), b0)
        dt = time.perf_counter() - t0
        cpu_b = {"value": nbt / dt, "unit": "frames/s", "cores": os.cpu_count() or 1, "kind": "port",
                 "sample": "%d frames in ONE batched DBNet forward + ONE recognize_batch over their %d crops, all cores as "
                           "intra-op threads (the generous bound of BASELINE.md section 4; the reference never batches)" % (nbt, nb2)}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": eng.dtype, "data": "synthetic", "config": workload_config(B, world),
                "inflight": NW,
                "e2e": {"value": e2e, "unit": "frames/s", "h2d_bytes_per_step": B * frame_bytes,
                        "d2h_bytes_per_step": blk_bytes, "ms_per_step": ms_e2e / args.steps, "clocks": info_e2e["clocks"]},
                "gpu_launches": int(launches), "clocks": clocks, "roofline": roof,
                "cpu_baseline": cpu, "cpu_baseline_batched": cpu_b,
                "hbm_stages": hbm_stages,
                "boxes_per_frame": float(counts.mean()),
                "alg_gflop_per_frame": alg_gf,
                "e2e_tensor_frac": (alg_gf * 1e9 * value / world) / (pk["tflops"] * 1e12)}
        line.update(extras)
        if world > 1:
            line["gather"] = {"collectives_per_step": 1, "bytes_per_rank_per_step": blk_bytes,
                              "ms_per_step_resident": info.get("gather_ms"), "ms_per_step_e2e": info_e2e.get("gather_ms"),
                              "checked": info_e2e.get("gather_checked") or info.get("gather_checked"),
                              "numa_bound_cpus": numa_cpus}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def overlay_leg(_lib, eng, recs0, counts, dev_pool, B, h, w):
    """ProcessingService._draw_detections (processing_service.py:188-218) for a whole batch through vtd_draw_detections: the
    boxes of the last batch with a label each, drawn into B resident BGR frames (a copy of the pool's first B); wall clock around
    the C-ABI call (it uploads the draw list and synchronises)."""
    import torch
    from video_text_detection_system_b200.sinks import overlay_items
    recs = recs0[0].reshape(B, KMAX, 128).view(_lib.RECORD_DTYPE).reshape(B, KMAX)
    dets = [[{"bbox": [int(v) for v in recs[i][j]["bbox"]], "text": "text%02d" % j, "detection_confidence": float(recs[i][j]["det_conf"])}
             for j in range(int(counts[i]))] for i in range(B)]
    items = overlay_items(dets)
    frames = dev_pool[:B].clone()
    ptrs = [frames[i].data_ptr() for i in range(B)]
    torch.cuda.synchronize()
    eng.draw_detections(ptrs, items, on_device=True, h=h, w=w, pitch=w * 3)          # warm-up (uploads the glyph tables)
    reps = 10
    t0 = time.perf_counter()
    for _ in range(reps):
        eng.draw_detections(ptrs, items, on_device=True, h=h, w=w, pitch=w * 3)
    dt = (time.perf_counter() - t0) / reps
    return {"value": B / dt, "unit": "frames/s", "ms_per_batch": dt * 1e3, "detections": int(len(items)),
            "call": "vtd_draw_detections on %d resident %dx%d BGR frames, %d labelled boxes" % (B, h, w, len(items))}


def trocr_leg(args, _lib, synthetic, device, recs0, counts, host_pool, B):
    """The reference's OTHER recogniser (TransformerRecognizer, text_recognizer.py:39-69; BASELINE configs[4] names
    "CRNN+Transformer"): microsoft/trocr-base-printed's architecture with random-init weights, greedy generate(max_length=50),
    on the crops the detector found in the first frames of the last batch (host BGR crops -> vtd_trocr_generate_crops: H2D,
    device-side 384x384 processor resize, ViT encoder, 49 decoder steps, ids back).  Wall clock around the C-ABI call."""
    model = synthetic.random_trocr_model("base", seed=0)
    chunk = 128                                                   # crops per chunk (one decode loop serves them all)
    eng = _lib.Engine(device=device, dtype=args.dtype if args.dtype != "fp32" else "fp16", det_h=32, det_w=32, max_batch=1, max_boxes=64,
                      max_src_h=32, max_src_w=32)
    eng.load_trocr(model.state_dict(), crops_per_chunk=chunk)
    del model
    recs = recs0[0].reshape(B, KMAX, 128).view(_lib.RECORD_DTYPE).reshape(B, KMAX)
    frames = host_pool.numpy()
    crops = []
    for i in range(B):
        for j in range(int(counts[i])):
            x1, y1, x2, y2 = (int(v) for v in recs[i][j]["bbox"])
            c = frames[i % frames.shape[0]][y1:y2, x1:x2]
            if c.size:
                crops.append(np.ascontiguousarray(c))
    crops = crops[:2 * chunk]
    if not crops:
        raise RuntimeError("no crops to recognise")
    eng.trocr_generate_crops(crops, 50)                           # warm-up with the timed chunk sizes (one decode-step graph per size)
    t0 = time.perf_counter()
    ids, lens = eng.trocr_generate_crops(crops, 50)
    dt = time.perf_counter() - t0
    eng.close()
    gf_crop = 2 * (577 * (768 * 768 * 4 + 2 * 768 * 3072) * 12 + 12 * 2 * 577 * 577 * 768) / 1e9       # encoder, 2 x MAC
    return {"value": len(crops) / dt, "unit": "crops/s", "crops": len(crops), "seconds": dt, "tokens_per_crop": float(lens.mean()),
            "frames_per_s_at_50_crops": len(crops) / dt / BOXES, "encoder_gflop_per_crop": gf_crop,
            "model": "ViT-B/16@384 encoder + 12-layer TrOCR decoder (microsoft/trocr-base-printed configuration, random-init), "
                     "greedy generate(max_length=50), %d crops per chunk" % chunk}


def api_leg(args, _lib, synthetic, det_sd, rec_sd, host_pool, bias, B, NW, device, bb):
    """frames/s through VideoTextPipeline.detect_and_recognize -- the call a user of the reference's surface makes: a list
    of NumPy BGR frames in, the reference's result dictionaries out (H2D, the whole device path, D2H of the records and the
    record -> dict conversion inside the timed region; wall clock, the outputs being host objects).  NW caller threads,
    one pipeline slot each, as process_video keeps them in flight."""
    from video_text_detection_system_b200 import VideoTextPipeline
    P = VideoTextPipeline(use_transformer_ocr=False, backbone=bb, pretrained=False, det_size=(DET_H, DET_W), dtype=args.dtype,
                          crop_w=CROP_W, max_boxes=KMAX, batch_size=B, inflight=NW)
    P.detector.model.load_state_dict(det_sd)
    P.recognizer.model.load_state_dict(rec_sd)
    P.logit_bias_dev = bias.data_ptr()
    POOL = host_pool.shape[0]
    pinned = host_pool.numpy()
    pageable = np.array(pinned, copy=True)
    out = {}
    for key, pool in (("e2e_api", pageable), ("e2e_api_pinned", pinned)):
        def work(w, steps, res):
            n = 0
            for i in steps:
                frames = [pool[(i * B + j) % POOL] for j in range(B)]
                regions = P.detect_and_recognize(frames, slot=w)
                n += sum(len(r) for r in regions)
            res[w] = n
        for timed_run in (False, True):
            nsteps = args.steps if timed_run else max(NW, 3)
            res = [0] * NW
            ts = [threading.Thread(target=work, args=(w, range(w, nsteps, NW), res)) for w in range(NW)]
            t0 = time.perf_counter()
            for t in ts:
                t.start()
            for t in ts:
                t.join()
            dt = time.perf_counter() - t0
        out[key] = {"value": args.steps * B / dt, "unit": "frames/s", "ms_per_step": 1e3 * dt / args.steps,
                    "h2d_bytes_per_step": B * SRC_H * SRC_W * 3, "d2h_bytes_per_step": B * KMAX * 128 + B * 4,
                    "detections_per_frame": sum(res) / (args.steps * B),
                    "call": "VideoTextPipeline.detect_and_recognize(list of %d NumPy frames, %s host memory) -> result dicts, "
                            "%d caller threads" % (B, "pinned" if key.endswith("pinned") else "pageable", NW)}
    return out


if __name__ == "__main__":
    main()

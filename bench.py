#!/usr/bin/env python
"""bench.py -- decoded images/sec of the RTM3D keypoint-heatmap decode path on B200 (+ fraction of the HBM roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg4] [--impl b200|reference] [--scaling weak|strong]
                    [--dtype f32|bf16] [--graph] [--verify]

A "step" = one pass of the hot path (Tier A main decode + Tier B keypoint decode + grouping; models/model.py:29-162)
over one batch of synthetic head outputs that are already resident in HBM.  The default workload is BASELINE.json
configs[3] ("cfg4": 256 images per GPU, 3x96x320 main heat-map + 9-keypoint heat-map + the 16/2/2-channel regression
maps, K=100, thresh 0.4), the configuration the metric's "1/2/4/8 B200" is quoted on.  Images are independent
(SURVEY.md 8e): `--scaling weak` (default) keeps 256 images per GPU as N grows, `--scaling strong` cuts the 256 images into
N shards; for N > 1 every rank's fixed-size detections are delivered to every rank (`--gather`: rows left in the rank's own
symmetric-memory buffer by the select + post kernel and moved by the copy engines on a second stream, rows stored by the
kernel itself into every rank's buffer, or rtm3d_pack_wire + NCCL all-gather).  One JSON line is
printed by rank 0 (contract in the task statement):

  value        whole-job images/s, device-timed (CUDA events around exactly K steps, max over ranks)
  roofline     per kernel: the bytes THAT kernel moves / its CUDA-event duration vs MEASURED_PEAKS.json hbm_gbs (the
               dominant kernel's figures at the top level); step_frac = algorithmic bytes of the whole path /
               ms_per_step -- the honest whole-path number; cold_ms = first launch on a fresh workspace; shift_ms =
               ms_per_step when consecutive batches come from different distributions (the kernels keep no state)
  e2e          same metric through the host-buffer C-ABI entry points (pinned host in -> pinned host out, H2D + D2H
               inside the timed region)
  cpu_baseline the oracle's torch port of the reference decoder on the box's host cores (bounded sample), N = 1 only
  clocks       SM clock / throttle reasons sampled with NVML while the timed region runs

`--impl reference` times the reference's CPU implementation of the same path (oracle port: the reference is Python and
does not travel to the GPU box) on a bounded sample per step and prints the same line with
"impl": "reference".
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "decoded images/sec"
UNIT = "images/s"
THRESH, DOWN = 0.4, 4.0
CPU_SAMPLE_IMAGES = 16          # bounded CPU sample per step of the reference arm / cpu_baseline leg

WORKLOADS = {
    # name: B per GPU (weak scaling) = per job (strong scaling), C, H, W, K, keypoint channels, regression channels of the box decode
    "cfg2": dict(B=32, C=3, H=96, W=320, K=50, kpt=9, note="BASELINE configs[1]: DLA-34 head outputs batch 32, K=50"),
    "cfg2x": dict(B=64, C=3, H=96, W=320, K=100, kpt=9, note="north_star's batch >= 64 point: 64 images, main + 9-kpt heat-map, K=100"),
    "cfg3": dict(B=64, C=3, H=96, W=320, K=100, kpt=0, reg=8,
                 note="BASELINE configs[2]: SMOKE-style centre-keypoint decode (depth/dim/orientation regression, closed-form box) "
                      "batch 64, K=100 -- Tier C parity unpinned (no reference code)"),
    "cfg4": dict(B=256, C=3, H=96, W=320, K=100, kpt=9, note="BASELINE configs[3]: ResNet-18 head outputs batch 256 per GPU, K=100, main + 9-kpt heat-map"),
    "cfg4main": dict(B=256, C=3, H=96, W=320, K=100, kpt=0, note="BASELINE configs[3], main branch only (what the reference's inference() runs today)"),
    "cfg5": dict(B=128, C=3, H=192, W=640, K=100, kpt=9, note="BASELINE configs[4]: 192x640 heat-maps batch 128 per GPU, K=100, main + 9-kpt heat-map"),
}


def algorithmic_bytes_per_image(w, elem=4, n_vert=8):
    """A = C_hm*H*W*e + K*C_reg*32 + out (gathers charged one 32-byte sector per scalar; regression planes
    are NOT counted in full).  Returns (total, main part, keypoint part)."""
    HW, K, Cv = w["H"] * w["W"], w["K"], w["kpt"]
    if w.get("reg"):      # cfg3: C_reg regression channels gathered at the K peaks, 128 B of box parameters per detection
        main = w["C"] * HW * elem + K * w["reg"] * 32 + K * 128 + 4
        return main, main, 0
    main = w["C"] * HW * elem + K * (2 * n_vert + 2) * 32 + K * 100 + 4
    kpt = (Cv * HW * elem + Cv * K * 2 * 32 + K * Cv * 12) if Cv else 0
    return main + kpt, main, kpt


def kernel_bytes_per_image(w, elem=4, n_vert=8):
    """The bytes each kernel of the path moves per image (what its own roofline fraction is computed from): the scan kernel
    streams the heat-maps once; everything else (gathers at 32 B per scalar, outputs) belongs to the kernels behind it."""
    total, _, _ = algorithmic_bytes_per_image(w, elem, n_vert)
    heat = (w["C"] + w["kpt"]) * w["H"] * w["W"] * elem
    return {"scan": heat, "post": total - heat}


def ncu_traffic(workload):
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel, per launch, from the committed ncu capture of
    this workload (profiles/traffic.json, written by tools/summarise_profiles.py); None when there is none."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        with open(path) as f:
            return json.load(f).get(workload)
    except Exception:
        return None


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """SM clock + throttle reasons of one GPU, sampled through NVML while the bench runs."""
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting"}

    def __init__(self, index, uuid=None):
        super().__init__(daemon=True)
        self.index, self.uuid, self.samples, self._stop_evt, self.error = index, uuid, [], threading.Event(), None
        self.max_mhz = None

    def run(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            try:
                h = pynvml.nvmlDeviceGetHandleByUUID(self.uuid.encode() if isinstance(self.uuid, str) else self.uuid)
            except Exception:
                h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            while not self._stop_evt.is_set():
                mhz = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
                try:
                    rs = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    rs = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.samples.append((time.perf_counter(), mhz, rs))
                time.sleep(0.002)
        except Exception as e:  # NVML missing: report, do not fail the bench
            self.error = repr(e)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)

    def summary(self, t0, t1):
        inside = [s for s in self.samples if t0 <= s[0] <= t1]
        window = "timed region"
        if not inside:                      # region shorter than one NVML poll: use every sample taken under load
            inside, window = self.samples, "whole bench (timed region shorter than one NVML poll)"
        if not inside:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "error": self.error or "no samples"}
        mhz = sorted(s[1] for s in inside)
        bits = 0
        for s in inside:
            bits |= s[2]
        return {"sm_mhz": mhz[len(mhz) // 2], "sm_max_mhz": self.max_mhz, "samples": len(inside), "window": window,
                "reasons": sorted(n for b, n in self.REASONS.items() if bits & b)}


def make_inputs(torch, w, device, seed, kind="randn", dtype="f32"):
    """every head map torch.randn from a seeded generator on the owning device.  dtype "bf16": the maps as
    a bf16-emitting head would hand them over; the decode widens them exactly and computes in fp32.
    kind "trained": the heat-maps as randn*3 - 6 (sparse, "trained-like": the shift variant alternates it with randn).
    Workloads with `reg` (cfg3): third return value = (regression map [B,reg,H,W], camera matrices [B,9], dim_ref [C,3])."""
    g = torch.Generator(device=device).manual_seed(seed)
    B, C, H, W = w["B"], w["C"], w["H"], w["W"]
    td = torch.bfloat16 if dtype == "bf16" else torch.float32
    raw = lambda c: torch.randn((B, c, H, W), generator=g, device=device, dtype=torch.float32)
    heat = (lambda c: raw(c) * 3 - 6) if kind == "trained" else raw
    logits = [heat(C).to(td), raw(16).to(td), raw(2).to(td), raw(2).to(td)]
    kpt = heat(w["kpt"]).to(td) if w["kpt"] else None
    box = None
    if w.get("reg"):
        cam = torch.tensor([721.54 / 4, 0, 609.56 / 4, 0, 721.54 / 4, 172.85 / 4, 0, 0, 1], dtype=torch.float32, device=device)
        dim_ref = torch.tensor([[1.53, 1.63, 3.88], [1.76, 0.66, 0.84], [1.74, 0.60, 1.76]], dtype=torch.float32, device=device)[:C]
        box = (raw(w["reg"]).to(td), cam.repeat(B, 1).contiguous(), dim_ref.contiguous())
    return logits, kpt, box


def cpu_reference_step(torch, decode_ref, logits, kpt, K):
    """The reference's eval forward after the heads: clone the maps (models/model.py:27), decode (:29-75 + Tier B wiring)."""
    cl = [p.clone() for p in logits]
    return decode_ref.decode(cl, THRESH, K, DOWN, None if kpt is None else kpt.clone())


def time_cpu_reference(torch, w, steps, warmup, seed=1234):
    """images/s of the oracle's torch port on the host cores, on a bounded sample of the workload."""
    from oracle import decode_ref  # bench.py's cpu_baseline / reference arm: the one place the product side may run oracle/
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    ws = dict(w, B=min(w["B"], CPU_SAMPLE_IMAGES))
    logits, kpt, _ = make_inputs(torch, ws, torch.device("cpu"), seed)
    with torch.no_grad():
        for _ in range(warmup):
            cpu_reference_step(torch, decode_ref, logits, kpt, w["K"])
        t0 = time.perf_counter()
        for _ in range(steps):
            cpu_reference_step(torch, decode_ref, logits, kpt, w["K"])
        dt = time.perf_counter() - t0
    return ws["B"] * steps / dt, dt / steps * 1e3, cores, torch.get_num_threads(), ws["B"]


def config_dict(name, w, n_gpus, scaling, images_per_gpu):
    """The workload, identical in both arms (the driver compares the two dicts)."""
    return {"workload": f"{name}: {w['note']}", "images_per_gpu": images_per_gpu, "global_images": images_per_gpu * n_gpus,
            "heatmap": [w["C"], w["H"], w["W"]], "kpt_channels": w["kpt"], "topk": w["K"], "score_thresh": THRESH,
            "down_sample": DOWN, "parallelism": f"image-sharded x{n_gpus} ({scaling} scaling)"}


def images_per_gpu(w, n_gpus, scaling):
    if scaling == "strong":
        if w["B"] % n_gpus:
            raise SystemExit(f"bench.py: --scaling strong needs the workload's {w['B']} images to divide by {n_gpus} GPUs")
        return w["B"] // n_gpus
    return w["B"]


def run_reference(args, w):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import torch
    ips, ms, cores, threads, nb = time_cpu_reference(torch, w, args.steps, args.warmup)
    line = {"impl": "reference", "metric": METRIC, "value": round(ips, 2), "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms, 3), "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": config_dict(args.workload, w, args.gpus, args.scaling, images_per_gpu(w, args.gpus, args.scaling)),
            "cpu_baseline": {"value": round(ips, 2), "unit": UNIT, "cores": cores, "kind": "port", "torch_threads": threads,
                             "sample": f"{nb} images of the workload per step (torch port of Model.inference incl. the clone of "
                                       f"models/model.py:27 and the keypoint branch when the workload has one), {args.steps} steps"},
            "e2e": {"value": round(ips, 2), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


def run_b200(args, w_job):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (rtm3d_b200 has no CPU path; use --impl reference for the CPU arm)")
    if world != args.gpus:
        raise SystemExit(f"bench.py: --gpus {args.gpus} but WORLD_SIZE={world}; launch N>1 with torch.distributed.run")
    B = images_per_gpu(w_job, world, args.scaling)
    w = dict(w_job, B=B)

    # ---- CPU baseline first (N = 1 only): before any CUDA / NCCL work competes for the host cores
    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        ips, ms, cores, threads, nb = time_cpu_reference(torch, w_job, steps=3, warmup=1)
        cpu_baseline = {"value": round(ips, 2), "unit": UNIT, "cores": cores, "kind": "port", "torch_threads": threads,
                        "sample": f"{nb} images of the workload, 1 warm-up + 3 timed passes of the torch port of "
                                  "Model.inference (incl. the clone of models/model.py:27 and the keypoint branch)"}

    import torch.distributed as dist
    from rtm3d_b200 import HeatmapDecoder, HostDecodeSession, _native
    from rtm3d_b200.decoder import PackedDetections
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    saved_stdout = None
    if world > 1:
        # NCCL writes its version banner to stdout on the first communicator: keep stdout for the ONE JSON line
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        # the gather's NCCL kernel shares the GPU with the next batch's decode: keep it to the few SMs the scan kernel leaves
        # free (--max-ctas), unless the environment says otherwise
        if args.nccl_ctas > 0:
            os.environ.setdefault("NCCL_MAX_CTAS", str(args.nccl_ctas))
            os.environ.setdefault("NCCL_MIN_CTAS", "1")
        dist.init_process_group("nccl", device_id=dev)

    K, Cv = w["K"], w["kpt"]

    if args.gather == "auto":
        # measured at 8 GPUs: rows stored by the kernel itself are a 20 MB burst at its end (+27 us on a 127 us step);
        # moved by the copy engines they overlap the next batch completely, at the price of n-1 copy calls per step on the host --
        # which smaller batches cannot hide (strong scaling, measured: 128 images per rank 106 us with the copies vs 84 us with the
        # stores; 32 images per rank 128 us vs 68 us)
        args.gather = "ce" if B * K * PackedDetections.WORDS * 4 >= (2 << 20) else "p2p"

    def setup_gather():
        """Gather buffers of the two result slots.  Preferred: symmetric memory (every rank's buffer mapped into every process)
        -- the select + post kernel then stores the wire rows straight into all ranks' buffers (rtm3d_decode_fused_gather) and
        the only other cross-GPU operation of a step is a barrier on the gather stream.  Else: rtm3d_pack_wire + NCCL all-gather."""
        per = K * PackedDetections.WORDS + 1
        go = dict(stream=torch.cuda.Stream(device=dev), gathered=[None, None], mode="nccl", peers=[None, None], hdl=None, n_slots=2)
        if args.gather in ("p2p", "p2pd", "ce") and Cv and not w.get("reg"):
            try:
                import ctypes
                import torch.distributed._symmetric_memory as symm_mem
                # deferred pushes (p2pd): a batch's rows travel during the NEXT step, its flags arrive a step later -> one more slot
                n_slots = 3 if args.gather == "p2pd" else 2
                slot_words = (world * B * per + world + 3) // 4 * 4     # rows of every rank + one arrival flag per source rank (16-byte slots)
                buf = symm_mem.empty((n_slots, slot_words), dtype=torch.int32, device=dev)
                hdl = symm_mem.rendezvous(buf, dist.group.WORLD.group_name)
                go["peers"] = [(ctypes.c_void_p * world)(*[int(ptr) + slot * slot_words * 4 for ptr in hdl.buffer_ptrs]) for slot in range(n_slots)]
                buf.zero_()
                go.update(mode=args.gather, hdl=hdl, buf=buf, full=[buf[s_][:world * B * per].view(world * B, per) for s_ in range(n_slots)],
                          mine=[None] * n_slots, step_id=0, n_slots=n_slots, gathered=[None] * n_slots)
                if args.gather == "ce":
                    # copy engines: this rank's rows of a slot, as seen in every rank's buffer (peer-mapped views of the same offsets)
                    go["wait_stream"] = torch.cuda.Stream(device=dev)
                    go["peer_rows"] = [[hdl.get_buffer(r, (n_slots, slot_words), torch.int32)[s_][rank * B * per:(rank + 1) * B * per]
                                        for r in range(world)] for s_ in range(n_slots)]
                torch.cuda.synchronize()
                dist.barrier()
                return go
            except Exception as e:
                print(f"bench.py: symmetric-memory gather not available ({e!r}); NCCL all-gather", file=sys.stderr)
        go["full"] = [torch.empty((world * B, per), dtype=torch.int32, device=dev) for _ in range(2)]
        go["mine"] = [torch.empty((B, per), dtype=torch.int32, device=dev) for _ in range(2)]
        return go

    gather_out = setup_gather() if world > 1 else None
    # N > 1 with the NCCL all-gather: its kernel runs beside the next batch's decode; the persistent scan kernel leaves it a
    # few SMs instead of queueing its last CTAs behind it (--max-ctas; 0 = one CTA per SM).  The fused peer-to-peer gather
    # has no second kernel to make room for.
    max_ctas = args.max_ctas if args.max_ctas >= 0 else (0 if (world == 1 or gather_out["mode"] != "nccl") else 144)
    # result buffers are reused from call to call (saves ~35 us of host time per step); at N > 1 two decoders alternate so
    # that a batch's results stay untouched while the gather stream packs them
    mk_dec = lambda: HeatmapDecoder(THRESH, K, DOWN, max_ctas=max_ctas, reuse_outputs=True)
    decs_dev = [mk_dec() for _ in range(1 if world == 1 else 2)]
    nsets = 2                                 # (== the two result slots: step i uses input set, decoder and slot i & 1)
    sets = [make_inputs(torch, w, dev, 1234 + rank + 100 * s, dtype=args.dtype) for s in range(nsets)]
    elem = 2 if args.dtype == "bf16" else 4
    heat_bytes = B * (w["C"] + Cv) * w["H"] * w["W"] * elem
    if w.get("reg"):
        names = ["scan+select(main)", "box3d"]
    elif Cv:
        names = ["scan_planes(main+kpt)", "select_post"]
        probe = []
        decs_dev[0].decode_with_keypoints(sets[0][0], sets[0][1], marks=probe)      # which kernels serve this shape
        torch.cuda.synchronize()
        if decs_dev[0].staged_path == "legacy":                   # planes larger than the scan kernel's ring (cfg5): round-1 kernels
            names = ["decode_planes(main+kpt; round-1 streaming kernel)", "post_fused"]
    else:
        names = ["scan+select+epilogue(main)"]
    launches_per_step = {"scan_planes(main+kpt)": 1, "select_post": 1, "decode_planes(main+kpt; round-1 streaming kernel)": 1, "post_fused": 1, "scan+select(main)": 2, "box3d": 1, "scan+select+epilogue(main)": 3}
    n_launches = sum(launches_per_step[n] for n in names)       # (+ the wire-packing kernel when the gather goes through NCCL: added below)

    capturing = False

    def decode(i, dec, inputs, marks=None, gather=None):
        logits, kpt, box = inputs

        def mark():
            if marks is not None:
                e = torch.cuda.Event(enable_timing=True)
                e.record()
                marks.append(e)
        if box is not None:
            mark()
            det = dec.select_main(logits[0])
            mark()
            dec.decode_box3d(det, box[0], box[1], box[2], w["C"], cached=True)
            mark()
            return det
        if Cv:
            det, _, _ = dec.decode_with_keypoints(logits, kpt, marks=marks, gather=gather)
            return det
        mark()
        det = dec.decode_packed(logits)
        mark()
        return det

    def step(i, marks=None, inputs=None, graph=None):
        """One step.  N > 1: the decode on the launching stream (from a CUDA graph when given) and the path's one exchange
        (SURVEY.md 8e), the gather of the fixed-size detections: copy engines behind the kernel (ce), fused into the select +
        post kernel (p2p / p2pd: peer-to-peer stores, arrival flags awaited on a second stream), or rtm3d_pack_wire + NCCL
        all-gather on the second stream -- always behind an event, so that the exchange of batch i overlaps the decode of batch
        i+1; the timed region ends behind the last exchange."""
        p2p = world > 1 and gather_out["mode"] in ("p2p", "p2pd", "ce")
        deferred = world > 1 and gather_out["mode"] == "p2pd"
        ce = world > 1 and gather_out["mode"] == "ce"
        n_slots = gather_out["n_slots"] if world > 1 else 2
        # p2p: the slot follows the batch counter (the id carried by the arrival flags); NCCL: the step index
        slot = (gather_out["step_id"] % n_slots) if (p2p and marks is None) else (i & 1)
        if world > 1 and not capturing and gather_out["gathered"][slot] is not None:
            torch.cuda.current_stream().wait_event(gather_out["gathered"][slot])    # the batch that used this slot last has been exchanged
        wait_slot, wait_id = slot, 0
        if graph is not None:
            graph.replay()
            det = None
        else:
            gt = None
            if p2p and marks is None:
                gather_out["step_id"] += 1
                sid = gather_out["step_id"]
                if deferred:
                    prev_slot = (sid - 2) % n_slots
                    gt = (gather_out["peers"][slot], gather_out["peers"][prev_slot] if sid > 1 else None, world, rank, sid - 1)
                    wait_slot, wait_id = prev_slot, sid - 1                          # (what this launch delivers: the previous batch)
                elif ce:
                    gt = (gather_out["peers"][slot], None, world, rank, 0)              # the rows stay in this rank's buffer
                    wait_id = sid
                else:
                    gt = (gather_out["peers"][slot], world, rank, sid)
                    wait_id = sid
                gather_out["last"] = (slot, sid)
            det = decode(i, decs_dev[i % len(decs_dev)], inputs if inputs is not None else sets[i % nsets], marks, gather=gt)
            if world > 1 and not p2p:
                det.to_wire(gather_out["mine"][slot])                                 # one launch of the library (rtm3d_pack_wire)
        if world > 1 and not capturing:
            ready = torch.cuda.Event()
            ready.record()
            with torch.cuda.stream(gather_out["stream"]):
                gather_out["stream"].wait_event(ready)
                if ce and marks is None and wait_id > 0:
                    # the copy engines carry the rows to the other ranks (7 cudaMemcpyAsync over NVLink, no SM involved), the flag
                    # follows in stream order; the wait for everybody's flags sits on a third stream
                    mine_rows = gather_out["peer_rows"][slot][rank]
                    for r in range(world):
                        if r != rank:
                            gather_out["peer_rows"][slot][r].copy_(mine_rows, non_blocking=True)
                    _native.check(_native.lib().rtm3d_signal_gather(gather_out["peers"][slot], world, rank, B, K, 8, wait_id,
                                                                    gather_out["stream"].cuda_stream), "rtm3d_signal_gather")
                    with torch.cuda.stream(gather_out["wait_stream"]):
                        _native.check(_native.lib().rtm3d_wait_gather(gather_out["buf"][wait_slot].data_ptr(), B, K, 8, world, wait_id,
                                                                      gather_out["wait_stream"].cuda_stream), "rtm3d_wait_gather")
                        done = torch.cuda.Event()
                        done.record()
                    gather_out["stream"].wait_event(done)
                elif p2p:
                    if marks is None and wait_id > 0:                                # every rank's rows of that batch have landed here
                        _native.check(_native.lib().rtm3d_wait_gather(gather_out["buf"][wait_slot].data_ptr(), B, K, 8, world, wait_id,
                                                                      gather_out["stream"].cuda_stream), "rtm3d_wait_gather")
                else:
                    dist.all_gather_into_tensor(gather_out["full"][slot], gather_out["mine"][slot])
                if gather_out["gathered"][wait_slot] is None:
                    gather_out["gathered"][wait_slot] = torch.cuda.Event()
                gather_out["gathered"][wait_slot].record()
        return det

    def flush_gather():
        """Deferred peer-to-peer gather: the last batch's rows are still on their rank -- push them (rtm3d_push_gather) and wait
        for every rank's.  Part of the timed region."""
        if world > 1 and gather_out["mode"] == "p2pd" and gather_out.get("last"):
            slot, sid = gather_out["last"]
            st = torch.cuda.current_stream()
            _native.check(_native.lib().rtm3d_push_gather(gather_out["peers"][slot], world, rank, B, K, 8, sid, st.cuda_stream), "rtm3d_push_gather")
            _native.check(_native.lib().rtm3d_wait_gather(gather_out["buf"][slot].data_ptr(), B, K, 8, world, sid, st.cuda_stream), "rtm3d_wait_gather")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local, uuid="GPU-" + str(torch.cuda.get_device_properties(dev).uuid))
    sampler.start()

    for i in range(args.warmup):
        step(i)
    barrier()

    # ---- the kernels of a step replayed from CUDA graphs (one per input set / result slot): N > 1 by default -- the host
    #      would otherwise need longer to issue a step (kernels, events, pack, all-gather call) than the GPU to run it -- and
    #      on request at N = 1 (launch-bound small batches).  The collective itself is issued eagerly on the gather stream.
    graphs = None
    # (the fused peer-to-peer gather carries a step id per launch: eager steps; its host side is light enough)
    use_graph = args.graph or (world > 1 and not args.no_graph and gather_out["mode"] == "nccl")
    if use_graph:
        try:
            graphs = []
            for j in range(2):
                gph = torch.cuda.CUDAGraph()
                capturing = True
                with torch.cuda.graph(gph):
                    step(j)
                capturing = False
                graphs.append(gph)
            for j in range(2):
                step(j, graph=graphs[j])
            barrier()
        except Exception as e:                       # capture not possible: eager steps
            capturing = False
            graphs = None
            print(f"bench.py: CUDA graph capture failed ({e!r}); eager steps", file=sys.stderr)
            torch.cuda.synchronize()

    # ---- timed region: exactly K steps, CUDA events on the launching stream, per-kernel marks on the same stream
    p2p_mode = world > 1 and gather_out["mode"] in ("p2p", "p2pd", "ce")
    marks = [] if (graphs is None and not p2p_mode) else None      # (p2p: ONE fused call per step; kernel times from a separate short run)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.perf_counter()
    e0.record()
    for i in range(args.steps):
        if graphs is None:
            step(i, marks)
        else:
            step(i, graph=graphs[i & 1])
    flush_gather()
    if world > 1:                                     # the timed region ends with every exchange complete
        for ev in gather_out["gathered"]:
            if ev is not None:
                torch.cuda.current_stream().wait_event(ev)
    e1.record()
    barrier()
    t_wall1 = time.perf_counter()
    ms_total = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / args.steps

    # per-kernel durations from the marks (each step: one event before every stage and one after the last)
    kernel_ms = {}
    if marks is not None:
        per = len(names) + 1
        kernel_ms = {n: 0.0 for n in names}
        for s in range(args.steps):
            for j, n in enumerate(names):
                kernel_ms[n] += marks[s * per + j].elapsed_time(marks[s * per + j + 1])
        kernel_ms = {n: v / args.steps for n, v in kernel_ms.items()}
    else:                                   # graph replay: per-kernel times from a short un-captured run afterwards
        m2 = []
        for i in range(8):
            step(i, m2)
        barrier()
        per = len(names) + 1
        kernel_ms = {n: sum(m2[s * per + j].elapsed_time(m2[s * per + j + 1]) for s in range(8)) / 8 for j, n in enumerate(names)}

    # ---- the gather delivers every rank's detections to every rank: check the last two batches on this rank
    verify = None
    if world > 1 and args.verify:
        barrier()
        ok = True
        for slot in range(2):
            # two more (eager) steps: this rank's own wire rows, packed by rtm3d_pack_wire, are the reference for its block
            det = step(slot)
            flush_gather()
            barrier()
            full = gather_out["full"][gather_out["last"][0] if p2p_mode else slot]
            mine = det.to_wire()
            ok &= bool(torch.equal(full[rank * B:(rank + 1) * B], mine))                  # my block is my own wire rows
            sums = full.reshape(world, -1).to(torch.int64).sum(dim=1)                      # checksum of every rank's block as I received it
            own = torch.zeros(world, dtype=torch.int64, device=dev)
            own[rank] = mine.to(torch.int64).sum()
            dist.all_reduce(own, op=dist.ReduceOp.SUM)                                     # the checksums the owners computed
            ok &= bool(torch.equal(sums, own))
            # ... and the rows decode back into detections with plausible counts
            back = PackedDetections.from_wire(full, K)
            ok &= bool(((back.counts >= 0) & (back.counts <= K)).all())
        flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        verify = {"ok": bool(flag.item()), "ranks": world, "rows_per_rank": B, "gather": gather_out["mode"],
                  "what": "gathered wire rows == every owner's rows (own block bit-equal with rtm3d_pack_wire's, all blocks by checksum), both result slots"}

    # ---- the kernels keep no state: first launch on a fresh workspace, and batches that alternate between two distributions
    cold_ms = shift_ms = None
    if world == 1 and not args.no_extras:
        # cold: every workspace of the decoder re-initialised (rtm3d_workspace_init: zeroed + tables), then ONE step
        from rtm3d_b200 import _native
        for wsb in decs_dev[0]._ws.values():
            _native.check(_native.lib().rtm3d_workspace_init(wsb.data_ptr(), wsb.numel(), torch.cuda.current_stream(dev).cuda_stream), "workspace_init")
        torch.cuda.synchronize()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        decode(0, decs_dev[0], sets[1])
        c1.record()
        torch.cuda.synchronize()
        cold_ms = c0.elapsed_time(c1)
        shifted = [sets[0], make_inputs(torch, w, dev, 999, kind="trained", dtype=args.dtype)]
        for i in range(4):
            step(i, inputs=shifted[i % 2])
        torch.cuda.synchronize()
        n_shift = max(10, min(args.steps, 50))
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for i in range(n_shift):
            step(i, inputs=shifted[i % 2])
        s1.record()
        torch.cuda.synchronize()
        shift_ms = s0.elapsed_time(s1) / n_shift

    # ---- end to end: pinned host buffers in, pinned host buffers out, through the host-buffer C-ABI entry points
    # Two sessions (each with its own decoder workspace, device staging and pinned result buffers) alternate on two
    # streams: the zero-copy gathers and the D2H of step i overlap the H2D of step i+1.  Every step's H2D, kernels and
    # D2H are inside the timed region and every step's result is read on the host.
    e2e = None
    if not args.no_e2e and not w.get("reg"):
        host_logits = [t.to("cpu").pin_memory() for t in sets[0][0]]
        host_kpt = sets[0][1].to("cpu").pin_memory() if Cv else None
        depth = 2
        decs = [HeatmapDecoder(THRESH, K, DOWN) for _ in range(depth)]
        sess = [HostDecodeSession(d, B, w["C"], w["H"], w["W"], n_vert=8, kpt_channels=Cv, dtype=sets[0][0][0].dtype, device=dev) for d in decs]
        streams = [torch.cuda.Stream(device=dev) for _ in range(depth)]
        e2e_steps = max(4, min(args.steps, 20))

        def e2e_run(n):
            seen = 0
            for i in range(n):
                j = i % depth
                if i >= depth:
                    streams[j].synchronize()                      # step i - depth is complete: read its result
                    seen += int(sess[j].det_host.counts[0])
                with torch.cuda.stream(streams[j]):
                    sess[j].run(host_logits, host_kpt, sync=False)
            for j in range(depth):
                streams[j].synchronize()
                seen += int(sess[j].det_host.counts[0])
            return seen

        e2e_run(2 * depth)
        barrier()
        t0 = time.perf_counter()
        e2e_run(e2e_steps)
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_s = float(t.item())
        e2e = {"value": round(B * world * e2e_steps / e2e_s, 1), "unit": UNIT, "h2d_bytes_per_step": sess[0].h2d_bytes(),
               "d2h_bytes_per_step": sess[0].d2h_bytes(), "steps": e2e_steps, "sessions_in_flight": depth,
               "api": "HostDecodeSession.run -> " + ("rtm3d_decode_fused_host" if Cv else "rtm3d_decode_main_host")}
    sampler.stop()

    if rank == 0:
        peak, peak_src = peaks()
        A, _, _ = algorithmic_bytes_per_image(w, elem=elem)
        kb = kernel_bytes_per_image(w, elem=elem)
        own = {}
        for j, n in enumerate(names):
            own[n] = B * (kb["scan"] if j == 0 else kb["post"]) if len(names) > 1 else B * A
        per_kernel = {n: {"ms": round(kernel_ms[n], 5), "bytes": own[n],
                          "gbs": round(own[n] / (kernel_ms[n] * 1e-3) / 1e9, 1) if kernel_ms[n] > 0 else 0.0,
                          "frac": round(own[n] / (kernel_ms[n] * 1e-3) / 1e9 / peak, 4) if kernel_ms[n] > 0 else 0.0} for n in names}
        dom = max(kernel_ms, key=kernel_ms.get)
        step_gbs = B * A / (ms_step * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": round(B * world / (ms_step * 1e-3), 1), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_step, 5), "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f32" if elem == 4 else "f32 (bf16 maps widened on load)", "data": "synthetic",
            "config": config_dict(args.workload, w_job, world, args.scaling, B),
            "inputs": f"{nsets} input sets rotated, {heat_bytes / 1e6:.0f} MB of heat-map per step vs 126 MB L2"
                      + ("" if heat_bytes > 130e6 else " (SMALLER than L2: later steps may hit L2)"),
            "run": {"cuda_graph": graphs is not None, "max_ctas": max_ctas, "gather": None if world == 1 else gather_out["mode"]},
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": per_kernel[dom]["gbs"], "peak": peak, "unit": "GB/s",
                         "frac": per_kernel[dom]["frac"], "traffic": ncu_traffic(args.workload), "peak_source": peak_src,
                         "bytes_per_launch": per_kernel[dom]["bytes"],
                         "bytes_note": "the bytes the kernel itself moves: the scan kernel streams the heat-maps once; gathers (32 B per "
                                       "scalar) and outputs belong to the kernels behind it",
                         "kernel_ms": {k: round(v, 5) for k, v in kernel_ms.items()}, "kernels": per_kernel,
                         "step_achieved": round(step_gbs, 1), "step_frac": round(step_gbs / peak, 4),
                         "algorithmic_bytes_per_image": A, "cold_ms": None if cold_ms is None else round(cold_ms, 5),
                         "shift_ms_per_step": None if shift_ms is None else round(shift_ms, 5),
                         "frac_of_nominal_8tbs": round(per_kernel[dom]["gbs"] / 8000.0, 4),
                         # every head map counted in full (the literal reading of north_star) -- NOT the headline:
                         # nobody streams the regression planes, the decode gathers K points from them
                         "literal_full_maps": {"gbs": round(B * (w["C"] + Cv + (w["reg"] if w.get("reg") else 20)) * w["H"] * w["W"] * elem
                                                            / (ms_step * 1e-3) / 1e9, 1), "note": "all head maps in full / ms_per_step; not the headline"},
                         "state": ("round-1 streaming kernels (planes larger than the scan kernel's ring): thresholds remembered per workspace"
                                   if names[0].startswith("decode_planes") else
                                   "none: thresholds come from each strip's own data (no memory across planes or launches)")},
            "e2e": e2e,
            # (+ per step at N > 1: the wire-packing kernel (nccl), the flag-wait kernel (p2p, p2pd), the signal and wait kernels (ce))
            "gpu_launches": (n_launches + (0 if world == 1 else {"nccl": 1, "p2p": 1, "p2pd": 1, "ce": 2}[gather_out["mode"]])) * args.steps,
            "clocks": sampler.summary(t_wall0, t_wall1),
        }
        if cpu_baseline is not None:
            line["cpu_baseline"] = cpu_baseline
        if verify is not None:
            line["verify"] = verify
        if saved_stdout is not None:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--workload", choices=sorted(WORKLOADS), default="cfg4")
    ap.add_argument("--scaling", choices=["weak", "strong"], default="weak", help="weak: the workload's batch per GPU; strong: cut into N shards")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--dtype", choices=["f32", "bf16"], default="f32", help="element type of the head maps handed to the decode")
    ap.add_argument("--no-e2e", action="store_true", help="device-resident timed loop only (the ncu passes)")
    ap.add_argument("--no-extras", action="store_true", help="skip the cold-launch and distribution-shift measurements")
    ap.add_argument("--graph", action="store_true", help="replay the steps from a CUDA graph (default at N > 1; launch-bound small batches at N = 1)")
    ap.add_argument("--no-graph", action="store_true", help="N > 1: eager steps instead of the graph replay")
    ap.add_argument("--verify", action="store_true", help="N > 1: check the all-gathered detections against every owner's rows")
    ap.add_argument("--gather", choices=["auto", "p2p", "p2pd", "ce", "nccl"], default="auto",
                    help="N > 1: auto = ce when a rank's rows are >= 2 MB per batch, else p2p.  Fused peer-to-peer gather (symmetric memory) -- p2pd: a batch's rows are pushed while the next batch is sorted, "
                         "p2p: stored at the end of their own launch; ce: rows left on their rank by the kernel and moved by the copy engines "
                         "on a second stream -- or pack + NCCL all-gather")
    ap.add_argument("--nccl-ctas", type=int, default=0, help="N > 1: cap the CTAs of NCCL's kernels (NCCL_MAX_CTAS; 0 = NCCL's choice)")
    ap.add_argument("--max-ctas", type=int, default=-1, help="CTAs of the scan kernel (-1: all SMs at N=1, 144 at N>1)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        return run_reference(args, w)
    return run_b200(args, w)


if __name__ == "__main__":
    sys.exit(main())

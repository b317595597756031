#!/usr/bin/env python
"""bench.py - headline benchmark of the mxdetection detection hot path on B200.

Metric (BASELINE.json): RoIAlign RoIs/s fwd+bwd, with proposals/s, assigner anchors/s and the HBM
roofline fraction beside it.  Workload at N GPUs (weak scaling): every rank owns 8 images of BASELINE
config 3 / 5 - the Faster R-CNN R50-FPN RoI stage, 512 RoIs/img x 8 imgs on 4 FPN maps of an
800x1344 image, 256 ch, 7x7, sample_ratio 2 - whose 731 MB of feature maps exceed the 126 MB L2, so
every step streams from HBM.  A step = one forward + one backward of that shard.

  python bench.py --gpus N --steps K --warmup W           (N>1: launched under torch.distributed.run)
  python bench.py --impl reference ...                    (the CPU port of the reference path, timed)

Prints ONE JSON line (rank 0).  Timing: CUDA events on the launching stream, barrier +
synchronize on both sides of the K timed steps, max over ranks.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

F = np.float32
METRIC = "roialign_fpn_fwd_bwd_rois_per_s"
UNIT = "RoIs/s"
IMGS_PER_GPU = 8
ROIS_PER_IMG = 512
POOLED = (7, 7)
WORKLOAD = ("BASELINE config 3 (= per-GPU shard of config 5): FPN RoI stage, 8 imgs x 512 RoIs, 800x1344 "
            "(1333x800 padded), 4 levels x 256 ch fp32 NCHW, 7x7, sample_ratio 2, level assignment in-kernel")


def config_dict(n_gpus):
    return {"workload": WORKLOAD, "images_per_gpu": IMGS_PER_GPU, "rois_per_image": ROIS_PER_IMG,
            "global_images": IMGS_PER_GPU * n_gpus, "parallelism": "dp%d (images sharded, no data-path collective)" % n_gpus,
            "l2_policy": "inputs larger than L2 (731 MB maps + 206 MB grad_out per step vs 126 MB L2)"}


# ------------------------------------------------------------------------ helpers --
def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


CURRENT_NCU_SUMMARY = "ncu_r2f.txt"     # `ncu --set full` capture of the kernels this tree ships (profiles/README.md)


def bind_to_gpu_numa_node(index):
    """Multi-rank runs: keep this rank's threads (and therefore its first-touch pinned host buffers of the e2e leg)
    on the NUMA node the GPU hangs off, so eight ranks do not pull their PCIe traffic across the socket link.
    Best effort - silently skipped when sysfs does not say."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(index)
        bus = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bus).read())
        if node < 0:
            return
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            sys.stderr.write("[bench] gpu %d (%s): bound to NUMA node %d, %d cpus\n" % (index, bus, node, len(cpus)))
    except Exception as e:
        sys.stderr.write("[bench] gpu %d: NUMA binding skipped (%r)\n" % (index, e))


def ncu_traffic(kernel_substr):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the newest committed
    `ncu --set full` summary under profiles/ (profiles/summarize.py writes them); None when there is none."""
    import glob
    import re
    best = None
    paths = sorted(glob.glob(os.path.join(ROOT, "profiles", "ncu_*.txt")))
    cur = os.path.join(ROOT, "profiles", CURRENT_NCU_SUMMARY)
    for path in [q for q in paths if q != cur] + ([cur] if os.path.exists(cur) else []):   # the current build's capture wins
        txt = open(path).read()
        for blk in txt.split("== ncu --set full:")[1:]:
            if kernel_substr not in blk.splitlines()[0]:
                continue
            tot = 0.0
            for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                m = re.search(key + r"\s+([0-9.]+)\s+(\w+)", blk)
                if not m:
                    tot = None
                    break
                tot += float(m.group(1)) * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(m.group(2), 1.0)
            if tot:
                best = (tot, os.path.relpath(path, ROOT))
    return best


class ClockSampler:
    """SM clock / throttle-reason sampling DURING the timed region (NVML polled from a thread every ~2 ms;
    same fields as the nvidia-smi recipe of B200_PROFILING.md, which is too slow to start for a 50 ms region)."""
    BITS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
            0x80: "hw_power_brake_slowdown"}

    def __init__(self, torch_device_index):
        import threading
        self.samples, self.reasons, self.power = [], set(), []
        self.stop_flag = threading.Event()
        self.thread = None
        self.max_mhz = None
        self.err = None
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            self.nv = pynvml
            try:
                uuid = str(torch.cuda.get_device_properties(torch_device_index).uuid)
                self.h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode())
            except Exception:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(torch_device_index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._run, daemon=True)
        except Exception as e:  # no NVML: report it instead of inventing clocks
            self.err = repr(e)

    def _run(self):
        nv = self.nv
        while not self.stop_flag.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
                mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for bit, name in self.BITS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception as e:
                self.err = repr(e)
                return
            time.sleep(0.002)

    def start(self):
        if self.thread is not None:
            self.thread.start()

    def stop(self):
        self.stop_flag.set()
        if self.thread is not None:
            self.thread.join(timeout=2)
        out = {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz,
               "power_w_max": max(self.power) if self.power else None, "samples": len(self.samples),
               "reasons": sorted(self.reasons)}
        if self.err:
            out["error"] = self.err
        return out


def touched_bytes(rois, levels, shapes, scales, nimg, channels, pooled, sr):
    """Unique feature pixels any tap reads (Spec A sample positions), times C*4 bytes - the tighter
    algorithmic-bytes variant of SURVEY.md 8(d).  Accounting only; not on the product path."""
    PH, PW = pooled
    masks = [np.zeros((nimg, h, w), bool) for h, w in shapes]

    def axis(start, binsz, P, size):
        t = np.arange(P * sr)
        c = (start + (t // sr).astype(F) * binsz) + (((t % sr).astype(F) + F(0.5)) * binsz) / F(sr)
        ok = ~((c < -1) | (c > size))
        c = np.maximum(c, 0)
        lo = np.minimum(c.astype(np.int64), size - 1)
        hi = np.minimum(lo + 1, size - 1)
        return np.unique(np.concatenate([lo[ok], hi[ok]]))

    for r, lv in zip(rois, levels):
        b = int(r[0]); sc = F(scales[lv]); H, W = shapes[lv]
        x1, y1, x2, y2 = [F(v) * sc for v in r[1:]]
        bw = max(x2 - x1, F(1)) / F(PW); bh = max(y2 - y1, F(1)) / F(PH)
        rows = axis(y1, bh, PH, H); cols = axis(x1, bw, PW, W)
        if len(rows) and len(cols):
            masks[lv][b][np.ix_(rows, cols)] = True
    return int(sum(int(m.sum()) for m in masks)) * channels * 4


# ------------------------------------------------------------------ reference arm --
def load_synthetic():
    """mxdetection_b200/synthetic.py loaded BY PATH: the reference arm must not import the product package (that
    would map libmxdet_sm100.so into the reference process)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("_mxd_synthetic", os.path.join(ROOT, "mxdetection_b200", "synthetic.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def cpu_reference_fpn(n_images, passes, first_image=0, keep=False):
    """Times the C port of the reference CPU path (oracle/cpu_ref.c: OpenMP-over-RoIs forward, serial
    backward - the parallelisation of mxnet 1.3's roi_align.cc) on `n_images` images of the workload."""
    from oracle import cref
    syn = load_synthetic()
    d = syn.cfg3(batch=n_images, first_image=first_image, with_features=True)
    lv = cref.map_roi_levels(d["rois"], 4)
    shapes = [f.shape for f in d["feats"]]
    times = []
    out = grads = None
    for _ in range(passes):
        t0 = time.perf_counter()
        out = cref.roi_align_forward(d["feats"], d["rois"], POOLED, d["scales"], 2, lv)
        grads = cref.roi_align_backward(d["grad_out"], d["rois"], shapes, POOLED, d["scales"], 2, lv)
        times.append(time.perf_counter() - t0)
    if keep:
        return n_images * ROIS_PER_IMG, times, cref.num_threads(), (d, out, grads)
    return n_images * ROIS_PER_IMG, times, cref.num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    # torch.distributed.run exports OMP_NUM_THREADS=1 to its workers; the reference arm is ONE process that may use
    # every host core, so the OpenMP runtime is told so before it starts (recorded in the line)
    os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
    rois, times, threads = cpu_reference_fpn(1, args.warmup + args.steps)
    times = times[args.warmup:]
    total = sum(times)
    value = rois * len(times) / total
    sample = ("each step = 1 of the 8 images per GPU (512 RoIs, 4 FPN maps x 256 ch): C port of the mxnet-1.3 "
              "CPU ROIAlign (oracle/cpu_ref.c), forward OpenMP over RoIs, backward serial")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": len(times),
            "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config_dict(args.gpus),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "host_cpus": os.cpu_count(), "omp_num_threads": os.environ.get("OMP_NUM_THREADS"),
            "product_library_loaded": "mxdetection_b200" in sys.modules}
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------ our arm --
def run_ours(args):
    import torch
    import torch.distributed as dist
    import mxdetection_b200 as m
    from mxdetection_b200 import synthetic as syn
    from mxdetection_b200.ops import (roi_align_backward, roi_align_forward, roi_align_fpn_backward,
                                      roi_align_fpn_forward)
    from mxdetection_b200.core.anchor import AnchorGenerator, anchor_assign, anchor_inside_flags
    from mxdetection_b200.models.roi_extractors import map_roi_levels
    from mxdetection_b200.models.rpn_heads import ProposalConfig, RPNHead
    from mxdetection_b200.parallel import gather_detections

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device: there is no CPU fallback"
    torch.cuda.set_device(local)
    if world > 1:
        bind_to_gpu_numa_node(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # NCCL prints its version banner to STDOUT when the communicator is created; the contract is ONE json
        # line on stdout, so fd 1 points at stderr until the first collective has run.
        sys.stdout.flush()
        saved = os.dup(1); os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            t0 = torch.zeros(1, device=dev); dist.all_reduce(t0); torch.cuda.synchronize()
        finally:
            sys.stdout.flush(); os.dup2(saved, 1); os.close(saved)
    K, Wm = args.steps, max(args.warmup, 3)
    global IMGS_PER_GPU
    if args.scaling == "strong":
        assert 64 % world == 0, "strong scaling splits the 64 images of BASELINE config 5 evenly"
        IMGS_PER_GPU = 64 // world
    first_image = rank * IMGS_PER_GPU
    peak_gbs, peak_src = measured_peaks()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return float(x)
        t = torch.tensor([float(x)], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def ev():
        return torch.cuda.Event(enable_timing=True)

    # ---------------- workload (synthetic, resident in HBM for `value`) ----------------
    d = syn.cfg3(batch=IMGS_PER_GPU, first_image=first_image, with_features=False)
    shapes = [(IMGS_PER_GPU, 256, h, w) for h, w in d["feat_shapes"]]
    gen = torch.Generator(device=dev); gen.manual_seed(1000 * 3 + rank)
    feats = [torch.randn(s, device=dev, generator=gen) for s in shapes]
    R = d["rois"].shape[0]
    rois = torch.from_numpy(d["rois"]).to(dev)
    grad_out = torch.randn((R, 256) + POOLED, device=dev, generator=gen)
    out = torch.empty((R, 256) + POOLED, device=dev)
    grads = [torch.empty(s, device=dev) for s in shapes]
    scales = d["scales"]

    def fwd():
        roi_align_fpn_forward(feats, rois, POOLED, scales, 2, out=out)

    def bwd():
        roi_align_fpn_backward(grad_out, rois, shapes, POOLED, scales, 2, grad_feats=grads, accumulate=False)

    for _ in range(Wm):
        fwd(); bwd()
    barrier()
    sampler = ClockSampler(local); sampler.start()
    launches0 = m.launch_count()
    marks = [(ev(), ev(), ev()) for _ in range(K)]
    barrier()
    for a, b, c in marks:
        a.record(); fwd(); b.record(); bwd(); c.record()
    barrier()
    launches = m.launch_count() - launches0
    clocks = sampler.stop()
    total_ms = marks[0][0].elapsed_time(marks[-1][2])
    fwd_ms = statistics.mean(a.elapsed_time(b) for a, b, _ in marks)
    bwd_ms = statistics.mean(b.elapsed_time(c) for _, b, c in marks)
    total_ms = max_over_ranks(total_ms)
    value = world * R * K / (total_ms * 1e-3)

    # ---------------- roofline of the dominant kernel ----------------
    levels = map_roi_levels(rois, 4).cpu().numpy()
    map_bytes = sum(int(np.prod(s)) * 4 for s in shapes)
    out_bytes = R * 256 * POOLED[0] * POOLED[1] * 4
    tb = touched_bytes(d["rois"], levels, d["feat_shapes"], scales, IMGS_PER_GPU, 256, POOLED, 2)
    alg_fwd = min(map_bytes, tb) + out_bytes          # read maps (or only the touched pixels) + write out
    alg_bwd = out_bytes + map_bytes                   # read grad_out + write every grad-map byte (req=write)
    dom = "roi_align_backward" if bwd_ms >= fwd_ms else "roi_align_forward"
    dom_ms = max(fwd_ms, bwd_ms); dom_bytes = alg_bwd if bwd_ms >= fwd_ms else alg_fwd
    traffic = ncu_traffic("roi_align_tile_bwd_kernel" if bwd_ms >= fwd_ms else "roi_align_ring_fwd_kernel")
    achieved = dom_bytes / (dom_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak_gbs, "unit": "GB/s",
                "frac": achieved / peak_gbs, "traffic": (traffic[0] if traffic else None),
                "traffic_source": (traffic[1] if traffic else None), "peak_source": peak_src,
                "algorithmic_bytes": dom_bytes, "avg_ms": dom_ms,
                "forward": {"ms": fwd_ms, "bytes": alg_fwd, "gbs": alg_fwd / fwd_ms / 1e6, "frac": alg_fwd / fwd_ms / 1e6 / peak_gbs,
                            "map_bytes": map_bytes, "touched_bytes": tb},
                "backward": {"ms": bwd_ms, "bytes": alg_bwd, "gbs": alg_bwd / bwd_ms / 1e6, "frac": alg_bwd / bwd_ms / 1e6 / peak_gbs,
                             "note": "includes the req=write zero-fill of the grad maps"},
                "fwd_bwd": {"bytes": alg_fwd + alg_bwd, "gbs": (alg_fwd + alg_bwd) / (fwd_ms + bwd_ms) / 1e6,
                            "frac": (alg_fwd + alg_bwd) / (fwd_ms + bwd_ms) / 1e6 / peak_gbs}}

    # ---------------- e2e: host buffers through the public API, copies inside the timed region ----
    e2e = None            # --no-e2e: profiling passes only (the ncu launch list then holds the headline step alone)
    if not args.no_e2e:
        e_steps = K if args.e2e_steps <= 0 else max(1, args.e2e_steps)
        feats_h = [torch.empty(s, pin_memory=True).normal_() for s in shapes]
        gout_h = torch.empty(tuple(grad_out.shape), pin_memory=True).normal_()
        rois_h = torch.from_numpy(d["rois"]).pin_memory()
        out_h = torch.empty(tuple(out.shape), pin_memory=True)
        grads_h = [torch.empty(s, pin_memory=True) for s in shapes]
        h2d = sum(t.numel() * 4 for t in feats_h) + gout_h.numel() * 4 + rois_h.numel() * 4
        d2h = out_h.numel() * 4 + sum(t.numel() * 4 for t in grads_h)

        from mxdetection_b200.ops import HostRoIStage
        stage = HostRoIStage(shapes, ROIS_PER_IMG, POOLED, scales, 2, dev, depth=2)   # whole images: profiles/e2e_sweep.py

        def e2e_step():
            # public host-buffer API: per-image pipeline H2D | fwd+bwd | D2H (every byte still crosses PCIe inside the step)
            stage.forward_backward(feats_h, rois_h, gout_h, out_h, grads_h)

        e2e_step()
        barrier()
        e0, e1 = ev(), ev()
        e0.record()
        for _ in range(e_steps):
            e2e_step()
        e1.record()
        barrier()
        e2e_ms = max_over_ranks(e0.elapsed_time(e1))
        e2e = {"value": world * R * e_steps / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "steps": e_steps, "ms_per_step": e2e_ms / e_steps,
               "api": "mxdetection_b200.ops.HostRoIStage.forward_backward: pinned host buffers in and out, per-image "
                      "H2D | roi_align_fpn_forward/backward (ctypes C ABI) | D2H pipelined over three streams"}
        del feats_h, gout_h, grads_h, out_h

    # ---------------- secondary metrics of BASELINE.json (each: events, max over ranks) ----------
    def timed(fn, iters, warm=3):
        for _ in range(warm):
            fn()
        barrier()
        es = [(ev(), ev()) for _ in range(iters)]
        for a, b in es:
            a.record(); fn(); b.record()
        barrier()
        return max_over_ranks(statistics.mean(a.elapsed_time(b) for a, b in es))

    secondary = {}
    if not args.no_secondary:
        it = args.secondary_iters
        # (a) config 1: 512 RoIs on one 256x200x272 map; 4 rotating buffer sets so the map is never L2-resident
        c1 = syn.cfg1()
        rois1 = torch.from_numpy(c1["rois"]).to(dev)
        maps1 = [torch.randn((1, 256, 200, 272), device=dev, generator=gen) for _ in range(4)]
        g1 = [torch.empty((1, 256, 200, 272), device=dev) for _ in range(4)]
        go1 = [torch.randn((512, 256, 7, 7), device=dev, generator=gen) for _ in range(4)]
        o1 = [torch.empty((512, 256, 7, 7), device=dev) for _ in range(4)]
        ctr = [0]

        def cfg1_step():
            i = ctr[0] % 4; ctr[0] += 1
            roi_align_forward(maps1[i], rois1, (7, 7), 0.25, 2, out=o1[i])
            roi_align_backward(go1[i], rois1, (1, 256, 200, 272), (7, 7), 0.25, 2, grad_data=g1[i])
        ms = timed(cfg1_step, 4 * max(it // 4, 2))
        tb1 = touched_bytes(c1["rois"], np.zeros(512, np.int64), [(200, 272)], [0.25], 1, 256, (7, 7), 2)
        b1 = (min(55705600, tb1) + 25690112) + (25690112 + 55705600)    # fwd: touched pixels (or the map) + out; bwd: grad_out + every map byte
        secondary["cfg1_roialign_fwd_bwd"] = {"rois_per_s": world * 512 / (ms * 1e-3), "ms": ms,
                                              "hbm_frac": b1 / ms / 1e6 / peak_gbs, "algorithmic_bytes": b1, "touched_bytes": tb1,
                                              "l2_policy": "4 rotating buffer sets (445 MB) > L2"}
        del maps1, g1, go1, o1
        # (b) config 2: RPN proposals, 800x1088, 217 413 anchors/img, batch 2 - and the 8-image shard
        head, pcfg = RPNHead(), ProposalConfig(nms_pre=2000, nms_post=1000, max_num=1000, nms_thr=0.7)
        for name, bsz, (ih, iw) in (("cfg2_rpn_proposals_b2", 2, (800, 1088)), ("cfg5_rpn_proposals_b8", 8, (800, 1344))):
            r = syn.rpn_inputs(2, bsz, ih, iw, first_image=first_image)
            sc = [torch.from_numpy(s).to(dev) for s in r["scores"]]; dl = [torch.from_numpy(x).to(dev) for x in r["deltas"]]
            shp = torch.from_numpy(r["img_shapes"]).to(dev)
            nv = [None]

            def prop_step():
                nv[0] = head.get_proposals(sc, dl, r["feat_shapes"], shp, pcfg)
            ms = timed(prop_step, it)
            n_anchor = sum(s.shape[1] for s in sc)
            bytes_ = bsz * (n_anchor * 4 + 8663 * 16 + 1000 * 20)
            secondary[name] = {"images_per_s": world * bsz / (ms * 1e-3), "input_anchors_per_s": world * bsz * n_anchor / (ms * 1e-3),
                               "output_proposals_per_s": world * float(nv[0][1].sum().item()) / (ms * 1e-3), "ms": ms,
                               "hbm_frac": bytes_ / ms / 1e6 / peak_gbs, "bound": "latency (dependent sort/NMS chain), not HBM"}
            if bsz == 8:
                pipe_prop = (sc, dl, r["feat_shapes"], shp)
        # (c) config 4b: max-IoU assigner, 268 569 anchors x 100 GT, 8 images
        a = syn.assigner_inputs(4, IMGS_PER_GPU, first_image=first_image)
        anchors, valid = [], []
        for (fh, fw), s in zip(a["feat_shapes"], a["strides"]):
            ag = AnchorGenerator(s, [8], [0.5, 1.0, 2.0])
            anchors.append(ag.grid_anchors((fh, fw), s, device=dev)); valid.append(ag.valid_flags((fh, fw), (fh, fw), device=dev))
        anchors = torch.cat(anchors); inside = anchor_inside_flags(anchors, torch.cat(valid), a["img_shape"], 0)
        gts = torch.from_numpy(a["gts"]).to(dev); ngt = torch.from_numpy(a["num_gts"]).to(dev); gl = torch.from_numpy(a["gt_labels"]).to(dev)

        def assign_step():
            anchor_assign(anchors, inside, gts, ngt, gl)
        ms = timed(assign_step, it)
        pairs = IMGS_PER_GPU * anchors.shape[0] * 100
        secondary["cfg4b_max_iou_assigner_b8"] = {"anchors_per_s": world * IMGS_PER_GPU * anchors.shape[0] / (ms * 1e-3), "ms": ms,
                                                   "gt_anchor_pairs_per_s": world * pairs / (ms * 1e-3),
                                                   "executed_fma_pipe_note": "ncu counts of the two kernels: profiles/README.md (the kernel skips most pairs; "
                                                                             "no flop-per-pair estimate is reported)",
                                                   "hbm_frac": IMGS_PER_GPU * anchors.shape[0] * 28 / ms / 1e6 / peak_gbs,
                                                   "bound": "latency / issue (bbox-pruned pair tests), not HBM"}
        # (d) config 4a: mask branch, 14x14 on 128 RoIs/img x 8 imgs (same maps)
        dm = syn.cfg4_mask(batch=IMGS_PER_GPU, first_image=first_image, with_features=False)
        rois_m = torch.from_numpy(dm["rois"]).to(dev)
        go_m = torch.randn((rois_m.shape[0], 256, 14, 14), device=dev, generator=gen); o_m = torch.empty_like(go_m)

        def mask_step():
            roi_align_fpn_forward(feats, rois_m, (14, 14), scales, 2, out=o_m)
            roi_align_fpn_backward(go_m, rois_m, shapes, (14, 14), scales, 2, grad_feats=grads)
        ms = timed(mask_step, max(it // 2, 3))
        lv_m = map_roi_levels(rois_m, 4).cpu().numpy()
        tbm = touched_bytes(dm["rois"], lv_m, d["feat_shapes"], scales, IMGS_PER_GPU, 256, (14, 14), 2)
        bm = (min(map_bytes, tbm) + o_m.numel() * 4) + (o_m.numel() * 4 + map_bytes)   # same rule as the headline
        secondary["cfg4a_mask_roialign_14x14_b8"] = {"rois_per_s": world * rois_m.shape[0] / (ms * 1e-3), "ms": ms,
                                                      "hbm_frac": bm / ms / 1e6 / peak_gbs, "algorithmic_bytes": bm,
                                                      "touched_bytes": tbm}
        # (f) SURVEY 8(f) "next" rows at reference sizes: N1 sampling + target packing of one image's anchors, N2 detection
        # post-processing of 1000 proposals x 81 classes
        from mxdetection_b200.core.bbox import MaxIoUAssigner, RandomSampler, pack_targets
        from mxdetection_b200.models.bbox_heads import get_det_bboxes
        rng_n = np.random.default_rng(3 + first_image)
        gts1 = gts[0, : int(a["num_gts"][0])].contiguous()
        asg = MaxIoUAssigner(0.7, 0.3, 0.3).assign(anchors, gts1)
        skeys = torch.rand(anchors.shape[0], device=dev, generator=gen)
        sampler = RandomSampler(256, 0.5, -1)

        def n1_step():
            pack_targets(anchors, asg.gt_inds, gts1, sampler.sample(asg.gt_inds, skeys))
        ms = timed(n1_step, it)
        secondary["n1_sample256_pack_targets"] = {"ms": ms, "anchors": int(anchors.shape[0]), "images_per_s": world / (ms * 1e-3)}
        nd, Cd = 1000, 81
        rois_d = torch.from_numpy(np.concatenate([np.zeros((nd, 1)), syn.gt_boxes(rng_n, 800, 1344, nd)], 1).astype(np.float32)).to(dev)
        lg = rng_n.normal(0, 2, (nd, Cd)); lg[:, 0] += 3
        score_d = torch.from_numpy((np.exp(lg) / np.exp(lg).sum(1, keepdims=True)).astype(np.float32)).to(dev)
        pred_d = torch.from_numpy(rng_n.normal(0, 1.0, (nd, 4 * Cd)).astype(np.float32)).to(dev)

        def n2_step():
            get_det_bboxes(rois_d, score_d, pred_d, (800, 1344), 1.0, 0.05, 0.5, 100)
        ms = timed(n2_step, it)
        secondary["n2_det_bboxes_1000x81"] = {"ms": ms, "images_per_s": world / (ms * 1e-3)}
        # (e) config 5 shard pipeline: assigner + proposals + RoI stage fwd/bwd (+ NCCL gather of detections at N>1)
        side = torch.cuda.Stream(dev)     # the assigner depends on nothing the proposal / RoI chain produces:
                                          # it runs beside the latency-bound top-k / NMS chain

        def pipeline_step():
            cur = torch.cuda.current_stream(dev)
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                anchor_assign(anchors, inside, gts, ngt, gl)
            p, n = head.get_proposals(pipe_prop[0], pipe_prop[1], pipe_prop[2], pipe_prop[3], pcfg)
            fwd(); bwd()
            cur.wait_stream(side)
            gather_detections(p, n, first_image)
        ms = timed(pipeline_step, max(it // 2, 3))
        secondary["cfg5_shard_pipeline"] = {"images_per_s": world * IMGS_PER_GPU / (ms * 1e-3), "ms": ms,
                                            "stages": "assigner (side stream) || rpn proposals + fpn roialign fwd/bwd, then detection all-gather"}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm,
            "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f32", "data": "synthetic (N(0,1) features / grads, COCO-shaped RoIs, seeded per image)",
            "config": config_dict(world), "roofline": roofline, "e2e": e2e, "gpu_launches": int(launches),
            "launches_per_step": launches / K, "clocks": clocks, "secondary": secondary}

    # ---------------- CPU baseline (rank 0, N=1 only): bounded sample of the same workload --------
    if world == 1 and not args.no_cpu_baseline:
        n_rois, times, threads, (dd, ref_out, ref_grads) = cpu_reference_fpn(2, 3, first_image, keep=True)
        t = sum(times[1:]) / len(times[1:])
        line["cpu_baseline"] = {"value": n_rois / t, "unit": UNIT, "cores": threads, "kind": "port",
                                "sample": "2 of the 8 images (1024 RoIs) of the same workload, best-effort 2 timed passes after 1 warm-up: "
                                          "oracle/cpu_ref.c, forward OpenMP over RoIs, backward serial (mxnet 1.3 roi_align.cc structure)",
                                "seconds_per_pass": t, "host_cpus": os.cpu_count()}
        # the oracle as CHECKER: the GPU path on the very inputs the CPU port just processed (1e-5 / 1e-4 bars of north_star)
        f2 = [torch.from_numpy(f).to(dev) for f in dd["feats"]]
        r2 = torch.from_numpy(dd["rois"]).to(dev)
        o2 = roi_align_fpn_forward(f2, r2, POOLED, scales, 2)
        g2 = roi_align_fpn_backward(torch.from_numpy(dd["grad_out"]).to(dev), r2, [tuple(f.shape) for f in dd["feats"]], POOLED, scales, 2)
        o2 = o2.cpu().numpy()
        ok_f = bool(np.all(np.abs(o2 - ref_out) <= 1e-5 * np.maximum(1, np.abs(ref_out))))
        ok_b = all(bool(np.all(np.abs(a.cpu().numpy() - b) <= 1e-4 * np.maximum(1, np.abs(b)))) for a, b in zip(g2, ref_grads))
        line["verified"] = {"forward_1e-5": ok_f, "backward_1e-4": ok_b, "checksum_out": float(np.float64(o2).sum()),
                            "checksum_ref": float(np.float64(ref_out).sum()),
                            "what": "GPU forward + backward of the 2-image sample against the C port's outputs of the cpu_baseline leg"}
        assert ok_f and ok_b, "bench: GPU output does not match the CPU port"
        del f2, g2
        # second, independently compiled CPU baseline: torchvision CPU kernels on BASELINE config 1 and NMS n=2000
        try:
            import torchvision
            c1 = syn.cfg1()
            x = torch.randn(1, 256, 200, 272, requires_grad=True)
            rr = torch.from_numpy(c1["rois"])
            t0 = time.perf_counter(); y = torchvision.ops.roi_align(x, rr, (7, 7), 0.25, 2, aligned=False); t1 = time.perf_counter()
            y.backward(torch.ones_like(y)); t2 = time.perf_counter()
            nb = 2000
            rng_t = np.random.default_rng(7)
            xy = rng_t.uniform(0, 1200, (nb, 2)); wh = rng_t.uniform(8, 200, (nb, 2))
            bx = torch.from_numpy(np.concatenate([xy, xy + wh], 1).astype(np.float32)); scs = torch.from_numpy(rng_t.uniform(0, 1, nb).astype(np.float32))
            t3 = time.perf_counter(); kept = torchvision.ops.nms(bx, scs, 0.7); t4 = time.perf_counter()
            line["cpu_baseline"]["torchvision_cpu"] = {"cfg1_roi_align_fwd_ms": 1e3 * (t1 - t0), "cfg1_roi_align_bwd_ms": 1e3 * (t2 - t1),
                                                       "cfg1_rois_per_s": 512 / (t2 - t0), "nms_2000_ms": 1e3 * (t4 - t3), "nms_kept": int(kept.numel()),
                                                       "threads": torch.get_num_threads(), "version": torchvision.__version__}
        except Exception as e:      # the second baseline is optional
            line["cpu_baseline"]["torchvision_cpu"] = {"unavailable": repr(e)}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=0, help="timed steps of the host-buffer leg (0 = --steps)")
    ap.add_argument("--secondary-iters", type=int, default=20)
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: 8 images per GPU (default); strong: 64 images in total (BASELINE config 5), 64/N per GPU")
    ap.add_argument("--no-secondary", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg (ncu passes of profiles/run_profile.sh)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())

#!/usr/bin/env python
"""bench.py -- distillation-step throughput of the LGD hot path (BASELINE.json metric: images/sec).

A "step" is ONE pass of the hot path over one synthetic batch: DynamicTeacher.forward -> BaseDistillator.distill_loss
-> backward of (loss_distill + sum_l <features_tea[l], G_l>) with fixed random cotangents G_l standing in for the
student-head gradient -> [N>1: one NCCL all-reduce of the hot-path gradients].
Workload at every N: configs[1] of BASELINE.json per GPU ("RetinaNet R-50 FPN, bs=16, synthetic COCO boxes", 800x1333
padded to 800x1344, context box on, stuGuided), i.e. weak scaling with 16 images per rank.

  python bench.py [--gpus N] [--steps K] [--warmup W]          # this repo's CUDA engine
  python bench.py --impl reference [...]                        # the reference algorithm on the host cores
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

IMG_H, IMG_W = 800, 1333
FLOPS_PER_PIXEL_CONV = 2 * 256 * 2304          # one 3x3 256->256 convolution, per output pixel
CFG_KW = dict(add_context_box=True, detach_appearance_embed=False, interact_pattern="stuGuided")
METRIC = "distillation_step_images_per_sec"
ACTIVE_KW = dict(CFG_KW)   # set from --workload in main(); the CPU arm samples 800x1333 images of that configuration


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="lgd_b200", choices=["lgd_b200", "reference"])
    ap.add_argument("--batch", type=int, default=16, help="images per GPU (BASELINE configs[1]: 16)")
    ap.add_argument("--cpu-sample-batch", type=int, default=0,
                    help="images per CPU step (0 = the full per-GPU batch when the run fits ~5 minutes, else the largest "
                         "of 16/8/4/2 that does; the number used is printed in config.images_per_step)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the parity block (engine vs CPU oracle, B=2, ~10 s)")
    ap.add_argument("--fwd-only", action="store_true", help="time teacher forward + loss only (no backward)")
    ap.add_argument("--layout", default="nchw", choices=["nchw", "channels_last"],
                    help="memory format of the FPN maps and of the teacher-pyramid cotangents at the boundary: nchw = what "
                         "the reference's detectron2 FPN hands over (default, the metric's configuration); channels_last "
                         "= the student running in channels_last memory format (no layout movers on the path)")
    ap.add_argument("--workload", default="retinanet", choices=["retinanet", "fcos", "multiscale"],
                    help="retinanet = BASELINE configs[1] (default, the metric's configuration); fcos = configs[2] "
                         "(no context box); multiscale = configs[4]'s per-GPU shape mix (short side 640..800 per step)")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(hbm=float(d["hbm_gbs"]), bf16_burst=float(d["bf16_tflops"]),
                    bf16_sustained=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), source="measured")
    return dict(hbm=6650.0, bf16_burst=1590.0, bf16_sustained=1400.0, source="fallback")


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi sampled every 200 ms while the timed region runs (B200_PROFILING.md 'clocks' line)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [x.strip() for x in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
                power.append(float(parts[2]))
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_step_fn(batch, seed=1234):
    """The reference algorithm (oracle port of the reference's PyTorch path) on the host cores: fwd + bwd."""
    import torch
    from lgd_b200 import synth
    from oracle import lgd_oracle as O
    sd = synth.synth_state_dict(0)
    bi, im, feats = synth.synth_batch(batch, IMG_H, IMG_W, seed=seed)
    hws = synth.pyramid_hw(synth.pad32(IMG_H), synth.pad32(IMG_W))
    cot = synth.synth_cotangents({k: torch.empty(batch, 256, h, w) for k, (h, w) in zip(feats, hws)})
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items()}

    def step():
        f = {k: v.detach().requires_grad_(True) for k, v in feats.items()}
        tea, _, _, loss, _ = O.distill_step(params, bi, im, f, **ACTIVE_KW)
        total = loss + sum((tea[k] * cot[k]).sum() for k in tea)
        torch.autograd.grad(total, list(f.values()) + list(params.values()), allow_unused=True)
        return float(loss)
    return step


def time_cpu_fwd(batch, steps):
    """fwd+loss only (no backward) of the oracle port: images/s"""
    import torch
    from lgd_b200 import synth
    from oracle import lgd_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    sd = synth.synth_state_dict(0)
    bi, im, feats = synth.synth_batch(batch, IMG_H, IMG_W, seed=1234)
    with torch.no_grad():
        O.distill_step(sd, bi, im, feats, **ACTIVE_KW)
        t0 = time.perf_counter()
        for _ in range(steps):
            O.distill_step(sd, bi, im, feats, **ACTIVE_KW)
    return batch * steps / (time.perf_counter() - t0)


def time_cpu(batch, steps, warmup):
    import torch
    torch.set_num_threads(os.cpu_count() or 1)
    step = cpu_step_fn(batch)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return batch * steps / dt, dt / steps


def cpu_model():
    try:
        with open("/proc/cpuinfo") as f:
            for ln in f:
                if ln.lower().startswith("model name"):
                    return ln.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


CPU_SEC_PER_IMAGE = 0.42   # fwd+bwd of the oracle port at 800x1344 on the pool's 16 host cores (measured 0.39-0.40)


def cpu_batch_for(args, nsteps, budget_s=300.0):
    """Images per CPU step: the full per-GPU batch when nsteps of it fit the time budget, else the largest power of two
    that does (the CPU images/s of this conv-bound path is flat in the batch size: 2.53 at B=2, 2.5 at B=16)."""
    if args.cpu_sample_batch > 0:
        return args.cpu_sample_batch
    b = args.batch
    while b > 2 and nsteps * b * CPU_SEC_PER_IMAGE > budget_s:
        b //= 2
    return max(b, 1)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    b = cpu_batch_for(args, args.steps + args.warmup)
    ips, sec = time_cpu(b, args.steps, args.warmup)
    cfg = workload_config(args)
    cfg["images_per_step"] = b
    sample = ("%d of the %d images of one batch per step (800x1344, %s, stuGuided), fwd+bwd of the oracle port of the "
              "reference's PyTorch path, torch CPU fp32, %d threads on %s"
              % (b, args.batch, "ctx box" if ACTIVE_KW.get("add_context_box") else "no ctx box", cores, cpu_model()))
    line = {
        "impl": "reference", "metric": METRIC, "value": ips, "unit": "images/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": cfg,
        "cpu_baseline": {"value": ips, "unit": "images/s", "cores": cores, "cpu": cpu_model(), "kind": "port",
                         "sample": sample},
        "e2e": {"value": ips, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


MULTISCALE_SHORT = (640, 672, 704, 736, 768, 800)   # configs/Base-RetinaNet.yaml:26


def cfg_kw(args):
    kw = dict(CFG_KW)
    if getattr(args, "workload", "retinanet") == "fcos":   # configs/Distillation/FCOS/fcos_R_50...yaml:22-24
        kw["add_context_box"] = False
    return kw


def image_sizes(args):
    if getattr(args, "workload", "retinanet") == "multiscale":
        return [(s, int(round(s * 1333 / 800))) for s in MULTISCALE_SHORT]
    return [(IMG_H, IMG_W)] * 2


def workload_config(args):
    wl = getattr(args, "workload", "retinanet")
    name = {"retinanet": "RetinaNet R-50 FPN distillation step (teacher fwd + distill loss + bwd), bs=%d per GPU, "
                         "800x1333->800x1344, P3-P7, synthetic COCO boxes (ctx box on, stuGuided)",
            "fcos": "FCOS R-50 FPN distillation step (teacher fwd + distill loss + bwd), bs=%d per GPU, "
                    "800x1333->800x1344, P3-P7, synthetic COCO boxes (no ctx box, stuGuided)",
            "multiscale": "RetinaNet (Swin-T recipe shapes) distillation step, bs=%d per GPU, short side cycling "
                          "through 640..800 (long = short*1333/800, padded to x32), P3-P7, ctx box on"}[wl] % args.batch
    return {"workload": name,
            "images_per_gpu": args.batch, "images_per_step": args.batch, "image_hw": [800, 1344] if wl != "multiscale" else "640x1088 .. 800x1344",
            "levels": "p3-p7", "boundary_layout": getattr(args, "layout", "nchw"),
            "step": "fwd+loss" if args.fwd_only else "fwd+loss+bwd",
            "l2": "inputs (367 MB of FPN maps per step) and every intermediate exceed the 126 MB L2"}


# ------------------------------------------------------------------------------------------------ GPU arm
def run_gpu(args):
    import torch
    import torch.distributed as dist
    from lgd_b200 import _lib, synth
    from lgd_b200.dist import ChainGradReducer, FlatGradBucket
    from lgd_b200.optim import publish_scalars
    from lgd_b200.step import HotPathDistillator

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the lgd_b200 hot path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    B = args.batch

    model = HotPathDistillator(synth.make_cfg(device="cuda", **cfg_kw(args)))
    model.load_hot_path_state_dict(synth.synth_state_dict(0))
    model = model.to(dev)
    model.teacher.return_masks = True        # the reference API returns the float masks; keep that work in
    # N > 1: one flat gradient buffer -> ONE all-reduce per step. N = 1: no collective, gradients are
    # dropped between steps like optimizer.zero_grad(set_to_none=True) in the reference's loop (train.py:201-202).
    params = list(model.parameters())
    from lgd_b200 import engine as _eng
    use_chain = _eng.chain_applicable(cfg_kw(args)["interact_pattern"])
    # native chains: one all-reduce per chain, started inside the backward (ChainGradReducer); per-kernel
    # orchestration (LGD_B200_CHAIN=0): one flat bucket, one all-reduce after the backward
    reducer = ChainGradReducer() if (world > 1 and use_chain) else None
    bucket = FlatGradBucket(params) if (world > 1 and not use_chain) else None
    comm_events = []

    # two synthetic batches (alternated), host copies pinned for the e2e leg
    batches = []
    cots = []
    for s, (ih, iw) in enumerate(image_sizes(args)):
        bi, im, feats = synth.synth_batch(B, ih, iw, seed=1234 + 1000 * rank + s)
        if args.layout == "channels_last":
            feats = {k: v.contiguous(memory_format=torch.channels_last) for k, v in feats.items()}
        host = {k: v.pin_memory() for k, v in feats.items()}
        batches.append((bi, im, host))
        cots.append({k: (v.to(dev).contiguous(memory_format=torch.channels_last) if args.layout == "channels_last"
                         else v.to(dev)) for k, v in synth.synth_cotangents(
            {k: torch.empty(B, 256, *v.shape[-2:]) for k, v in feats.items()}).items()})
    NB = len(batches)
    # pixels per image over the pyramid: the mean over the shape mix (all equal except for --workload multiscale)
    P = sum(sum(v.shape[-2] * v.shape[-1] for v in host.values()) for _, _, host in batches) / NB
    resident = [{k: v.to(dev) for k, v in host.items()} for _, _, host in batches]
    h2d_bytes = sum(sum(v.numel() * 4 for v in host.values()) for _, _, host in batches) / NB

    copy_stream = torch.cuda.Stream(dev)

    NBUF = 3
    stage_buf = [[{k: torch.empty_like(v, device=dev) for k, v in host.items()} for _, _, host in batches]
                 if args.workload == "multiscale" else
                 [{k: torch.empty_like(v, device=dev) for k, v in batches[0][2].items()}] * NB
                 for _ in range(NBUF)]   # [buffer][batch shape]
    stage_free = [None] * NBUF   # event: the step that last consumed the buffer has finished

    def stage_from_host(i):
        """Queue the pinned-host -> device copy of step i's FPN maps on the copy stream into one of three fixed device
        buffers; it runs underneath step i-1. The buffer was last read by step i-3, which the host already knows to be
        finished (it has read that step's loss), so the wait below returns at once and the copy stream never holds a
        pending event wait -- measured: a GPU-side wait on an unfinished event in front of the H2D copy serialises the
        copy with the compute kernels (+6.6 ms/step). Returns the device tensors and the event the compute stream has
        to wait for."""
        _, _, host = batches[i % NB]
        j = i % NBUF
        if stage_free[j] is not None:
            stage_free[j].synchronize()
        with torch.cuda.stream(copy_stream):
            for k, v in host.items():
                stage_buf[j][i % NB][k].copy_(v, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return stage_buf[j][i % NB], ev, j

    def one_step(i, staged=None):
        bi, im, host = batches[i % NB]
        if staged is not None:
            f, ev, j = staged
            # HOST-side wait for the input copy (already finished in steady state: it was queued a whole step earlier), so
            # the compute stream never holds a wait on the copy stream. Copy engines work in order: whatever small
            # host->device upload a step makes through one (it used to be the box table) queues behind the 367 MB
            # input copies in flight and stalls the compute stream until they end -- the first steps of a pass took
            # 24 ms instead of 11 (tools/e2e_probe.py). The library's own uploads (box table, token programs) are
            # therefore pulled from pinned memory by kernels (lgd_upload_from_host).
            ev.synchronize()
            f = {k: v.detach().requires_grad_(not args.fwd_only) for k, v in f.items()}
        else:
            f = {k: v.detach().requires_grad_(not args.fwd_only) for k, v in resident[i % NB].items()}
        if bucket is not None:
            bucket.zero_()
        else:
            for p in params:
                p.grad = None
        if args.fwd_only:
            with torch.no_grad():
                _, _, _, loss = model.forward(bi, im, f)
        else:
            _, loss = model.step(bi, im, f, cots[i % NB])
            if bucket is not None:
                bucket.all_reduce_mean()
            if reducer is not None:
                # exposed communication = how long the compute stream has to wait for the collectives at this point
                c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                c0.record()
                reducer.finish()
                c1.record()
                comm_events.append((c0, c1))
        if staged is not None:
            done = torch.cuda.Event()
            done.record()
            stage_free[staged[2]] = done
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(n, from_host, read_loss):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = lib.lgd_launch_count()
        e0.record()
        last = None
        if from_host:
            # end to end: every step's inputs come from pinned host memory (copy of step i+1 queued while step i
            # computes) and every step's loss is copied back to pinned host memory and read (one step late, so the
            # read does not drain the queue; the last one is read before the clock stops)
            copy_stream.wait_stream(torch.cuda.current_stream(dev))
            staged = stage_from_host(0)
            pending = None
            dbg = [] if os.environ.get("LGD_BENCH_DEBUG") else None
            for i in range(n):
                # the next step's input copy is queued first, so that it runs underneath this whole step
                nxt = stage_from_host(i + 1) if i + 1 < n else None
                if dbg is not None:
                    d0 = torch.cuda.Event(enable_timing=True)
                    d0.record()
                    dbg.append((d0, staged[1], time.perf_counter()))
                loss = one_step(i, staged)
                t_enq = time.perf_counter()
                staged = nxt
                slot = loss_host[i % 2]
                # read-back by a kernel store into pinned memory (keeps it off the copy engines)
                ev = publish_scalars(loss, slot)
                if pending is not None:
                    pending[1].synchronize()
                    last = float(pending[0][0])
                pending = (slot, ev)
                if dbg is not None:
                    dbg[-1] = dbg[-1] + (t_enq, time.perf_counter())
            pending[1].synchronize()
            last = float(pending[0][0])
            if dbg:
                torch.cuda.synchronize()
                print("e2e debug: step-start gaps (ms):", ["%.1f" % dbg[k][0].elapsed_time(dbg[k + 1][0])
                                                           for k in range(min(12, len(dbg) - 1))], file=sys.stderr)
                print("e2e debug: host loop (ms):", ["%.1f" % ((dbg[k + 1][2] - dbg[k][2]) * 1e3)
                                                     for k in range(min(12, len(dbg) - 1))], file=sys.stderr)
                print("e2e debug: host enqueue of the step (ms):", ["%.1f" % ((dbg[k][3] - dbg[k][2]) * 1e3)
                                                                    for k in range(min(12, len(dbg)))], file=sys.stderr)
                print("e2e debug: host wait for previous loss (ms):", ["%.1f" % ((dbg[k][4] - dbg[k][3]) * 1e3)
                                                                       for k in range(min(12, len(dbg)))], file=sys.stderr)
        else:
            for i in range(n):
                one_step(i)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        launches = lib.lgd_launch_count() - l0
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms, launches, last

    loss_host = [torch.empty(1, dtype=torch.float32).pin_memory() for _ in range(2)]
    # every shape of the mix is seen twice before the clock starts (the caching allocator settles per shape)
    n_warm = max(args.warmup, 2 * NB) if args.workload == "multiscale" else args.warmup
    for i in range(n_warm):
        one_step(i)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    comm_events.clear()
    ms, launches, _ = timed(args.steps, False, False)
    comm_ms = (sum(a.elapsed_time(b) for a, b in comm_events) / len(comm_events)) if comm_events else None
    # end-to-end: host buffers in, loss out, every step
    timed(min(4, args.steps), True, True)   # warm the pipelined path (pinned staging, allocator) before timing it
    ms_e2e, _, last_loss = timed(args.steps, True, True)
    # The device-resident loop does strictly less work per step than the end-to-end loop. When it nevertheless came out
    # more than 5 % slower, its pass was disturbed (seen once in ~30 runs on the shared pool: 22.7 vs 19.1 ms): it is
    # re-measured ONCE, the same K steps, and both timings are reported.
    remeasured = None
    if ms > 1.05 * ms_e2e:
        remeasured = {"first_pass_ms_per_step": ms / args.steps, "reason": "device-resident pass slower than the "
                      "end-to-end pass of the same run; re-measured once"}
        ms, launches, _ = timed(args.steps, False, False)
    clocks = sampler.stop() if rank == 0 else None

    # live per-entry-point device time (CUDA events on the launching stream) for the roofline. The wgrad side stream
    # is switched off for this pass only, so that every duration is that of a kernel running alone on the GPU.
    from lgd_b200 import engine as _engine
    _engine.WGRAD_SIDE_STREAM = False
    _lib.profile = []
    barrier()
    nprof = NB if args.workload == "multiscale" else min(args.steps, 3)   # whole shape cycles only
    pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pe0.record()
    for i in range(nprof):
        one_step(i)
    pe1.record()
    torch.cuda.synchronize()
    prof_pass_ms = pe0.elapsed_time(pe1) / nprof
    prof = _engine.drain_profile()     # per-kernel calls made from Python + the calls inside the native chains
    _lib.profile = None
    _engine.WGRAD_SIDE_STREAM = True
    per = {}
    for name, ms_call in prof:
        d = per.setdefault(name, [0.0, 0])
        d[0] += ms_call
        d[1] += 1
    total_prof_ms = sum(v[0] for v in per.values()) / nprof

    # SURVEY 8(d) asks for both step definitions: the headline is fwd+loss+bwd; fwd+loss (teacher forward + distill
    # loss, no backward -- the definition of BASELINE configs[0]) is reported beside it
    fwd_loss = None
    if not args.fwd_only:
        nf = min(args.steps, 20)

        def fwd_step(i):
            bi, im, _ = batches[i % NB]
            with torch.no_grad():
                model.forward(bi, im, resident[i % NB])
        for i in range(2):
            fwd_step(i)
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for i in range(nf):
            fwd_step(i)
        f1.record()
        barrier()
        ms_f = f0.elapsed_time(f1)
        if world > 1:
            t = torch.tensor([ms_f], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_f = float(t)
        fwd_loss = {"value": world * B * nf / (ms_f * 1e-3), "unit": "images/s", "ms_per_step": ms_f / nf, "steps": nf}

    # SURVEY 8(f) rank 4, reported beside the metric (the metric's step excludes the optimizer): one
    # optimizer step over the hot-path parameters with the reference's grouping (one group per parameter), torch.optim.SGD
    # vs the multi-tensor kernel
    opt_ms = None
    if not args.fwd_only and rank == 0:
        from lgd_b200.optim import FusedSGD
        for p in params:
            if p.grad is None:
                p.grad = torch.zeros_like(p)
        groups = lambda: [{"params": [p], "lr": 0.0, "weight_decay": 1e-4} for p in params]   # lr 0: weights stay put
        opt_ms = {}
        for label, opt in (("torch_sgd_per_param_groups", torch.optim.SGD(groups(), 0.0, momentum=0.9)),
                           ("lgd_fused_sgd", FusedSGD(groups(), 0.0, momentum=0.9))):
            for _ in range(3):
                opt.step()
            torch.cuda.synchronize()
            o0, o1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            o0.record()
            for _ in range(10):
                opt.step()
            o1.record()
            t1 = time.perf_counter()
            torch.cuda.synchronize()
            opt_ms[label] = {"device_ms": o0.elapsed_time(o1) / 10, "host_enqueue_ms": (t1 - t0) * 100.0}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    pk = peaks()
    ips = world * B * args.steps / (ms * 1e-3)
    ips_e2e = world * B * args.steps / (ms_e2e * 1e-3)
    flops_launch = float(FLOPS_PER_PIXEL_CONV) * B * P
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "conv3x3_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f).get("dram_bytes_per_launch")

    def tensor_roofline(names, label, peak, peak_note, burst):
        t = sum(per[n][0] for n in names if n in per)
        c = sum(per[n][1] for n in names if n in per)
        if c == 0:
            return None
        ms1 = t / c
        ach = flops_launch / (ms1 * 1e-3) / 1e12
        return {"kernel": label, "bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                "frac_of_burst_peak": ach / burst, "avg_launch_ms": ms1, "launches_per_step": c / nprof,
                "flops_per_launch": flops_launch, "share_of_step": (t / nprof) / total_prof_ms if total_prof_ms > 0 else None,
                "peak_note": peak_note}

    note16 = ("%s sustained bf16 cuBLAS peak of MEASURED_PEAKS.json (16-bit operands, fp32 accumulate; the kernel is "
              "timed inside a long step); burst figure %.1f TF/s" % (pk["source"], pk["bf16_burst"]))
    note32 = "TF32 = 1/2 of the %s sustained bf16 cuBLAS peak in MEASURED_PEAKS.json" % pk["source"]
    per_kernel = {
        "forward (conv3x3_tc_kernel<f16>, kind::f16)": tensor_roofline(
            ["lgd_conv3x3_fwd_f16"], "conv3x3_tc_kernel<f16> forward", pk["bf16_sustained"], note16, pk["bf16_burst"]),
        "dgrad (conv3x3_tc_kernel<f16>, scaled fp16 gradients)": tensor_roofline(
            ["lgd_conv3x3_dgrad_f16", "lgd_conv3x3_dgrad_f16_gnsums", "lgd_conv3x3_dgrad_f16_gnsums_y"], "conv3x3_tc_kernel<f16> dgrad", pk["bf16_sustained"],
            note16, pk["bf16_burst"]),
        "wgrad (conv3x3_wgrad_kernel<f16>, MN-major)": tensor_roofline(
            ["lgd_conv3x3_wgrad_f16"], "conv3x3_wgrad_kernel<f16>", pk["bf16_sustained"], note16, pk["bf16_burst"]),
        "tf32 forward/dgrad (fallback paths)": tensor_roofline(
            ["lgd_conv3x3_fwd", "lgd_conv3x3_fwd_addend"], "conv3x3_tc_kernel<tf32>", pk["bf16_sustained"] / 2, note32,
            pk["bf16_burst"] / 2),
        "tf32 wgrad (fallback path)": tensor_roofline(
            ["lgd_conv3x3_wgrad"], "conv3x3_wgrad_kernel<tf32>", pk["bf16_sustained"] / 2, note32, pk["bf16_burst"] / 2),
    }
    per_kernel = {k: v for k, v in per_kernel.items() if v is not None}
    # the dominant kernel family of the step: all 3x3 convolutions on 16-bit operands (falls back to the TF32 family
    # when LGD_B200_BWD_F16=0 / tf32x3 made those the majority)
    f16_names = ["lgd_conv3x3_fwd_f16", "lgd_conv3x3_dgrad_f16", "lgd_conv3x3_dgrad_f16_gnsums",
                 "lgd_conv3x3_dgrad_f16_gnsums_y", "lgd_conv3x3_wgrad_f16"]
    tf32_names = ["lgd_conv3x3_fwd", "lgd_conv3x3_fwd_addend", "lgd_conv3x3_wgrad"]
    t16 = sum(per[n][0] for n in f16_names if n in per)
    t32 = sum(per[n][0] for n in tf32_names if n in per)
    if t16 >= t32:
        roofline = tensor_roofline(f16_names, "3x3 convolutions on tcgen05 cta_group::2 kind::f16 (fp16 operands, fp32 "
                                   "accumulate): conv3x3_tc_kernel<f16> forward + dgrad, conv3x3_wgrad_kernel<f16>",
                                   pk["bf16_sustained"], note16, pk["bf16_burst"])
    else:
        roofline = tensor_roofline(tf32_names, "3x3 convolutions on tcgen05 cta_group::2 kind::tf32",
                                   pk["bf16_sustained"] / 2, note32, pk["bf16_burst"] / 2)
    roofline["traffic"] = traffic
    n_conv = roofline["launches_per_step"]   # of the profiled steps (forward-only steps with --fwd-only)
    roofline["timing_note"] = ("CUDA events around every launch in a separate pass of %d steps with the side streams "
                               "disabled (kernels run alone); the timed region itself overlaps the chains" % nprof)
    roofline16 = per_kernel
    F1 = 1024.0 * P * B   # bytes of one fp32 pyramid tensor for the whole batch
    hbm = {}
    # algorithmic bytes (fp32 = F1 per pyramid tensor). With the fp16 backward the tensors that only feed convolutions
    # are written as fp16 only (F1/2): GroupNorm apply / backward outputs, the IN-MSE gradient, the rendering.
    lean = _engine._bwd_f16()
    for name, nbytes in (("lgd_in_mse_moments_fwd", 2 * F1), ("lgd_maskpool_fwd", F1),
                         ("lgd_gn_apply", ((1.5 + 1.5 + 2) / 3 if lean else 2) * F1),   # the third call writes fp32
                         ("lgd_gn_bwd", (4.5 if lean else 5) * F1), ("lgd_gn_bwd_tile_sums", 2.5 * F1),
                         ("lgd_in_mse_bwd", (2.5 if lean else 3) * F1),
                         ("lgd_render_fwd", (0.5 if lean else 1) * F1), ("lgd_render_bwd", F1),
                         ("lgd_maskpool_bwd", F1)):
        if name in per and per[name][0] > 0:
            t = per[name][0] / per[name][1]
            gbs = nbytes / (t * 1e-3) / 1e9
            hbm[name] = {"avg_ms": t, "algorithmic_bytes": nbytes, "achieved_gbs": gbs, "frac": gbs / pk["hbm"]}
    if "lgd_maskpool_fwd" in hbm and "lgd_in_mse_moments_fwd" in hbm:
        # "pooling + L2 reduction" forward = 3 F1 per batch (projected student map read once by the
        # pooling; adapter output and teacher pyramid read once by the loss, InstanceNorm statistics included)
        t = hbm["lgd_maskpool_fwd"]["avg_ms"] + hbm["lgd_in_mse_moments_fwd"]["avg_ms"]
        gbs = 3 * F1 / (t * 1e-3) / 1e9
        hbm["pooling_plus_l2_reduction_fwd"] = {"avg_ms": t, "algorithmic_bytes": 3 * F1, "achieved_gbs": gbs,
                                                "frac": gbs / pk["hbm"]}
    breakdown = {k: {"ms_per_step": v[0] / nprof, "calls_per_step": v[1] / nprof}
                 for k, v in sorted(per.items(), key=lambda kv: -kv[1][0])}

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        cores = os.cpu_count() or 1
        cb = cpu_batch_for(args, 3, budget_s=24.0)
        cips, csec = time_cpu(cb, 2, 1)
        cpu = {"value": cips, "unit": "images/s", "cores": cores, "cpu": cpu_model(), "kind": "port",
               "fwd_loss_only_value": time_cpu_fwd(min(cb, 4), 2),
               "sample": "%d of the %d images per step, 1 warm-up + 2 timed fwd+bwd steps of the oracle port "
                         "(torch CPU fp32, %d threads), %.2f s/step" % (cb, B, cores, csec)}

    # parity of THIS build on THIS box, next to the throughput: engine vs the CPU oracle on a B=2 sample of the workload
    # at full image size (oracle/parity.py; the assertions live in tests/test_gpu_parity.py)
    par = None
    if not args.no_parity and world == 1 and not args.fwd_only:
        from oracle import parity as _parity
        r = _parity.step_parity(cfg_kw(args), 2, (IMG_H, IMG_W), seed=77)
        par = {"sample": "B=2, 800x1333->800x1344, seed 77, vs the fp32 CPU oracle", "masks_bit_exact": r["masks_exact"],
               "loss_rel_err": r["loss_err"], "teacher_pyramid_rel_l2": r["fwd_err"],
               "relu_flip_fraction": r["flip_fraction"], "relu_flips": r["flips"], "relu_decisions": r["activations"],
               "flip_margin_rel_rms": r["flip_margin"],
               "worst_grad_rel_l2_same_activation_pattern": r["grad_err_pattern"],
               "worst_grad_same_pattern_tensor": r["grad_err_pattern_worst"],
               "worst_grad_rel_l2_plain_fp32": r["grad_err_plain"], "worst_grad_plain_tensor": r["grad_err_plain_worst"]}

    line = {
        "metric": METRIC, "value": ips, "unit": "images/s", "n_gpus": world, "steps": args.steps,
        "warmup": n_warm, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "fp16 tensor-core operands (forward, dgrad, wgrad; gradients power-of-two scaled), f32 accumulate; f32 storage of everything that is not a conv operand", "data": "synthetic",
        "config": workload_config(args), "clocks": clocks,
        "e2e": {"value": ips_e2e, "unit": "images/s", "h2d_bytes_per_step": h2d_bytes + 0, "d2h_bytes_per_step": 4,
                "ms_per_step": ms_e2e / args.steps, "loss": last_loss,
                "h2d_gbs_aggregate": world * h2d_bytes / (ms_e2e / args.steps * 1e-3) / 1e9,
                "note": "every step: FPN maps copied from pinned host memory (copy of step i+1 queued on a copy stream "
                        "while step i computes), plugin API DynamicTeacher.forward / distill_loss / backward, loss "
                        "copied to pinned host memory and read; all inside the timed region; cotangents stay on device"},
        "fwd_loss_only": fwd_loss, "gpu_launches": launches, "optimizer_step_ms": opt_ms,
        # local_inst_proj_2D evaluated from per-box tap vectors (csrc/taprender.cu) instead of a convolution over the
        # rendered map: 21 instead of 24 convolution launches per step (LGD_B200_TAP_RENDER=0 restores the convolution)
        "tap_render": bool(_engine.TAP_RENDER),
        "comm": None if world == 1 else {
            "exposed_ms_per_step": comm_ms, "collectives_per_step": 3 if reducer is not None else 1,
            "bytes_per_step": sum(p.numel() for p in params) * 4,
            "note": "gradient average of the hot-path parameters over NCCL; native chains: adapter gradients (7 MB) "
                    "all-reduced underneath the teacher backward, teacher gradients except student_proj_2D (31 MB) "
                    "underneath the student-side end of the teacher backward, student_proj_2D (2.4 MB) at its end; exposed = time "
                    "the compute stream waits for them (CUDA events around the wait, rank 0)"}, "roofline": roofline, "roofline_per_kernel": roofline16, "roofline_hbm": hbm, "cpu_baseline": cpu, "parity": par,
        # convolution launches actually made per step (24 = 8 forward + 8 dgrad + 8 wgrad; 21 when local_inst_proj_2D is
        # evaluated from per-box tap vectors instead of a convolution, LGD_B200_TAP_RENDER=1): FLOPs executed, not the
        # FLOPs of the reference's formulation
        "flops_per_step": n_conv * flops_launch,
        "step_tflops": n_conv * flops_launch * world / (ms / args.steps * 1e-3) / 1e12,
        "serial_pass": {"ms_per_step": prof_pass_ms, "sum_of_library_calls_ms": total_prof_ms,
                        "note": "wgrad side stream off + an event pair per call; the difference is torch's own kernels "
                                "(gradient accumulation, zero_) and launch gaps"},
        "breakdown_ms": breakdown, "remeasured": remeasured,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    ACTIVE_KW.clear()
    ACTIVE_KW.update(cfg_kw(args))
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()

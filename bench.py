#!/usr/bin/env python
"""Benchmark of the PoET deformable encoder/decoder hot path on B200 (driver contract).

    python bench.py --gpus N --steps K --warmup W            # our CUDA path (one rank per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K --warmup W   # reference arm: CPU oracle port on host cores

Workload (config.workload): BASELINE.json configs[1] — YCB-V shape: 5 enc / 5 dec layers, 16 heads,
d=256, 10 queries, 22 class slots, batch 16 per GPU, 640x480 REF pyramid (S=1600), forward+backward
through the fixed-cotangent loss of SURVEY.md §8d, fp32.  A "step" = one forward+backward over one
batch (N>1: plus the flat-buffer gradient all-reduce).  Metric: images/s.

Timing: per-step CUDA events on the launching stream, an L2 flush (256 MiB write) between timed
steps outside the events, barrier + synchronize on both sides, MAX over ranks.

Keys beyond the driver contract: `e2e` (pinned host inputs -> GraphedStep.prefetch on a copy stream -> replay ->
read-back of loss + final predictions every step; `serial_value` = without the copy pipeline), `roofline`
(+ `alt`: the reduction-issue bound of the MSDA backward), `kernels` (per-kernel table; GEMM rows carry
`issued_tflops` / `tensor_pipe_frac`), `cpu_baseline`, `clocks`.
Optional rows, each stated in `config`: `--from-features` (input_proj inside the step),
`--criterion` (on-device SetCriterion + 'gt' matcher instead of the fixed-cotangent loss), `--optimizer`
(clip_grad_norm_ + AdamW inside the step); `--micro-batches`, `--no-graph`, `--precision` are A/B knobs.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

WORKLOAD = "cfg2"            # default; --workload selects another BASELINE.json config (module-level so helpers see it)
POET_SMS = 148

# BASELINE.json `configs` -> (synthetic config, what a step is, description)
WORKLOADS = {
    "cfg1": dict(cfg="cfg1", backward=False,
                 text="cfg1: 1x640x480, 2enc/2dec/8h d256 Q5 (reference CPU eval case), eval forward only"),
    "cfg2": dict(cfg="cfg2", backward=True,
                 text="cfg2: YCB-V 5enc/5dec/16h d256 Q10 22 class slots, fwd+bwd, synthetic cotangent loss"),
    "cfg3": dict(cfg="cfg3", backward=True,
                 text="cfg3: LM-O 5enc/5dec/16h d256 Q10 9 class slots (heads 27/54), fwd+bwd, synthetic cotangent loss"),
    "cfg4": dict(cfg="cfg4", backward=True,
                 text="cfg4: YCB-V training step in the bf16 throughput mode (single-pass bf16 MMAs in the encoder, bf16x3 "
                      "elsewhere), fwd+bwd + gradient all-reduce + fused clip+AdamW"),
    "cfg5": dict(cfg="cfg5", backward=True,
                 text="cfg5: 1280x960 REF pyramid (S=6380), 6enc/6dec/8h d256 Q25, fwd+bwd, global batch 32 (strong scaling)"),
}


def metric_name(workload, cfg, per_gpu_batch):
    from poet_b200 import synthetic as S
    res = "1280x960" if cfg["pyramid"] == "REF1280" else "640x480"
    what = "fwd+bwd" if WORKLOADS[workload]["backward"] else "eval fwd"
    return (f"images/sec, PoET deformable enc/dec {what} ({cfg['enc_layers']}enc/{cfg['dec_layers']}dec/{cfg['nheads']}h, "
            f"{res} REF pyramid, batch {per_gpu_batch}/GPU)")


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            p = json.load(fh)
        return dict(hbm=float(p["hbm_gbs"]), tf_burst=float(p["bf16_tflops"]), tf_sust=float(p["bf16_tflops_sustained"]),
                    source="measured")
    except Exception:
        return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, source="fallback")


def config_dict(cfg, extra=None):
    from poet_b200 import synthetic as S
    d = {"workload": WORKLOADS[WORKLOAD]["text"],
         "batch_per_gpu": cfg["batch"], "pyramid": S.pyramid_of(cfg), "tokens": S.n_tokens(cfg),
         "dropout": 0.0, "l2": "flushed between timed steps (256 MiB write, outside the events)"}
    if extra:
        d.update(extra)
    return d


# ------------------------------------------------------------------------------------------
# CPU legs (the only places bench.py executes oracle/)
# ------------------------------------------------------------------------------------------
def default_batch(workload, cfg, world):
    """Per-GPU batch: BASELINE.json quotes cfg5 at a GLOBAL batch of 32 swept over 1/2/4/8 GPUs (strong scaling);
    every other config keeps its per-GPU batch as GPUs are added (weak scaling)."""
    return max(1, 32 // world) if workload == "cfg5" else cfg["batch"]


def scaling_of(workload):
    return "strong" if workload == "cfg5" else "weak"


def cpu_step_fn(cfg, batch, backward=True):
    from oracle import poet_oracle as O
    from poet_b200 import synthetic as S
    P = {k: v.requires_grad_(backward) for k, v in S.make_params(cfg).items()}
    inp = S.make_inputs(cfg, batch=batch)
    g_t, g_R = S.make_cotangents(cfg, batch=batch)

    def step():
        for v in P.values():
            v.grad = None
        cap = {}
        with torch.set_grad_enabled(backward):
            O.poet_path_forward(P, cfg, inp["srcs"], inp["masks"], inp["boxes"], inp["labels"], capture=cap)
            loss = O.synthetic_loss((cap["translation_all"], cap["rotation_all"]), g_t, g_R)
        if backward:
            loss.backward()
        return float(loss.detach())
    return step


def cpu_baseline(cfg, batch=4, reps=2, backward=True):
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    if cfg["pyramid"] == "REF1280":
        batch, reps = 1, 1
    batch = min(batch, cfg["batch"])
    step = cpu_step_fn(cfg, batch, backward=backward)
    step()                                             # warm-up
    t0 = time.perf_counter()
    for _ in range(reps):
        step()
    dt = (time.perf_counter() - t0) / reps
    return {"value": batch / dt, "unit": "images/s", "cores": cores, "kind": "port",
            "sample": f"{reps} {'fwd+bwd' if backward else 'eval fwd'} steps of {batch} images of the same workload "
                      f"(oracle/poet_oracle.py, torch CPU fp32, {cores} threads) after 1 warm-up"}


def parity_check(cfg, inp, out, n_images=4):
    """The benchmarked path against the CPU oracle on the first images of the benchmarked batch (images are
    independent): max |translation|, |rotation| error over ALL decoder layers; tolerances of BASELINE.json."""
    from oracle import poet_oracle as O
    from poet_b200 import synthetic as S
    n = min(n_images, len(inp["boxes"]))
    P = S.make_params(cfg)
    cap = {}
    with torch.no_grad():
        O.poet_path_forward(P, cfg, [s[:n] for s in inp["srcs"]], [m[:n] for m in inp["masks"]], inp["boxes"][:n],
                            inp["labels"][:n], capture=cap)
    t = torch.stack([a["pred_translation"] for a in out["aux_outputs"]] + [out["pred_translation"]]).detach().cpu()
    R = torch.stack([a["pred_rotation"] for a in out["aux_outputs"]] + [out["pred_rotation"]]).detach().cpu()
    dt = float((t[:, :n] - cap["translation_all"]).abs().max())
    dR = float((R[:, :n] - cap["rotation_all"]).abs().max())
    return {"checked": f"first {n} images of the benchmarked batch, all {t.shape[0]} decoder layers, vs oracle/poet_oracle.py (fp32 CPU)",
            "max_abs_translation": dt, "max_abs_rotation": dR, "tol_translation": 1e-4, "tol_rotation": 1e-3,
            "ok": bool(dt <= 1e-4 and dR <= 1e-3)}


def run_reference(args):
    """--impl reference: the reference's algorithm on the host cores.  The reference is pure Python and
    /root/reference does not exist on the GPU box, so this is the oracle port (kind 'port')."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from poet_b200 import synthetic as S
    wl = WORKLOADS[WORKLOAD]
    cfg = dict(S.CONFIGS[wl["cfg"]])
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    batch = 1 if cfg["pyramid"] == "REF1280" or cfg["batch"] == 1 else 2      # bounded sample per step
    step = cpu_step_fn(cfg, batch, backward=wl["backward"])
    for _ in range(max(1, min(args.warmup, 2))):
        step()
    times = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
    total = sum(times)
    value = batch * args.steps / total
    what = "fwd+bwd" if wl["backward"] else "eval fwd"
    sample = f"each step = {what} of {batch} images of {WORKLOAD} (bounded sample), oracle port, {cores} threads"
    per_gpu = args.batch or default_batch(WORKLOAD, cfg, max(1, args.gpus))
    line = {"impl": "reference", "metric": metric_name(WORKLOAD, cfg, per_gpu), "value": value, "unit": "images/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
            "higher_is_better": True, "scaling": scaling_of(WORKLOAD), "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
            "config": config_dict(dict(cfg, batch=per_gpu), {"batch_per_step": batch}),
            "cpu_baseline": {"value": value, "unit": "images/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


THROUGHPUT_TOL = (2e-2, 5e-2)        # cfg4 mixed mode: |translation|, |rotation| vs the fp32 oracle (stated in DESIGN.md section 2)


def parity_check_throughput(cfg, inp, dev):
    model = build_gpu_model(cfg, dev)
    model.transformer.set_throughput_mode(True)
    with torch.no_grad():
        out, _ = model.forward_pyramid([s.to(dev) for s in inp["srcs"]], [m.to(dev) for m in inp["masks"]],
                                       [b.to(dev) for b in inp["boxes"]], [l.to(dev) for l in inp["labels"]])
    p = parity_check(cfg, inp, out)
    p["tol_translation"], p["tol_rotation"] = THROUGHPUT_TOL
    p["ok"] = bool(p["max_abs_translation"] <= THROUGHPUT_TOL[0] and p["max_abs_rotation"] <= THROUGHPUT_TOL[1])
    p["checked"] += " -- throughput (bf16) mode, its own tolerance"
    return p


# ------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu, self.proc = gpu_index, None
        self.path = tempfile.mktemp(prefix="poet_clocks_", suffix=".csv")

    def start(self):
        try:
            self.fh = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.gpu)], stdout=self.fh, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.fh.close()
        sm, mx, reasons, power = [], [], set(), []
        try:
            for ln in open(self.path):
                f = [x.strip() for x in ln.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
                except ValueError:
                    continue
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            busy = [s for s, p in zip(sm, power) if p >= 0.5 * max(power)] or sm
            out.update(sm_mhz=statistics.median(busy), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm),
                       power_w_max=max(power))
        return out


class _FeatureBackbone(torch.nn.Module):
    """Sizes PoET.input_proj like the reference's Mask R-CNN FPN backbone (three 256-channel maps, backbone_maskrcnn.py:41-42)."""

    def __init__(self, channels):
        super().__init__()
        self.strides, self.num_channels = [8, 16, 32], [channels] * 3


def build_gpu_model(cfg, dev, with_input_proj=False, dropout=0.0):
    from poet_b200 import synthetic as S
    from poet_b200.deformable_transformer import DeformableTransformer
    from poet_b200.pose_estimation_transformer import PoET
    tr = DeformableTransformer(cfg["d_model"], cfg["nheads"], cfg["enc_layers"], cfg["dec_layers"], cfg["dim_ff"], dropout,
                               "relu", True, cfg["n_levels"], cfg["n_points"], cfg["n_points"])
    backbone = _FeatureBackbone(cfg["d_model"]) if with_input_proj else None
    model = PoET(backbone, tr, cfg["num_queries"], cfg["n_levels"], cfg["n_classes"], class_mode=cfg["class_mode"])
    model.load_state_dict(S.make_params(cfg, with_input_proj=with_input_proj), strict=True)
    return model.to(dev).train()


def run_gpu(args):
    import torch.distributed as dist
    from poet_b200 import ops, synthetic as S
    from poet_b200.data_parallel import FlatGradReducer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run")
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    ops.set_gemm_precision(args.precision)

    wl = WORKLOADS[WORKLOAD]
    cfg = dict(S.CONFIGS[wl["cfg"]])
    cfg["batch"] = B = args.batch or default_batch(WORKLOAD, cfg, world)
    do_backward = wl["backward"]
    # layer-count ablations (marginal cost of one encoder / decoder layer inside the replayed graph); never a bench line:
    # the printed config carries "ablation" and the metric name is suffixed
    ablation = {k: int(os.environ[e]) for k, e in (("enc_layers", "POET_BENCH_ENC_LAYERS"), ("dec_layers", "POET_BENCH_DEC_LAYERS"))
                if os.environ.get(e)}
    cfg.update(ablation)
    model = build_gpu_model(cfg, dev, with_input_proj=args.from_features, dropout=args.dropout)
    if not do_backward:
        model.eval()
    throughput = WORKLOAD == "cfg4"
    if throughput:                                          # cfg4: bf16 training step (mixed mode + optimizer + all-reduce)
        model.transformer.set_throughput_mode(True)
        args.optimizer = True
    model.micro_batches = args.micro_batches
    reducer = FlatGradReducer(model.parameters())
    inp = S.make_inputs(cfg, seed=1234 + rank)           # each rank owns a different image shard (weak scaling)
    g_t, g_R = (t.to(dev) for t in S.make_cotangents(cfg))
    if args.from_features:                                  # three feature maps + their masks + the padded-image mask
        H0, W0 = S.pyramid_of(cfg)[0]
        inp["srcs"] = inp["srcs"][:3]
        inp["masks"] = inp["masks"][:3] + [torch.zeros(B, H0 * 16, W0 * 16, dtype=torch.bool)]
    d_srcs = [s.to(dev) for s in inp["srcs"]]
    d_masks = [m.to(dev) for m in inp["masks"]]
    d_boxes = [b.to(dev) for b in inp["boxes"]]
    d_labels = [l.to(dev) for l in inp["labels"]]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def synthetic_loss(out):
        t = torch.stack([a["pred_translation"] for a in out["aux_outputs"]] + [out["pred_translation"]])
        R = torch.stack([a["pred_rotation"] for a in out["aux_outputs"]] + [out["pred_rotation"]])
        return (t * g_t).sum() + (R * g_R).sum()

    loss_fn = synthetic_loss
    if args.criterion:
        from poet_b200.criterion import PoseCriterion
        Q = cfg["num_queries"]
        gen = torch.Generator().manual_seed(777 + rank)
        tgt_t = torch.randn(B, Q, 3, generator=gen).to(dev)
        q4 = torch.nn.functional.normalize(torch.randn(B, Q, 4, generator=gen), dim=-1)      # random unit quaternions -> SO(3)
        w, x, y, z = q4.unbind(-1)
        tgt_R = torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w),
                             2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
                             2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], -1).view(B, Q, 3, 3).to(dev)
        n_boxes_dev = torch.tensor([min(int(b.shape[0]), Q) for b in inp["boxes"]], dtype=torch.int32, device=dev)
        crit = PoseCriterion({"loss_trans": 1.0, "loss_rot": 1.0})                            # reference defaults (main.py:121-122)

        def loss_fn(out):
            return crit(out, tgt_t, tgt_R, n_boxes_dev)[1]

    def fwd(srcs, masks, boxes, labels):
        if args.from_features:
            return model.forward_features(srcs, masks[:-1], masks[-1], boxes, labels)
        return model.forward_pyramid(srcs, masks, boxes, labels)

    def eager_step(srcs, masks, boxes, labels):
        if not do_backward:
            with torch.no_grad():
                out, _ = fwd(srcs, masks, boxes, labels)
                return loss_fn(out), out
        reducer.zero()
        out, _ = fwd(srcs, masks, boxes, labels)
        loss = loss_fn(out)
        loss.backward()
        reducer.all_reduce()
        return loss, out

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    opt = None
    if args.optimizer:
        from poet_b200.optim import FusedClipAdamW
        opt = FusedClipAdamW(model, reducer, lr=2e-4, weight_decay=1e-4, max_norm=0.1)     # reference defaults (main.py)

    graphed = None
    if args.graph:
        from poet_b200.graph import GraphedStep
        graphed = GraphedStep(model, loss_fn, d_srcs, d_masks, d_boxes, d_labels, reducer=reducer, optimizer=opt,
                              entry="features" if args.from_features else "pyramid", backward=do_backward,
                              overlap_allreduce=world > 1 and args.overlap_allreduce)

        def step(srcs=None, masks=None, boxes=None, labels=None):
            loss, out = graphed.run(srcs, masks, boxes, labels)
            if do_backward and not graphed.reduces:          # else: reduced segment by segment inside the replayed graph
                reducer.all_reduce()
            if opt is not None:
                opt.step()
            return loss, out
    else:
        def step(srcs=None, masks=None, boxes=None, labels=None):
            if srcs is None:
                srcs, masks, boxes, labels = d_srcs, d_masks, d_boxes, d_labels
            r = eager_step(srcs, masks, boxes, labels)
            if opt is not None:
                opt.step()
            return r

    for _ in range(max(args.warmup, 3)):
        loss0, out0 = step()
    sync_all()
    # parity of the benchmarked path itself (same model, same inputs, same graph replay) against the CPU oracle
    parity = None
    if rank == 0 and not args.no_parity and not args.from_features and (opt is None or throughput) and args.dropout == 0.0:
        # cfg4's own tolerance (tests/test_gpu_model.py::test_throughput_mode_tolerance): the optimizer has already moved the
        # weights by the warm-up steps there, so the check runs on a fresh copy of the model in the same mode
        parity = parity_check(cfg, inp, out0) if not throughput else parity_check_throughput(cfg, inp, dev)
        if not parity["ok"]:
            print(f"bench.py: PARITY FAILED on the benchmarked path: {parity}", file=sys.stderr, flush=True)

    # ---- device-resident timed region -------------------------------------------------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    evs = []
    sync_all()
    for _ in range(args.steps):
        flush.fill_(1)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        step()
        e.record()
        evs.append((s, e))
    sync_all()
    total_ms = sum(s.elapsed_time(e) for s, e in evs)
    tt = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    total_ms = float(tt.item())

    # ---- end-to-end: host buffers in, host result out, through the public module API ---------
    h_srcs = [s.pin_memory() for s in inp["srcs"]]
    h_masks = [m.pin_memory() for m in inp["masks"]]
    h2d = sum(t.numel() * t.element_size() for t in h_srcs + h_masks) + sum(b.numel() * 4 + l.numel() * 8 for b, l in zip(inp["boxes"], inp["labels"]))
    d2h = 0

    d2h_pinned = torch.empty(1 + B * cfg["num_queries"] * 12, dtype=torch.float32).pin_memory()

    def e2e_loop(pipelined: bool) -> float:
        """Wall-clock ms of args.steps end-to-end steps (one untimed pass first).  pipelined: the H2D copy of
        step i+1's inputs is issued on the copy stream before step i is replayed (GraphedStep.prefetch), so every
        timed step still contains one full input copy and one result read-back, overlapped with compute."""
        nonlocal d2h
        total = 0.0
        sync_all()
        if pipelined:
            graphed.prefetch(h_srcs, h_masks, inp["boxes"], inp["labels"])
        for it in range(args.steps + 1):
            flush.fill_(1)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            if pipelined:
                loss, out = step()                                                   # consumes the oldest prefetched set
                graphed.prefetch(h_srcs, h_masks, inp["boxes"], inp["labels"])     # inputs of the NEXT step, during this replay
            elif graphed is not None:
                loss, out = step(h_srcs, h_masks, inp["boxes"], inp["labels"])     # H2D into the graph's static buffers
            else:
                srcs = [s.to(dev, non_blocking=True) for s in h_srcs]
                masks = [m.to(dev, non_blocking=True) for m in h_masks]
                loss, out = step(srcs, masks, inp["boxes"], inp["labels"])          # host box lists: padded on host, one H2D
            res = torch.cat((loss.detach().reshape(1), out["pred_translation"].detach().reshape(-1),
                             out["pred_rotation"].detach().reshape(-1)))
            d2h_pinned.copy_(res, non_blocking=True)                            # one read-back of loss + final predictions
            torch.cuda.synchronize()
            host = [d2h_pinned]
            if it > 0:                                                         # first pass warms the pinned path
                total += 1e3 * (time.perf_counter() - t0)
            d2h = sum(t.numel() * t.element_size() for t in host)
        if pipelined:
            step()                                                             # drain the last prefetched set
            torch.cuda.synchronize()
        t = torch.tensor([total], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    e2e_serial_ms = e2e_loop(False)
    e2e_ms = e2e_loop(True) if graphed is not None else e2e_serial_ms
    clocks = sampler.stop() if rank == 0 else None      # sampled over the device-timed and the end-to-end timed regions

    # ---- per-kernel table: an eager pass with CUDA events around every library call ----------
    # (events cannot bracket nodes inside a replayed graph; the kernels and their arguments are identical)
    ktimes, ksteps = {}, min(args.steps, 5)

    def local_step():                                   # no collective: only rank 0 runs the table pass
        if not do_backward:
            with torch.no_grad():
                fwd(d_srcs, d_masks, d_boxes, d_labels)
            return
        reducer.zero()
        out, _ = fwd(d_srcs, d_masks, d_boxes, d_labels)
        loss_fn(out).backward()

    l0 = ops.launch_count()
    local_step()
    launches = (ops.launch_count() - l0) * args.steps          # library launches replayed per step x timed steps
    if args.kernel_table and rank == 0:
        torch.cuda.synchronize()
        ops.set_parallel_streams(False)       # single stream: the event brackets must see each kernel alone
        ops.kernel_timing(True)
        for _ in range(ksteps):
            flush.fill_(1)
            # park the GPU (~40 ms spin) while the host enqueues the whole step, so the events measure
            # back-to-back device execution instead of Python launch gaps
            torch.cuda._sleep(80_000_000)
            local_step()
        torch.cuda.synchronize()
        ktimes = ops.kernel_times_ms()
        ops.kernel_timing(False)
        ops.set_parallel_streams(True)
    if world > 1:
        dist.barrier()

    if rank == 0:
        peaks = load_peaks()
        ms_per_step = total_ms / args.steps
        value = B * world * args.steps / (total_ms / 1e3)
        line = {"metric": metric_name(WORKLOAD, cfg, B) + (f" [ABLATION {ablation}: not a bench line]" if ablation else ""), "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": scaling_of(WORKLOAD),
                "vs_baseline": None,
                "dtype": ("bf16 (single-pass tcgen05 MMAs, fp32 accumulate, on the token-row GEMMs; bf16x3 on query rows / heads)"
                          if throughput else "fp32" if args.precision == "fp32" else f"fp32 ({args.precision} tensor-core GEMMs)"),
                "data": "synthetic",
                "config": config_dict(cfg, {"global_batch": B * world, "parallelism": f"dp{world}",
                                            "grad_allreduce_bytes": reducer.nbytes() if world > 1 else 0,
                                            "grad_allreduce": ("none (1 GPU)" if world == 1 else
                                                               "NCCL AVG, per-segment, inside the replayed graph, overlapped with backward"
                                                               if (graphed is not None and graphed.reduces) else "NCCL AVG, one blocking call after the step"),
                                            "gemm_precision": args.precision,
                                            "dropout": args.dropout,
                                            "launch": "one CUDA graph per step" if args.graph else "eager",
                                            "micro_batches": args.micro_batches,
                                            "entry": ("backbone feature maps (input_proj inside the step)" if args.from_features
                                                      else "post-input_proj pyramid"),
                                            "loss": ("on-device PoseCriterion (SetCriterion + 'gt' matcher), synthetic targets" if args.criterion
                                                     else "fixed-cotangent loss (SURVEY.md section 8d)"),
                                            "optimizer": "fused clip_grad_norm_(0.1) + AdamW inside every step" if opt is not None else "none (forward + backward [+ all-reduce])",
                                            "kernel_table": ("eager single-stream pass, CUDA events around every library call; `share` is "
                                                             "of the table's own sum (the replayed graph overlaps branches on "
                                                             "side streams: kernel_table_ms / ms_per_step = serial-to-graph ratio)")}),
                "e2e": {"value": B * world * args.steps / (e2e_ms / 1e3), "unit": "images/s",
                        "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "pipeline": ("pinned host inputs of step i+1 copied on a copy stream while step i replays "
                                     "(GraphedStep.prefetch); result read back every step") if graphed is not None else "serial",
                        "serial_value": B * world * args.steps / (e2e_serial_ms / 1e3)},
                "gpu_launches": launches, "clocks": clocks}
        if parity is not None:
            line["parity"] = parity
        if ktimes:
            line["roofline"], line["kernels"] = roofline_from(ktimes, peaks, ksteps, ms_per_step,
                                                              mma_passes=3 if args.precision == "bf16x3" else 1,
                                                              msda_sparse_fraction=msda_sparse_fraction(cfg))
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(cfg, backward=do_backward)
        print(json.dumps(line), flush=True)
    if world > 1:
        graphed = None                      # a captured graph may hold NCCL nodes: release it before the communicator
        torch.cuda.synchronize()
        dist.destroy_process_group()


def msda_sparse_fraction(cfg):
    """Share of the encoder MSDA backward's sampling points whose grad_value contributions go through L2 atomics: the tile
    kernel (csrc/msda.cu, D = 16 or 32, L = P = 4, at least two such levels) turns the trailing levels that hold <= 104
    pixels together into dense products."""
    from poet_b200 import synthetic as S
    pyr = S.pyramid_of(cfg)
    if cfg["d_model"] // cfg["nheads"] not in (16, 32) or len(pyr) != 4 or cfg["n_points"] != 4:
        return 1.0
    sizes = [h * w for h, w in pyr]
    for l0 in range(3):                           # at least two dense levels (POET_MSDA_TILE_MAX_LD = 2)
        if sum(sizes[l0:]) <= 104:
            return l0 / 4.0
    return 1.0


def roofline_from(ktimes, peaks, steps, ms_per_step, mma_passes=1, msda_sparse_fraction=1.0):
    """Per-kernel table from the CUDA-event brackets; `roofline` = the kernel with the largest time share.
    GEMMs are judged against the tensor pipe (sustained bf16 peak: the kernel is timed inside a long step),
    everything else against HBM.  Algorithmic bytes/flops per launch are the figures of DESIGN.md."""
    rows = []
    table_ms = sum(ms for ms, _n, _b, _f in ktimes.values()) / steps          # serialised library time per step
    for name, (ms, n, nbytes, flops) in ktimes.items():
        if ms <= 0:
            continue
        is_gemm = name.startswith("poet_gemm")
        if is_gemm:
            ach, peak, unit, bound = flops / (ms * 1e-3) / 1e12, peaks["tf_sust"], "TFLOP/s", "tensor"
        else:
            ach, peak, unit, bound = nbytes / (ms * 1e-3) / 1e9, peaks["hbm"], "GB/s", "hbm"
        row = {"kernel": name, "launches_per_step": n / steps, "ms_per_step": ms / steps,
               "share": ms / steps / table_ms, "bound": bound, "achieved": ach, "peak": peak, "unit": unit,
               "frac": ach / peak if nbytes or flops else None,
               "hbm_gbs": nbytes / (ms * 1e-3) / 1e9 if nbytes else None}
        if is_gemm and flops:
            # split-bf16 issues `mma_passes` tensor-core MMAs per fp32 product: tensor-pipe utilisation counts them all
            row["issued_tflops"] = ach * mma_passes
            row["tensor_pipe_frac"] = ach * mma_passes / peak
            row["hbm_frac"] = row["hbm_gbs"] / peaks["hbm"] if row["hbm_gbs"] else None
        rows.append(row)
    rows.sort(key=lambda r: -r["ms_per_step"])
    # group GEMM shapes into one dominant-kernel line as well
    gemm = [r for r in rows if r["kernel"].startswith("poet_gemm")]
    top = rows[0]
    traffic, traffic_src = None, None
    try:                                        # measured DRAM bytes per launch of that kernel (ncu --set full)
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            tj = json.load(fh)
        traffic = tj.get(top["kernel"])
        traffic_src = tj.get("_source")
    except Exception:
        pass
    roof = {"kernel": top["kernel"], "bound": top["bound"], "achieved": top["achieved"], "peak": top["peak"],
            "unit": top["unit"], "frac": top["frac"], "traffic": traffic, "traffic_source": traffic_src,
            "peak_source": peaks["source"], "share_of_table": top["share"],
            "kernel_table_ms": table_ms, "serial_to_graph_ratio": table_ms / ms_per_step}
    if gemm:
        roof["all_gemm_share_of_table"] = sum(r["share"] for r in gemm)
    if top["kernel"].startswith("poet_msda_bwd"):
        # What bounds the scatter side of this kernel: every bilinear corner of every sampling point
        # of a SPARSE level is added to grad_value by L2 atomics (red.global.add.v4.f32, 64 B per corner at D = 16); the
        # chip retires ~5.3 TB/s of such payload whatever the instruction form or layout (tools/red_micro.cu,
        # profiles/r02_red_micro.txt).  The dense (low-resolution) levels of the tile kernel leave through ~1 MB of
        # reductions.  `achieved` assumes every sparse-level corner in range (an upper bound on the payload).
        ms, n, _nbytes, flops = ktimes[top["kernel"]]
        payload = flops / 30.0 * 16.0 * msda_sparse_fraction            # 16-byte lane-ops x sparse share
        roof["alt"] = {"bound": "l2_atomic_payload", "achieved": payload / (ms * 1e-3) / 1e9, "peak": 5300.0,
                       "unit": "GB/s of fp32 reduction payload (upper bound: all corners in range)",
                       "frac": payload / (ms * 1e-3) / 1e9 / 5300.0, "sparse_level_share": msda_sparse_fraction,
                       "peak_source": "tools/red_micro.cu on B200: 1.68 GB of 64-byte corner reductions in 310-320 us"}
    return roof, rows[:24]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=os.environ.get("POET_GEMM_PRECISION", "bf16x3"), choices=["fp32", "bf16x3", "bf16"])
    ap.add_argument("--no-kernel-table", dest="kernel_table", action="store_false")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", dest="graph", action="store_false")
    ap.add_argument("--from-features", action="store_true",
                    help="start from three backbone feature maps [B,256,H_l,W_l]: input_proj (SURVEY.md 8f N1) is part of the step")
    ap.add_argument("--criterion", action="store_true",
                    help="back-propagate the on-device PoseCriterion (reference SetCriterion + 'gt' matcher) instead of the "
                         "fixed-cotangent loss of SURVEY.md section 8d")
    ap.add_argument("--optimizer", action="store_true",
                    help="include the fused clip_grad_norm_(0.1) + AdamW step in every step (training step of cfg4)")
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS),
                    help="BASELINE.json config to run (default cfg2, the configuration the metric is quoted on)")
    ap.add_argument("--overlap-allreduce", action="store_true",
                    help="multi-GPU: all-reduce the gradient arena segment by segment inside the replayed graph, overlapped with "
                         "backward (measured slower than the single call after the replay at 2 GPUs: profiles/r02_allreduce_overlap.txt)")
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch override (default: the workload's)")
    ap.add_argument("--dropout", type=float, default=0.0,
                    help="train-mode dropout probability (reference default 0.1, main.py:94); 0 = the parity configuration")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle parity check of the benchmarked path")
    ap.add_argument("--micro-batches", type=int, default=int(os.environ.get("POET_MICRO_BATCHES", "1")),
                    help="slices of the per-GPU batch issued on separate streams (decoder chain of one overlaps the encoder of another)")
    args = ap.parse_args()
    global WORKLOAD
    WORKLOAD = args.workload
    if args.impl == "reference":
        run_reference(args)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the poet_b200 path has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    run_gpu(args)


if __name__ == "__main__":
    main()

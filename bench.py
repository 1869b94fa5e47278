#!/usr/bin/env python
"""bench.py — Aria-25.3B bf16 prefill tokens/s on B200 (BASELINE.json metric), one process per GPU.

    python bench.py --gpus 1 --steps 10 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the reference's algorithm on the host CPU (oracle port)

Workload (BASELINE.json configs[1], SURVEY.md §8d cfg 2): random-init Aria-25.3B (seed 0), one synthetic
980x980 image (4900 patches -> 256 image tokens) + 512 random text tokens => T = 768 prefill tokens,
num_logits_to_keep=1, batch 1, no KV cache in.  A "step" = one AriaForConditionalGeneration.forward().

Printed JSON (rank 0, one line):
  value       whole-job prefill tokens/s with inputs already resident in HBM (CUDA events, max over ranks)
  e2e         same metric through the public API with HOST (pinned) buffers: H2D of pixel_values + input_ids and
              D2H of the logits inside the timed region, every step
  roofline    dominant kernel = fc1 grouped expert GEMM (+SwiGLU): HBM-bound at 72 rows/expert; algorithmic bytes
              (weights + A + out) / mean launch duration measured live with CUDA events inside the timed region
  cpu_baseline the oracle port (oracle/aria_oracle.py) timed on the host cores on a bounded sample
Multi-GPU: the prefill path of one request does not shard (ViT/attention single-GPU per north_star); --gpus N runs N
independent replicas (one request each, weak scaling, no data-path collective).  See DESIGN.md §5.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

T_TEXT, T_IMG = 512, 256
T_TOTAL = T_TEXT + T_IMG
METRIC = "Aria-25.3B bf16 prefill tokens/sec"
WORKLOAD = ("cfg2: Aria-25.3B, one 980px image (4900 patches -> 256 image tokens) + 512 text tokens, T=768 prefill, "
            "batch 1, num_logits_to_keep=1, random-init weights")


def _peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        self.idx = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ CPU oracle leg
class CpuReference:
    """Times the oracle port (the reference's algorithm, torch CPU bf16) on a bounded sample of the SAME workload:
    one full-width ViT layer (N=4900) and one full-width MoE decoder layer (T=768), extrapolated to 27 + 28 layers."""

    def __init__(self, threads=None):
        import torch
        from oracle import configs as C

        try:
            import psutil
            cores = psutil.cpu_count(logical=False) or os.cpu_count()
        except Exception:
            cores = os.cpu_count()
        self.threads = threads or cores
        torch.set_num_threads(self.threads)
        self.cfg = C.with_layers(C.ARIA_25B, lm_layers=1, vit_layers=1)
        gen = torch.Generator().manual_seed(0)
        sd = {}
        sd.update(C.vit_state(self.cfg["vision_config"], gen))
        sd.update(C.projector_state(self.cfg["projector"], gen))
        sd.update(C.lm_state(self.cfg["text_config"], gen))
        self.sd = {k: v.bfloat16() for k, v in sd.items()}
        self.pv = torch.randn(1, 3, 980, 980, generator=gen).bfloat16()
        self.emb = torch.randn(1, T_TOTAL, 2560, generator=gen).bfloat16()

    def sample(self):
        import torch
        import torch.nn.functional as F
        from oracle import aria_oracle as O

        cfg, sd = self.cfg, self.sd

        def timed(fn):
            t0 = time.perf_counter()
            r = fn()
            return time.perf_counter() - t0, r

        vp, lp = "vision_tower.vision_model.", "language_model.model.layers.0."
        tc = cfg["text_config"]
        pos = torch.arange(T_TOTAL)[None]
        with torch.no_grad():
            pm = torch.ones(1, 70, 70, dtype=torch.bool)
            t_emb, x = timed(lambda: O.vit_embeddings(self.pv, pm, sd, cfg["vision_config"], vp))
            t_vl, x = timed(lambda: O.vit_encoder_layer(x, sd, vp + "encoder.layers.0.", cfg["vision_config"], None))
            t_p, _ = timed(lambda: O.projector_forward(x, None, sd, cfg["projector"]))
            t_ll, (h, _) = timed(lambda: O.moe_decoder_layer(self.emb, sd, lp, tc, pos))
            t_head, _ = timed(lambda: F.linear(O.rms_norm(h[:, -1:], sd["language_model.model.norm.weight"], tc["rms_norm_eps"]),
                                               sd["language_model.lm_head.weight"]))
        vit_layer, lm_layer = t_vl, t_ll
        total = t_emb + 27 * vit_layer + t_p + 28 * lm_layer + t_head
        cpu_work = t_emb + t_vl + t_p + t_ll + t_head
        return {"value": T_TOTAL / total, "unit": "tokens/s", "cores": self.threads, "kind": "port",
                "sample": (f"oracle port (oracle/aria_oracle.py), full-width bf16 on the host CPU: ViT patch-embed + ONE "
                           f"encoder layer (N=4900), projector, ONE MoE decoder layer (T=768), final norm + lm_head (1 row); "
                           f"total = embed + 27 x vit_layer + projector + 28 x lm_layer + head; vit_layer={vit_layer:.3f}s "
                           f"lm_layer={lm_layer:.3f}s cpu_work={cpu_work:.1f}s/sample")}


def run_reference(args, rank):
    if rank != 0:
        return
    ref = CpuReference()
    for _ in range(args.warmup):
        ref.sample()
    vals = [ref.sample() for _ in range(args.steps)]
    v = statistics.median([x["value"] for x in vals])
    last = vals[-1]
    line = {"metric": METRIC, "value": v, "unit": "tokens/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1000.0 * T_TOTAL / v, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic", "impl": "reference",
            "config": {"workload": WORKLOAD, "global_batch": 1, "seq_len": T_TOTAL, "parallelism": "host CPU"},
            "cpu_baseline": {"value": v, "unit": "tokens/s", "cores": last["cores"], "kind": "port", "sample": last["sample"]},
            "e2e": {"value": v, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ GPU leg
def run_aria(args, rank, local_rank, world):
    import torch

    from aria_b200 import _lib as L
    from aria_b200 import ops
    from aria_b200.modeling_aria import AriaConfig, AriaForConditionalGeneration, GraphedPrefill, init_random_
    from aria_b200 import configs as C

    L.load()  # fail loudly if the CUDA extension is missing
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    cfg = C.ARIA_25B
    torch.set_grad_enabled(False)
    model = AriaForConditionalGeneration(AriaConfig.from_dict(cfg), device=dev)
    init_random_(model, seed=0)
    n_params = sum(p.numel() for p in model.parameters())
    # N > 1: every rank prefills its own request (data parallel).  ARIA_BENCH_MULTI=ep additionally shards the routed
    # experts over the ranks (token rows exchanged by our NVLink peer-memory kernels): each GPU then streams 1/N of the
    # expert weights per layer.  Default "replicas": N independent model replicas, no data-path collective.
    multi = os.environ.get("ARIA_BENCH_MULTI", "replicas") if world > 1 else "single"
    if multi == "ep":
        model.enable_expert_parallel(T_TOTAL)

    g = torch.Generator().manual_seed(1234 + rank)
    pv_host = torch.randn(1, 3, 980, 980, generator=g).bfloat16().pin_memory()
    text = torch.randint(10, cfg["text_config"]["vocab_size"], (T_TEXT,), generator=g)
    ids_host = torch.cat([text[:16], torch.full((T_IMG,), cfg["image_token_index"]), text[16:]])[None].contiguous().pin_memory()
    logits_host = torch.empty(1, 1, cfg["text_config"]["vocab_size"], dtype=torch.bfloat16).pin_memory()
    pv_dev, ids_dev = pv_host.to(dev), ids_host.to(dev)

    # live timing of the dominant kernel: fc1 grouped GEMM + SwiGLU (one launch per MoE layer)
    rec = {"on": False, "ev": []}
    orig_gg = ops.grouped_gemm

    def timed_gg(a, b, off, swiglu=False, **kw):
        if rec["on"] and swiglu:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            y = orig_gg(a, b, off, swiglu=swiglu, **kw)
            e1.record()
            rec["ev"].append((e0, e1, a.shape[0]))
            return y
        return orig_gg(a, b, off, swiglu=swiglu, **kw)

    ops.grouped_gemm = timed_gg

    # public API: eager forward() for the instrumented leg, GraphedPrefill (CUDA-graph replay of the same forward)
    # for the throughput legs
    def step_eager():
        return model(ids_dev, pv_dev, None, num_logits_to_keep=1, input_ids_host=ids_host).logits

    graphed = GraphedPrefill(model, ids_host, pv_host, num_logits_to_keep=1)

    def step_resident():
        return graphed.replay()  # inputs already resident in HBM

    def step_e2e():
        out = graphed(ids_host, pv_host)            # H2D of ids + pixels from pinned host memory, then replay
        logits_host.copy_(out, non_blocking=False)  # D2H read of the step's result (synchronises)
        return logits_host

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step_resident()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    # kernels per step (counted on one eager forward; the graph replays exactly these launches)
    l0 = L.launch_count
    step_eager()
    launches_per_step = L.launch_count - l0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step_resident()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1) / args.steps
    launches = launches_per_step * args.steps
    # dominant kernel, timed live with CUDA events on the launching stream over the same K steps (eager launches of
    # the identical kernels; events cannot be recorded inside a graph replay)
    rec["on"] = True
    x0, x1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    x0.record()
    for _ in range(args.steps):
        step_eager()
    x1.record()
    barrier()
    rec["on"] = False
    ms_eager = x0.elapsed_time(x1) / args.steps
    fc1_ms = [a.elapsed_time(b) for a, b, _ in rec["ev"]]
    fc1_rows = rec["ev"][0][2] if rec["ev"] else 0

    # end-to-end leg (host buffers, copies inside the timed region)
    for _ in range(2):
        step_e2e()
    barrier()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(args.steps):
        step_e2e()
    t1.record()
    barrier()
    ms_e2e = t0.elapsed_time(t1) / args.steps
    clocks = sampler.stop()

    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([ms, ms_e2e], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = float(t[0]), float(t[1])
    if rank != 0:
        return
    peaks, peak_src = _peaks()
    tc = cfg["text_config"]
    E, d, I = tc["moe_num_experts"], tc["hidden_size"], tc["moe_intermediate_size"]
    if multi == "ep":  # each rank streams its E/world experts; rows = this rank's share of all ranks' (token, slot) pairs
        E = E // world
        fc1_rows = T_TOTAL * tc["moe_topk"]
    fc1_bytes = E * d * 2 * I * 2 + fc1_rows * d * 2 + fc1_rows * I * 2  # weights + A read + out write
    fc1_avg = statistics.mean(fc1_ms) if fc1_ms else float("nan")
    achieved = fc1_bytes / (fc1_avg * 1e-3) / 1e9
    try:  # DRAM traffic of the same kernel from the committed `ncu --set full` capture (per launch)
        tr = json.load(open(os.path.join(ROOT, "profiles", "fc1_traffic.json")))
        traffic = tr["dram_bytes_read"] + tr["dram_bytes_write"]
    except Exception:
        traffic = None
    roofline = {"bound": "hbm", "kernel": "gemm_kernel<128,MN-major,SWIGLU> (fc1 grouped expert GEMM)",
                "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
                "traffic": traffic, "peak_source": peak_src, "bytes_per_launch": fc1_bytes,
                "avg_launch_ms": fc1_avg, "launches_timed": len(fc1_ms),
                "share_of_step": sum(fc1_ms) / args.steps / ms_eager if fc1_ms else None,
                "eager_ms_per_step": ms_eager}
    line = {"metric": METRIC, "value": world * T_TOTAL / (ms * 1e-3), "unit": "tokens/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": WORKLOAD, "global_batch": world, "seq_len": T_TOTAL,
                       "parallelism": (f"dp{world} + experts sharded ep{world} (NVLink peer-memory token exchange)" if multi == "ep"
                                       else f"replicas x{world} (no data-path collective)"), "params": n_params,
                       "l2": "per-step working set = 50.6 GB of weights >> 126 MB L2, no flush needed"},
            "e2e": {"value": world * T_TOTAL / (ms_e2e * 1e-3), "unit": "tokens/s", "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": pv_host.numel() * 2 + ids_host.numel() * 8,
                    "d2h_bytes_per_step": logits_host.numel() * 2},
            "gpu_launches": launches, "launch_mode": "CUDA graph replay of the eager forward (same kernels)",
            "clocks": clocks, "roofline": roofline, "impl": "aria_b200"}
    if world == 1 and not args.no_cpu_baseline:
        ref = CpuReference()
        ref.sample()  # warm-up
        line["cpu_baseline"] = ref.sample()
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="aria", choices=["aria", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
    else:
        run_aria(args, rank, local_rank, world)
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()

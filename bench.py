#!/usr/bin/env python
"""bench.py -- MAG-BERT samples/s (seq=50, bs=48 per GPU) on N B200s, BASELINE.json's metric.

    python bench.py --gpus 1 --steps 200 --warmup 10                 # our arm
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W       # N GPUs, one rank per GPU (NCCL)
    python bench.py --impl reference --gpus 1 --steps K --warmup W   # reference arm: the CPU oracle port

A "step" is one optimisation step (forward + MSE + backward + AdamW, dropout on) of
MAG_BertForSequenceClassification (bert-base geometry, random init, synthetic MOSI-shaped batch:
seq 50, V 47, A 74, bs 48 per GPU) -- the unit the reference's train_epoch executes
(multimodal_driver.py:354-388).  One JSON line is printed by rank 0:

  value      whole-job samples/s with the batch already resident in HBM (CUDA-graph replay, no copies)
  e2e        the same through the public API with HOST batches: pinned-host -> H2D -> step -> loss D2H
  roofline   the dominant kernel (tcgen05 GEMM, FFN-in shape) timed alone with CUDA events vs the measured
             bf16/fp16 dense peak (MEASURED_PEAKS.json), plus per-shape entries and the whole-step fraction
  cpu_baseline  oracle/restatement.py (torch fp32 on the host cores) on a bounded sample of the same workload
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEQ, BS, V, A = 50, 48, 47, 74
FWD_FLOP_PER_SAMPLE = 8.7234e9            # SURVEY.md section 8d (S=50, MOSI), algorithmic, padding not counted
TRAIN_FLOP_PER_SAMPLE = 3 * FWD_FLOP_PER_SAMPLE
METRIC = "MAG-BERT train-step samples/sec (seq=50, bs=48/GPU)"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            d = json.load(fh)
        return dict(hbm_gbs=d["hbm_gbs"], tflops_burst=d["bf16_tflops"], tflops_sustained=d["bf16_tflops_sustained"],
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm_gbs=6650.0, tflops_burst=1590.0, tflops_sustained=1400.0, source="fallback (B200_PROFILING.md)")


# ------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples SM clock + throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {getattr(nv, k): k for k in dir(nv) if k.startswith("nvmlClocksEventReason") or
                 k.startswith("nvmlClocksThrottleReason")}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if isinstance(bit, int) and bit and (r & bit) == bit and "None" not in name and "All" not in name:
                        short = name.replace("nvmlClocksEventReason", "").replace("nvmlClocksThrottleReason", "")
                        if short not in ("GpuIdle", "ApplicationsClocksSetting"):
                            self.reasons.add(short)
            except Exception:
                pass
            time.sleep(0.01)

    def __enter__(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thr is not None:
            self._thr.join(timeout=1.0)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ------------------------------------------------------------------------------------------- CPU arm
def cpu_oracle_throughput(train_samples_budget, batch, steps, warmup, mode="train"):
    """Times oracle/restatement.py on the host cores.  Returns (samples/s, cores, description)."""
    import torch
    from oracle import inputs as gen, restatement as R, weights
    sd = weights.bert_state_dict(seed=0)
    cores = pick_cpu_threads(torch, R, gen, sd)
    b = gen.bert_batch(batch, SEQ, V, A, seed=1234)
    if mode == "train":
        params = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
        plist = list(params.values())
        opt = torch.optim.AdamW(plist, lr=1e-5, eps=1e-6, weight_decay=0.01)
        gens = torch.Generator().manual_seed(0)

        def drop(site, x):   # the reference trains with dropout on (p = 0.1 / 0.5 for MAG)
            p = 0.5 if site == "mag" else 0.1
            return torch.nn.functional.dropout(x, p, True)

        def step():
            opt.zero_grad(set_to_none=True)
            logits = R.mag_bert_forward(params, b["input_ids"], b["visual"], b["acoustic"], b["attention_mask"],
                                        b["token_type_ids"], drop=drop)
            loss = R.mse_loss(logits, b["labels"])
            loss.backward()
            opt.step()
            return float(loss.detach())
    else:
        def step():
            with torch.no_grad():
                R.mag_bert_forward(sd, b["input_ids"], b["visual"], b["acoustic"], b["attention_mask"],
                                   b["token_type_ids"])
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return batch * steps / dt, cores, dt / steps


def pick_cpu_threads(torch, R, gen, sd):
    """The CPU arm should be as fast as the host allows: torch's intra-op pool does not scale to every
    hardware thread on small matrices, so time one eval forward of 8 sequences at a few thread counts and
    keep the best (all cores are always among the candidates)."""
    ncpu = os.cpu_count() or 1
    cands = sorted({c for c in (8, 16, 32, 64, ncpu) if c <= ncpu} | {ncpu})
    b = gen.bert_batch(8, SEQ, V, A, seed=7)
    best, best_t = ncpu, float("inf")
    for c in cands:
        torch.set_num_threads(c)
        with torch.no_grad():
            R.mag_bert_forward(sd, b["input_ids"], b["visual"], b["acoustic"], b["attention_mask"], b["token_type_ids"])
            t0 = time.perf_counter()
            R.mag_bert_forward(sd, b["input_ids"], b["visual"], b["acoustic"], b["attention_mask"], b["token_type_ids"])
            dt = time.perf_counter() - t0
        if dt < best_t:
            best, best_t = c, dt
    torch.set_num_threads(best)
    return best


def cpu_model_name():
    try:
        with open("/proc/cpuinfo") as fh:
            for line in fh:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # bounded sample: keep the whole run within a few minutes of CPU time
    per_step = max(1, min(BS, int(1500 / max(1, args.steps + args.warmup))))
    sps, cores, sec = cpu_oracle_throughput(0, per_step, args.steps, args.warmup, "train")
    line = {
        "impl": "reference", "metric": METRIC, "value": sps, "unit": "samples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"MAG-BERT bert-base MOSI-synthetic train step (fwd+MSE+bwd+AdamW, dropout on) seq={SEQ} "
                               f"V={V} A={A}; CPU sample of {per_step} sequences per step", "global_batch": per_step,
                   "seq_len": SEQ, "parallelism": "cpu"},
        "cpu_baseline": {"value": sps, "unit": "samples/s", "cores": cores, "kind": "port",
                         "sample": f"{args.steps} timed train steps of {per_step} sequences (oracle/restatement.py, torch "
                                   f"fp32, {cores} threads, {cpu_model_name()})"},
        "e2e": {"value": sps, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------- GPU arm
def time_gemm_shapes(torch, L, lib, peaks, T):
    """The tcgen05 GEMM alone on the step's four forward shapes (CUDA events, back-to-back launches over
    rotating buffers larger than L2)."""
    shapes = [("ffn_in", T, 3072, 768, L.EPI_GELU_F16), ("ffn_out", T, 768, 3072, L.EPI_STORE_F32),
              ("qkv", T, 2304, 768, L.EPI_STORE_F16), ("attn_out", T, 768, 768, L.EPI_STORE_F32)]
    out = []
    st = L.current_stream()
    for name, M, N, K, epi in shapes:
        nbuf = 12   # 12 x (A + W + out) > 126 MB L2 for every shape
        As = [torch.randn(M, K, device="cuda").half() for _ in range(nbuf)]
        Ws = [(torch.randn(N, K, device="cuda") * 0.02).half() for _ in range(nbuf)]
        bias = torch.zeros(N, device="cuda")
        o_dt = torch.float32 if epi == L.EPI_STORE_F32 else torch.float16
        Os = [torch.empty(M, N, device="cuda", dtype=o_dt) for _ in range(nbuf)]
        O1 = torch.empty(M, N, device="cuda", dtype=torch.float16)

        def launch(i):
            j = i % nbuf
            L.check(lib.magb200_gemm(L.ptr(As[j]), L.ptr(Ws[j]), M, N, K, K, K, L.F16, L.F16, 0, 0, epi, L.ptr(bias),
                                     L.ptr(Os[j]), L.ptr(O1) if epi == L.EPI_GELU_F16 else None, None, N, 1.0, 1, 0, 0,
                                     st))
        for i in range(nbuf):
            launch(i)
        torch.cuda.synchronize()
        reps = 60
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(reps):
            launch(i)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / reps
        flops = 2.0 * M * N * K
        tf = flops / (us * 1e-6) / 1e12
        # forward Linears run on the persistent cta_group::2 kernel (profiles/r01_gemm_policy.txt)
        kern = "gemm2p_tcgen05_kernel<256x%d pair tiles>" % (256 if (N % 256 == 0 and N >= 1024) else 128)
        out.append({"kernel": kern, "shape": name, "M": M, "N": N, "K": K, "us": round(us, 2),
                    "achieved": round(tf, 1), "frac": round(tf / peaks["tflops_burst"], 4)})
    return out


def read_profile_traffic():
    p = os.path.join(ROOT, "profiles", "latest_traffic.json")
    if os.path.exists(p):
        try:
            with open(p) as fh:
                return json.load(fh).get("ffn_in_dram_bytes_per_launch")
        except Exception:
            return None
    return None


def run_gpu_arm(args):
    import torch
    import torch.distributed as dist
    from transformers import BertConfig

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback (use --impl reference for "
                         "the CPU oracle arm)")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import bert as mag_bert
    from bert_multimodal_transformer_b200 import _lib as L
    from bert_multimodal_transformer_b200.trainer import MagBertTrainer
    from oracle import inputs as gen

    lib = L.load()
    L.check(lib.magb200_device_check())
    peaks = load_peaks()

    class MC:
        beta_shift, dropout_prob = 1.0, 0.5

    torch.manual_seed(0)
    cfg = BertConfig(num_labels=1)
    import io, contextlib
    with contextlib.redirect_stdout(io.StringIO()):
        model = mag_bert.MAG_BertForSequenceClassification(cfg, MC())
    model = model.to(torch.device("cuda", local)).train()
    total_steps = args.steps * 2 + args.warmup * 2 + 64
    trainer = MagBertTrainer(model, BS, SEQ, lr=1e-5, weight_decay=0.01, warmup_steps=total_steps // 10,
                             total_steps=total_steps, use_cuda_graph=not args.no_graph, seed=rank)
    batch = gen.bert_batch(BS, SEQ, V, A, seed=1234 + rank)
    trainer._stage(batch["input_ids"], batch["visual"], batch["acoustic"], batch["attention_mask"],
                   batch["token_type_ids"], batch["labels"])
    trainer.upload_staged()     # "value" runs on this HBM-resident batch

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # launches per step (counted by the library while the step is recorded / run once)
    c0 = lib.magb200_launch_count()
    trainer.launch_train_step(transfers=False)
    trainer.stream.synchronize()
    # graph capture ran the body twice (warm-up + capture); eager mode once
    launches_per_step = (lib.magb200_launch_count() - c0) // (2 if trainer.use_graph else 1)

    # ---------------- value: inputs resident in HBM, CUDA events on the launching stream
    for _ in range(max(args.warmup, 3)):
        trainer.launch_train_step(transfers=False)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        with torch.cuda.stream(trainer.stream):
            e0.record()
        for _ in range(args.steps):
            trainer.launch_train_step(transfers=False)
        with torch.cuda.stream(trainer.stream):
            e1.record()
        trainer.stream.synchronize()
    barrier()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = world * BS * args.steps / (ms * 1e-3)
    loss_after = float(trainer.loss.item())
    if args.profile_mode:
        if rank == 0:
            print(json.dumps({"profile_mode": True, "value_under_profiler": value, "ms_per_step": ms / args.steps,
                              "gpu_launches_per_step": int(launches_per_step)}), flush=True)
        if world > 1:
            finish_distributed(trainer, dist)
        return

    # ---------------- e2e: host batch -> pinned -> H2D -> step -> loss D2H, every step, through train_step()
    for _ in range(3):
        trainer.train_step(batch["input_ids"], batch["visual"], batch["acoustic"], batch["attention_mask"],
                           batch["token_type_ids"], batch["labels"])
    barrier()
    t0 = time.perf_counter()
    with torch.cuda.stream(trainer.stream):
        e0.record()
    # every step stages its own host batch into pinned memory, copies it H2D inside the step and reads its
    # loss back D2H; the host stages batch k+1 while the GPU runs step k (submit/result, two pinned slots)
    prev = None
    for _ in range(args.steps):
        tk = trainer.submit_train_step(batch["input_ids"], batch["visual"], batch["acoustic"], batch["attention_mask"],
                                       batch["token_type_ids"], batch["labels"])
        if prev is not None:
            last_loss = trainer.result(prev)
        prev = tk
    last_loss = trainer.result(prev)
    with torch.cuda.stream(trainer.stream):
        e1.record()
    trainer.stream.synchronize()
    wall = time.perf_counter() - t0
    ms_e2e = max(e0.elapsed_time(e1), wall * 1e3)
    if world > 1:
        t = torch.tensor([ms_e2e], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_e2e = float(t.item())
    e2e_value = world * BS * args.steps / (ms_e2e * 1e-3)

    # ---------------- eval forward (secondary number)
    for _ in range(3):
        trainer.launch_eval_step(transfers=False)
    barrier()
    n_eval = max(20, args.steps)
    with torch.cuda.stream(trainer.stream):
        e0.record()
    for _ in range(n_eval):
        trainer.launch_eval_step(transfers=False)
    with torch.cuda.stream(trainer.stream):
        e1.record()
    trainer.stream.synchronize()
    ms_eval = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms_eval], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_eval = float(t.item())
    eval_value = world * BS * n_eval / (ms_eval * 1e-3)

    if rank != 0:
        if world > 1:
            finish_distributed(trainer, dist)
        return

    # ---------------- rank 0 only: dominant-kernel roofline + CPU baseline
    torch.cuda.set_stream(torch.cuda.default_stream())
    kernels = time_gemm_shapes(torch, L, lib, peaks, BS * SEQ)
    dom = kernels[0]
    roofline = {
        "bound": "tensor", "kernel": "gemm2p_tcgen05_kernel<fp16, K-major, persistent cta_group::2, 256x256 pair tiles> "
                                     "(FFN-in: M=2400 N=3072 K=768, bias + GELU + saved GELU' epilogue)",
        "achieved": dom["achieved"], "peak": peaks["tflops_burst"], "unit": "TFLOP/s", "frac": dom["frac"],
        "traffic": read_profile_traffic(), "peak_source": peaks["source"] + ", burst figure (kernel timed alone)",
        "per_shape": kernels,
        "whole_step": {"flop_per_sample": TRAIN_FLOP_PER_SAMPLE,
                       "achieved_tflops": round(value / world * TRAIN_FLOP_PER_SAMPLE / 1e12, 2),
                       "frac_of_sustained_peak": round(value / world * TRAIN_FLOP_PER_SAMPLE / 1e12
                                                       / peaks["tflops_sustained"], 4)},
    }
    cpu = None
    if not args.no_cpu_baseline:
        sps, cores, sec = cpu_oracle_throughput(0, 16, 3, 1, "train")
        cpu = {"value": round(sps, 2), "unit": "samples/s", "cores": cores, "kind": "port",
               "sample": f"3 timed + 1 warm-up train steps of 16 sequences (oracle/restatement.py, torch fp32 on "
                         f"{cores} threads of {cpu_model_name()})"}
    line = {
        "metric": METRIC, "value": round(value, 1), "unit": "samples/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": round(ms / args.steps, 4), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None,
        "dtype": "fp16 tensor-core operands; fp32 accumulate / residual stream / LayerNorm / softmax / master weights",
        "data": "synthetic",
        "config": {"workload": f"MAG-BERT bert-base-uncased geometry, MOSI-synthetic (seq={SEQ}, V={V}, A={A}), "
                               f"bs={BS}/GPU, train step = fwd + MSE + bwd + AdamW, dropout on (BASELINE.json configs[1])",
                   "global_batch": BS * world, "seq_len": SEQ, "parallelism": f"dp{world}",
                   "l2": "no flush needed: each step streams ~2.5 GB (params, grads, Adam moments, activations) "
                         ">> 126 MB L2",
                   "cuda_graph": trainer.use_graph},
        "clocks": clk.summary(),
        "e2e": {"value": round(e2e_value, 1), "unit": "samples/s", "h2d_bytes_per_step": trainer.h2d_bytes,
                "d2h_bytes_per_step": trainer.d2h_bytes, "ms_per_step": round(ms_e2e / args.steps, 4),
                "api": "MagBertTrainer.submit_train_step(host tensors) / result(ticket) -> loss float, 2 pinned "
                       "staging slots (host stages batch k+1 while the GPU runs step k)"},
        "gpu_launches": int(launches_per_step * args.steps),
        "gpu_launches_per_step": int(launches_per_step),
        "eval_forward": {"value": round(eval_value, 1), "unit": "samples/s", "ms_per_step": round(ms_eval / n_eval, 4),
                         "roofline_frac_of_sustained": round(eval_value / world * FWD_FLOP_PER_SAMPLE / 1e12
                                                             / peaks["tflops_sustained"], 4)},
        "roofline": roofline,
        "cpu_baseline": cpu,
        "loss_after": loss_after, "last_loss": last_loss,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        finish_distributed(trainer, dist)


def finish_distributed(trainer, dist):
    """Orderly exit of a multi-rank run: drop the captured graphs (they reference NCCL work), meet at a barrier,
    then leave without tearing NCCL down from under them (destroy_process_group can block on captured
    collectives); the process exit releases everything."""
    import torch
    trainer.release()
    dist.barrier()
    torch.cuda.synchronize()
    sys.stdout.flush()
    sys.stderr.flush()
    os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-graph", action="store_true", help="eager launches instead of CUDA-graph replay")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-mode", action="store_true",
                    help="only run the warm-up + timed train steps (for ncu captures); prints a reduced line")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()

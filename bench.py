#!/usr/bin/env python
"""
bench.py -- raw-signal samples/sec basecalled (forward + decode) on synthetic chunks.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1]): hac-shaped LSTM-CRF (H=384, 5-mer scores, seeded random weights -- the
real checkpoint needs the network), batch 512 chunks per GPU, 10 000-sample chunks trimmed to a stride multiple
(9996) exactly as `_load_model(use_koi=True)` does (bonito/util.py:288-291).  A step = one batch through
conv stem -> strided conv GEMM -> 5 x (input GEMM + persistent LSTM) -> CRF linear + clamp -> CRF decode.

  value  whole-job samples/s with the batch already resident in HBM (CUDA events, max over ranks)
  e2e    the same through the reference-facing call `compute_scores(model, batch)` with HOST float32 input:
         fp16 cast into pinned memory, H2D, forward, decode, D2H of moves/sequence/qstring inside the timing
  roofline / stages: per-kernel CUDA-event durations recorded inside the timed region

Multi-GPU: chunks shard by batch (one process per GPU, weights broadcast once over NCCL, no steady-state
collective) => weak scaling.
--impl reference: the reference's PyTorch-CPU execution of the same path (oracle/cpu_reference.py) on the host
cores, a bounded sample of the workload per step.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

# The engine drives one CUDA stream per 32-chunk tile (16 at batch 512).  With the default of 8 hardware work queues,
# streams that share a queue serialise behind each other's not-yet-dispatched cluster launches; this must be set
# before the CUDA context exists.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "raw-signal samples/sec basecalled (forward+decode), 10k-sample chunks, hac-shaped LSTM-CRF"
CHUNK = 10000
BATCH = 512
MODEL = "hac"


def log(msg):
    print(f"[bench {time.strftime('%H:%M:%S')}] {msg}", file=sys.stderr, flush=True)


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            p = json.load(fh)
        return dict(hbm=p["hbm_gbs"], tflops=p["bf16_tflops_sustained"], tflops_burst=p["bf16_tflops"], source="measured")
    except Exception:
        return dict(hbm=6650.0, tflops=1400.0, tflops_burst=1590.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms while the timed region runs."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index
        self.t_begin = self.t_end = None

    def mark_begin(self):
        self.t_begin = time.time()

    def mark_end(self):
        self.t_end = time.time()

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.thread.join(timeout=2)
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for stamp, r in self.rows:
            if self.t_begin is not None and not (self.t_begin <= stamp <= (self.t_end or stamp) + 0.06):
                continue   # only samples taken while the timed region ran
            try:
                pw.append(float(r[2]))
                sm.append(float(r[0])); mx.append(float(r[1]))
                for name, flag in zip(names, r[3:7]):
                    if flag.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "sm_mhz_min": sm[0] if sm else None, "power_w_max": max(pw) if pw else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def build_model(device, rank, world):
    from bonito_b200.crf.model import Model
    from bonito_b200 import synth
    spec = synth.model_spec(MODEL)
    weights = synth.make_weights(spec, seed=25)
    cfg = synth.model_config(spec, batchsize=BATCH, chunksize=CHUNK, overlap=500)
    model = Model(cfg)
    if rank == 0:
        model.load_state_dict(synth.state_dict_from_weights(spec, weights))
    chunksize = CHUNK - CHUNK % model.stride
    model.use_koi(batchsize=BATCH, chunksize=chunksize, quantize=False)
    model = model.half().eval().to(device)
    if world > 1:  # the one collective of the path: weights from rank 0 (NCCL over NVLink)
        from bonito_b200.distributed import broadcast_parameters
        broadcast_parameters(model, src=0)
    return model, spec, weights, chunksize


def cpu_baseline(spec, weights, chunksize, n_chunks=32, threads=None):
    from oracle import synth
    from oracle.cpu_reference import CpuReferenceModel
    threads = threads or min(os.cpu_count(), 32)
    torch.set_num_threads(threads)
    os.environ["OMP_NUM_THREADS"] = str(threads)
    log(f"cpu baseline: {n_chunks} chunks on {threads} threads (host has {os.cpu_count()})")
    ref = CpuReferenceModel(spec, weights)
    x = synth.squiggle(n_chunks, chunksize, seed=25)
    ref.forward(x[:2])  # warm-up
    log("cpu baseline: warm-up done")
    _, _, _, tf, td = ref.basecall_batch(x)
    log(f"cpu baseline: forward {tf:.2f}s decode {td:.2f}s")
    samples = n_chunks * chunksize
    return {"value": samples / (tf + td), "unit": "samples/s", "cores": threads, "kind": "port",
            "forward_only": samples / tf,
            "sample": f"{n_chunks} chunks x {chunksize} samples, torch {torch.__version__} fp32 modules as bonito.nn builds "
                      f"them + OpenMP C posterior-Viterbi decode (forward {tf:.2f}s, decode {td:.2f}s)"}


def run_reference(args, rank, world):
    """Reference arm: PyTorch-CPU path on the host cores; rank 0 only."""
    if rank != 0:
        return
    from oracle import synth
    spec = synth.model_spec(MODEL)
    weights = synth.make_weights(spec, seed=25)
    chunksize = CHUNK - CHUNK % 6
    from oracle.cpu_reference import CpuReferenceModel
    threads = min(os.cpu_count(), 32)
    torch.set_num_threads(threads)
    os.environ["OMP_NUM_THREADS"] = str(threads)
    ref = CpuReferenceModel(spec, weights)
    n_chunks = 32
    x = synth.squiggle(n_chunks, chunksize, seed=25)
    for _ in range(max(args.warmup, 1)):
        ref.forward(x[:2])
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ref.basecall_batch(x)
    dt = time.perf_counter() - t0
    value = n_chunks * chunksize * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{MODEL}-shaped LSTM-CRF, {chunksize}-sample chunks, forward+decode",
                   "sample_chunks_per_step": n_chunks},
        "cpu_baseline": {"value": value, "unit": "samples/s", "cores": threads, "kind": "port",
                         "sample": f"{n_chunks} chunks x {chunksize} samples per step"},
        "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--decode-overlap", action="store_true",
                    help="enqueue each tile's decode on its stream (measured slower: the decode CTAs share SMs with the "
                         "latency-critical recurrent clusters)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch.distributed as dist
    from bonito_b200.crf.basecall import compute_scores
    from bonito_b200.decode import _decoder
    from bonito_b200 import synth

    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)
    peaks = load_peaks()
    model, spec, weights, chunksize = build_model(device, rank, world)
    N, L = args.batch, chunksize
    host_batch = synth.squiggle(64, L, seed=100 + rank).repeat(N // 64 + 1, 1, 1)[:N].contiguous()  # float32 host
    x_dev = host_batch.to(device, torch.float16)
    log("model built")
    plan = model.native_plan(device)
    T = plan.frames(L)
    qs = model.config["qscore"]

    def step_resident(events=None):
        scores = plan.forward(x_dev, events=events, decode=(qs["scale"], qs["bias"]) if args.decode_overlap else None)
        return _decoder(scores, spec["state_len"], blank_score=plan.blank_score, qscale=qs["scale"], qbias=qs["bias"],
                        events=events)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.inference_mode():
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()          # nvidia-smi needs a moment before its first sample: start it before the warm-up
        for _ in range(max(args.warmup, 3)):
            step_resident()
        barrier()
        log("warm-up done")
        sampler.mark_begin()
        events = []
        t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_start.record()
        host_t0 = time.perf_counter()
        for _ in range(args.steps):
            step_resident(events)
        host_enqueue_ms = (time.perf_counter() - host_t0) * 1e3 / args.steps
        t_end.record()
        barrier()
        sampler.mark_end()
        elapsed_ms = t_start.elapsed_time(t_end)
        log(f"resident: {elapsed_ms / args.steps:.2f} ms/step")
        clocks = sampler.stop() if rank == 0 else None

        # end to end through the reference-facing calls, host float32 batches in, host byte arrays out:
        #   score_batches  the generator basecall() runs (staging / H2D of batch k+1 and D2H of k-1 overlap the kernels of k)
        #   compute_scores one synchronous call per batch (reported next to it)
        from bonito_b200.crf.basecall import score_batches
        feed = lambda n: ((i, host_batch) for i in range(n))
        for _ in score_batches(model, feed(2), scale=qs["scale"], offset=qs["bias"]):
            pass
        barrier()
        t0 = time.perf_counter()
        n_out = 0
        for _, out in score_batches(model, feed(args.steps), scale=qs["scale"], offset=qs["bias"]):
            n_out += int(out["moves"].shape[0])
        torch.cuda.synchronize()
        e2e_ms = (time.perf_counter() - t0) * 1e3
        assert n_out == N * args.steps
        barrier()
        for _ in range(2):
            compute_scores(model, host_batch, scale=qs["scale"], offset=qs["bias"])
        t0 = time.perf_counter()
        for _ in range(args.steps):
            out = compute_scores(model, host_batch, scale=qs["scale"], offset=qs["bias"])
        torch.cuda.synchronize()
        single_ms = (time.perf_counter() - t0) * 1e3 / args.steps
        barrier()
        log(f"e2e: {e2e_ms / args.steps:.2f} ms/step pipelined, {single_ms:.2f} ms/step one synchronous call per batch")

    if world > 1:
        t = torch.tensor([elapsed_ms, e2e_ms], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms, e2e_ms = t.tolist()

    if rank == 0:
        stage_ms = {}
        for name, a, b in events:
            stage_ms.setdefault(name, []).append(a.elapsed_time(b))
        # NOTE: the engine runs one CUDA stream per 32-chunk tile, so launches of different tiles overlap; these are
        # sums of per-launch durations (start/end events on the launch's own stream), not shares of the wall time.
        per_step = {k: sum(v) / args.steps for k, v in stage_ms.items()}
        launches = {k: len(v) // args.steps for k, v in stage_ms.items()}
        step_ms = elapsed_ms / args.steps
        H = spec["hidden"]
        tile_mode = bool(plan.tile) and os.environ.get("B200_LSTM_TILE", "1") != "0"
        tile_chunks = plan.tile if tile_mode else plan.TILE
        cluster = plan.tile_cs if tile_mode else 8
        n_tiles = -(-N // tile_chunks)
        flops_step = {  # algorithmic FLOPs per step, all launches of the kernel
            "lstm_rec": spec["n_lstm"] * 2.0 * N * T * 4 * H * H,
            "lstm_in_gemm": spec["n_lstm"] * 2.0 * N * T * 4 * H * H,
            "conv_gemm": 2.0 * N * T * H * plan.k3 * plan.c2,
            "crf_gemm": 2.0 * N * T * plan.n_scores * H,
        }
        flops = {k: v / max(launches.get(k, 1), 1) for k, v in flops_step.items()}
        sms = torch.cuda.get_device_properties(device).multi_processor_count
        dominant = "lstm_rec"
        dur = per_step[dominant] / launches[dominant] * 1e-3
        ach = flops[dominant] / dur / 1e12
        # one launch = one cluster (6 or 8 SMs) working on one tile-layer, or all tiles of a layer: compare with that share of the chip
        launch_sms = cluster if launches[dominant] > spec["n_lstm"] else min(cluster * n_tiles, sms)
        peak_share = peaks["tflops"] * launch_sms / sms
        chip_ach = flops_step[dominant] / (step_ms * 1e-3) / 1e12
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as fh:
                traffic = json.load(fh)["lstm_rec"]["bytes"] if N == BATCH else None
        except Exception:
            pass
        roof = {"kernel": ("lstm_rec_tc6_kernel (persistent tcgen05 LSTM layer, one 6-CTA cluster per 48-chunk tile)" if tile_mode else
                           "lstm_rec_tc_kernel (persistent tcgen05 LSTM layer, one 8-CTA cluster per 32-chunk tile)"),
                "bound": "tensor", "achieved": ach, "peak": peak_share, "unit": "TFLOP/s", "frac": ach / peak_share,
                "traffic": traffic, "traffic_algorithmic": 2.0 * T * (N / launches[dominant] * spec["n_lstm"]) * 5 * H,
                "peak_source": f"{peaks['source']} sustained bf16 GEMM {peaks['tflops']} TFLOP/s x {launch_sms}/{sms} SMs "
                               "(the share of the chip one launch occupies)",
                "launch_ms": dur * 1e3, "launches_per_step": launches[dominant],
                "flops_per_launch": flops[dominant],
                "chip_level": {"achieved": chip_ach, "peak": peaks["tflops"], "frac": chip_ach / peaks["tflops"],
                               "note": "all lstm_rec FLOPs of a step / whole step time (other kernels overlap)"}}
        total_flops = sum(flops_step.values())
        line = {
            "metric": METRIC, "value": world * N * L * args.steps / (elapsed_ms * 1e-3), "unit": "samples/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": step_ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16", "data": "synthetic",
            "config": {"workload": f"{MODEL}-shaped LSTM-CRF (H={H}, {spec['n_lstm']} LSTM, {plan.n_scores} scores/frame), "
                                   f"batch {N}/GPU, {CHUNK}->{L}-sample chunks ({T} frames), forward+decode",
                       "weights": "seeded synthetic (bonito_b200/synth.py)", "l2": "per-step tensors (0.16-2.6 GB) exceed the 126 MB L2",
                       "parallelism": f"chunk-sharded replicas x{world}"},
            "e2e": {"value": world * N * L * args.steps / (e2e_ms * 1e-3), "unit": "samples/s",
                    "h2d_bytes_per_step": N * L * 2, "d2h_bytes_per_step": 3 * N * T, "ms_per_step": e2e_ms / args.steps,
                    "api": "bonito_b200.crf.basecall.score_batches(model, float32 host batches): the loop basecall() runs", "single_call_ms_per_step": single_ms},
            "gpu_launches": sum(len(v) for v in stage_ms.values()),
            "roofline": roof,
            "stage_launch_ms_summed_per_step": {k: round(v, 4) for k, v in per_step.items()},
            "launches_per_step": launches,
            "host_enqueue_ms_per_step": round(host_enqueue_ms, 3),
            "model_tflops_per_s": total_flops / (step_ms * 1e-3) / 1e12,
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(spec, weights, chunksize)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

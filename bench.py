#!/usr/bin/env python
"""Headline benchmark: audio-seconds/sec (RTF^-1) of Fr-En simultaneous S2ST at chunk = 320 ms (BASELINE.json).

A "step" is one pass of the hot path over one synthetic 10 s / 16 kHz utterance, streamed in 320 ms chunks
(32 policy() calls: fbank -> chunk-Conformer encoder -> 2 CTC heads -> policy gate -> MT decoder -> T2U ->
NAR unit decoder -> CodeHiFiGAN vocoder), batch 1, exactly BASELINE.json configs[1].

  value : audio-seconds / second with the utterance already resident in HBM and results left on the device
  e2e   : the same through the reference-facing agent API (StreamSpeechS2STAgent.pushpop with python-list
          SpeechSegments): host->device copy of every new chunk and device->host read of every emitted waveform
  N > 1 : one process per GPU (torchrun), utterances sharded over ranks, no data-path collective; the synthetic
          checkpoint is NCCL-broadcast from rank 0 once before timing (weak scaling: every rank streams K utterances)

  --impl reference : the CPU restatement of the reference agent (oracle/, kind "port": the reference itself cannot be
          installed here) with all host threads, same utterance, same metric.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

UTT_SECONDS = 10.0
CHUNK_MS = 320
SAMPLE_RATE = 16000
METRIC = "audio-seconds/sec (RTF^-1) Fr-En simul S2ST @chunk=320ms"
WORKLOAD = "Fr-En simultaneous S2ST, chunk_size=320ms, batch=1, synthetic 10 s 16 kHz audio (BASELINE.json configs[1])"


def shard_utterances(n_utts: int, rank: int, world: int):
    """utterance ids of this rank: round-robin, no collective on the data path"""
    return [i for i in range(n_utts) if i % world == rank]


def agent_args(device_index=0, encoder_mode="cached"):
    return argparse.Namespace(model_path="synthetic", vocoder="synthetic", vocoder_cfg=None, data_bin=".", config_yaml=None,
                              multitask_config_yaml=None, sample_rate=SAMPLE_RATE, max_len=200, force_finish=False,
                              dur_prediction=True, lagging_k1=0, lagging_k2=0, segment_size=CHUNK_MS, stride_n=1, stride_n2=1,
                              unit_per_subword=15, source_segment_size=CHUNK_MS, vocoder_context="receptive-field",
                              device_index=device_index, encoder_mode=encoder_mode)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


def run_reference(args):
    """CPU arm: the reference agent's semantics (full-prefix recompute every chunk) in PyTorch fp32 on the host cores."""
    import torch

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle.agent_oracle import OracleS2STAgent
    from oracle.streamspeech_oracle import StreamSpeechOracle
    from streamspeech_b200 import synth
    from streamspeech_b200.config import ModelConfig

    torch.set_grad_enabled(False)
    cores = torch.get_num_threads()
    cfg = ModelConfig()
    o = StreamSpeechOracle(cfg, synth.make_model_state_dict(cfg, 0), synth.make_vocoder_state_dict(cfg.vocoder, 1), synth.make_gcmvn(cfg))
    wav = synth.make_audio(UTT_SECONDS, seed=1234)
    n = SAMPLE_RATE * CHUNK_MS // 1000

    def one():
        ag = OracleS2STAgent(o, CHUNK_MS)
        out = 0
        for i in range(0, len(wav), n):
            ag.push(wav[i:i + n].tolist(), finished=i + n >= len(wav))
            a = ag.policy()
            out += len(a.wav) if a.wav else 0
        return out

    for _ in range(args.warmup):
        one()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        one()
    dt = time.perf_counter() - t0
    value = args.steps * UTT_SECONDS / dt
    sample = f"{args.steps} x one {UTT_SECONDS:.0f} s utterance streamed in {CHUNK_MS} ms chunks (full workload per step)"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "audio-s/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1000 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": {"workload": WORKLOAD},
            "cpu_baseline": {"value": value, "unit": "audio-s/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def skinny_roofline(engine, peaks):
    """Dominant kernel of the streaming path = skinny_gemm_kernel (every encoder / MT projection at M <= 16 rows), a
    weight-streaming kernel -> HBM roofline.  Measured live with CUDA events on the launching stream at the encoder FFN
    shape (M = 16 active rows, K = 256, N = 2048); 96 distinct weight matrices (201 MB > 126 MB L2) are cycled so every
    launch streams its 2 MB of weights from HBM like the real step does.  algorithmic bytes = N*K*4 + M*K*4 + M*N*4."""
    import torch

    M, K, N, NW = 16, 256, 2048, 96
    x = torch.randn(M, K, device=engine.device)
    ws = [torch.randn(N, K, device=engine.device) * K ** -0.5 for _ in range(NW)]
    b = torch.zeros(N, device=engine.device)
    out = torch.empty(M, N, device=engine.device)

    def launch_all():
        for i in range(NW):
            engine.lib.ss_op_linear(engine._h, engine._stream(), x.data_ptr(), M, K, ws[i].data_ptr(), b.data_ptr(), N, 2, out.data_ptr())

    launch_all()
    torch.cuda.synchronize()
    # the 96 launches are replayed from a CUDA graph so that the host enqueue rate (python/ctypes, ~8 us per call) does not
    # hide the kernel duration; events bracket the replay on the launching stream
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        launch_all()
    graph.replay()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    graph.replay()
    e.record()
    torch.cuda.synchronize()
    us = s.elapsed_time(e) / NW * 1e3
    nbytes = N * K * 4 + M * K * 4 + M * N * 4
    ach = nbytes / (us * 1e-6) / 1e9
    peak = peaks.get("hbm_gbs", 6650.0)
    traffic = None
    tp = os.path.join(ROOT, "profiles", "r1_dominant_kernel_traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get("dram_bytes_per_launch")
    return {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
            "kernel": "skinny_gemm_kernel<16,2,1> (fp32 weight-streaming GEMM, fused SiLU)", "algorithmic_bytes_per_launch": nbytes,
            "peak_source": "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6.65 TB/s",
            "shape": {"M": M, "N": N, "K": K}, "us_per_launch": us,
            "note": "batch-1 streaming step: per-kernel latency (launch + one DRAM round trip + reduction) dominates, not bandwidth"}


def gemm_rooflines(engine, peaks):
    """Secondary: the two large-M GEMM kernels at the vocoder's heaviest conv shape (M = 5*500, N = 256, K = 11*256)."""
    import torch

    M, C, k = 5 * 500, 256, 11
    x = torch.randn(M, C * k, device=engine.device)
    w = torch.randn(C, C * k, device=engine.device) / (C * k) ** 0.5
    b = torch.zeros(C, device=engine.device)
    out = {}
    for name, fn in (("gemm_kernel<128,64> fp32 CUDA cores", lambda: engine.op_linear(x, w, b)),
                     ("umma_gemm_kernel<128,2> tcgen05 bf16x3 (opt-in)", lambda: engine.op_linear_umma(x, w, b, 0, 2))):
        for _ in range(3):
            fn()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        s.record()
        for _ in range(20):
            fn()
        e.record()
        torch.cuda.synchronize()
        ms = s.elapsed_time(e) / 20
        ach = 2.0 * M * C * C * k / (ms * 1e-3) / 1e12
        peak = peaks.get("bf16_tflops", 1590.0)
        out[name] = {"bound": "tensor", "achieved_fp32_equivalent": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                     "us_per_launch": ms * 1e3, "shape": {"M": M, "N": C, "K": C * k}}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", type=str, default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--encoder-mode", type=str, default="cached", choices=["cached", "recompute"])
    ap.add_argument("--ncu-window", action="store_true",
                    help="after the timed runs, stream one more resident utterance between cudaProfilerStart/Stop (for `ncu --profile-from-start off`)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch

    from streamspeech_b200 import synth
    from streamspeech_b200.agent import StreamSpeechS2STAgent
    from streamspeech_b200.simuleval_compat import SpeechSegment

    torch.set_grad_enabled(False)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    agent = StreamSpeechS2STAgent(agent_args(local, args.encoder_mode))
    eng = agent.engine
    if world > 1:
        # the one collective of the path: initial weight broadcast over NVLink.  Every rank built the
        # same seeded checkpoint; broadcasting rank 0's packed copy is what a real deployment does and costs ~0.3 s once.
        from streamspeech_b200.config import ModelConfig

        sd = synth.make_model_state_dict(ModelConfig(), 0)
        flat = torch.cat([v.flatten() for v in sd.values() if v.is_floating_point()]).cuda()
        dist.broadcast(flat, src=0)
        torch.cuda.synchronize()
        del flat

    n = SAMPLE_RATE * CHUNK_MS // 1000
    utts = [synth.make_audio(UTT_SECONDS, seed=1234 + 7 * rank + i) for i in range(2)]  # rank-specific utterances
    utts_dev = [u.cuda() for u in utts]
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")  # > 126 MB L2

    def stream_resident(u):
        agent.reset()
        out = 0
        for i in range(0, u.numel(), n):
            end = min(i + n, u.numel())
            w, wav = agent.step_resident(u, end, end >= u.numel())
            if w and wav is not None:
                out += wav.numel()
        return out

    def stream_e2e(u_host):
        agent.reset()
        out = 0
        for i in range(0, len(u_host), n):
            seg = agent.pushpop(SpeechSegment(content=u_host[i:i + n], sample_rate=SAMPLE_RATE, finished=i + n >= len(u_host)))
            if not seg.is_empty:
                out += len(seg.content)
        return out

    def timed(fn, inputs, steps, warmup):
        for i in range(warmup):
            fn(inputs[i % len(inputs)])
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        total_ms, out = 0.0, 0
        l0 = eng.launch_count()
        for i in range(steps):
            flush.fill_(float(i))  # L2 flush between timed iterations (outside the events)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            out += fn(inputs[i % len(inputs)])
            e.record()
            torch.cuda.synchronize()
            total_ms += s.elapsed_time(e)
        launches = eng.launch_count() - l0
        t = torch.tensor([total_ms], device="cuda")
        if dist is not None:
            dist.barrier()
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), out, launches

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_res, out_res, launches = timed(stream_resident, utts_dev, args.steps, args.warmup)
    utts_host = [u.tolist() for u in utts]
    ms_e2e, out_e2e, _ = timed(stream_e2e, utts_host, args.steps, max(1, args.warmup // 2))
    clocks = sampler.stop() if rank == 0 else None
    if args.ncu_window and rank == 0:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        stream_resident(utts_dev[0])
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    peaks = {}
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peaks = json.load(open(pk))
    audio_s = args.steps * UTT_SECONDS * world
    value = audio_s / (ms_res * 1e-3)
    e2e = audio_s / (ms_e2e * 1e-3)
    line = {"metric": METRIC, "value": value, "unit": "audio-s/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_res / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic 16 kHz audio + seeded random-init weights of the StreamSpeech architecture (calibrated token rates, see DESIGN.md)",
            "config": {"workload": WORKLOAD, "utterance_s": UTT_SECONDS, "chunk_ms": CHUNK_MS, "policy_calls_per_step": int(UTT_SECONDS * 1000 // CHUNK_MS) + 1,
                       "l2": "256 MiB flush between timed steps", "encoder_mode": args.encoder_mode, "parallelism": f"utterance-sharded x{world}",
                       "output_audio_s_per_step": out_res / args.steps / SAMPLE_RATE},
            "e2e": {"value": e2e, "unit": "audio-s/s", "h2d_bytes_per_step": int(UTT_SECONDS * SAMPLE_RATE * 4),
                    "d2h_bytes_per_step": int(out_e2e / args.steps * 4), "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches), "clocks": clocks}
    line["roofline"] = skinny_roofline(eng, peaks)
    line["roofline_large_gemm"] = gemm_rooflines(eng, peaks)
    if world == 1 and not args.no_cpu_baseline:
        # bounded CPU sample: the first 4 s of the same utterance through the oracle agent (reference semantics)
        from oracle.agent_oracle import OracleS2STAgent
        from oracle.streamspeech_oracle import StreamSpeechOracle
        from streamspeech_b200.config import ModelConfig

        cfg = ModelConfig()
        o = StreamSpeechOracle(cfg, synth.make_model_state_dict(cfg, 0), synth.make_vocoder_state_dict(cfg.vocoder, 1), synth.make_gcmvn(cfg))
        ag = OracleS2STAgent(o, CHUNK_MS)
        secs = 4.0
        w = utts[0][: int(secs * SAMPLE_RATE)]
        t0 = time.perf_counter()
        for i in range(0, len(w), n):
            ag.push(w[i:i + n].tolist(), finished=i + n >= len(w))
            ag.policy()
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": secs / dt, "unit": "audio-s/s", "cores": torch.get_num_threads(), "kind": "port",
                                "sample": f"first {secs:.0f} s of the same utterance, {CHUNK_MS} ms chunks, oracle agent (reference semantics: full-prefix recompute)"}
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

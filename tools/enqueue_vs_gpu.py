#!/usr/bin/env python
"""Is the streaming step CPU-launch-bound or GPU-bound?  Times the enqueue (host) and the execution (CUDA events)
of ss_encoder_stream_step and of an MT decode burst separately."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from streamspeech_b200 import synth
from streamspeech_b200.agent import StreamSpeechS2STAgent

torch.set_grad_enabled(False)
agent = StreamSpeechS2STAgent(bench.agent_args(0, "cached"))
eng = agent.engine
u = synth.make_audio(10.0, seed=1234).cuda()
feats = eng.fbank(u)
buf = torch.zeros(1024, 256, device="cuda")
for rep in range(2):
    eng.encoder_stream_reset()
    host, gpu, launches = [], [], []
    for k in range(1, 32):
        F = 32 * k - 2
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = eng.launch_count()
        t0 = time.perf_counter()
        s.record()
        eng.encoder_stream_step(feats[:F], buf)
        e.record()
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        host.append((t1 - t0) * 1e3); gpu.append(s.elapsed_time(e)); launches.append(eng.launch_count() - l0)
print("encoder_stream_step per call: host enqueue ms", sum(host[5:]) / len(host[5:]), "gpu ms", sum(gpu[5:]) / len(gpu[5:]), "launches", launches[10])
enc = buf[:250].contiguous()
for rep in range(2):
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = eng.launch_count()
    t0 = time.perf_counter(); s.record()
    toks, f = eng.mt_greedy(enc, None, 40)
    e.record(); t1 = time.perf_counter(); torch.cuda.synchronize()
print("mt_greedy 40 tokens: wall ms", (t1 - t0) * 1e3, "gpu ms", s.elapsed_time(e), "launches", eng.launch_count() - l0, "ntok", len(toks))
# pure launch-rate test: 2000 tiny layer norms
x = torch.randn(16, 256, device="cuda"); g = torch.ones(256, device="cuda"); b = torch.zeros(256, device="cuda")
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(2000): eng.op_layer_norm(x, g, b)
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print("2000 LN launches via ctypes: enqueue us/launch", (t1 - t0) / 2000 * 1e6, "total us/launch", (t2 - t0) / 2000 * 1e6)
# T2U + unit decoder and vocoder: host enqueue time vs GPU time of one call
toks, f = eng.mt_greedy(enc[:160].contiguous(), None, 30)
f = f.contiguous()
codes = torch.randint(0, 1000, (40,), device="cuda")
for name, fn in (("t2u_unit_decode S=%d" % f.shape[0], lambda: eng.t2u_unit_decode(f)),
                 ("vocoder 30 new frames", None)):
    if fn is None:
        dur, cum = eng.vocoder_durations(codes)
        total = int(cum[-1].item())
        fn = lambda: eng.vocoder_generate(total, total - 30, 30, -1)
    for _ in range(3):
        fn()
    hs, gs, ls = [], [], []
    for _ in range(10):
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = eng.launch_count(); t0 = time.perf_counter(); s.record()
        fn()
        e.record(); t1 = time.perf_counter(); torch.cuda.synchronize()
        hs.append((t1 - t0) * 1e3); gs.append(s.elapsed_time(e)); ls.append(eng.launch_count() - l0)
    hs.sort(); gs.sort()
    print(name, ": host enqueue ms", round(hs[5], 3), "gpu ms", round(gs[5], 3), "launches", ls[0])

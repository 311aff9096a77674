#!/usr/bin/env python
"""One steady-state invocation of each stage inside a cudaProfilerStart/Stop window, for
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/stage_launches.csv python tools/ncu_targets.py
"""
import os, sys
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
eng.encoder_stream_reset()
for k in range(1, 20):
    eng.encoder_stream_step(feats[: 32 * k - 2], buf)
enc = buf[:160].contiguous()
toks, f = eng.mt_greedy(enc, None, 30)
r = eng.t2u_unit_decode(f.contiguous())
codes = torch.randint(0, 1000, (40,), device="cuda")
dur, cum = eng.vocoder_durations(codes)
total = int(cum[-1].item())
eng.vocoder_generate(total, total - 30, 30, -1)
torch.cuda.synchronize()
torch.cuda.profiler.start()
eng.encoder_stream_step(feats[: 32 * 20 - 2], buf)           # steady-state encoder step: 16 active rows, T = 160
eng.ctc_greedy(0, buf[:160])
toks2, f2 = eng.mt_greedy(enc, toks[:27], 3)                   # prefill of 28 tokens + 3 decode steps
r = eng.t2u_unit_decode(f.contiguous())                        # S = 31 -> 775 unit positions
dur, cum = eng.vocoder_durations(codes)
eng.vocoder_generate(total, total - 30, 30, -1)                # 30 new frames + 24 context
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done", len(toks), total)

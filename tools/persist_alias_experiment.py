#!/usr/bin/env python
"""Timing experiment: persistent encoder step with every layer aliased to layer 0's weights (11 MB instead of 130 MB of
weights per step; results are numerically meaningless).  Separates memory-system cost (DRAM / TLB misses on weights that
are touched once per step) from the intrinsic latency of the phase chain."""
import json, os, sys
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
res = {}
for alias, bar in ((0, 0), (0, 1), (1, 0), (1, 1), (0, 0), (0, 1)):
    eng.set_option("persistent_alias", alias)
    eng.set_option("persistent_barrier", bar)
    gpu = []
    for rep in range(2):
        eng.encoder_stream_reset()
        for k in range(1, 32):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            s.record()
            eng.encoder_stream_step(feats[:32 * k - 2], buf)
            e.record()
            torch.cuda.synchronize()
            if rep == 1 and k > 4:
                gpu.append(s.elapsed_time(e))
    print("alias", alias, "own_barrier", bar, "gpu ms/step", sum(gpu) / len(gpu))

#!/usr/bin/env python
"""Host-side profile of the end-to-end path (StreamSpeechS2STAgent.pushpop with python-list segments): cProfile over 3 utterances."""
import cProfile, os, pstats, sys, io
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from streamspeech_b200 import synth
from streamspeech_b200.agent import StreamSpeechS2STAgent
from streamspeech_b200.simuleval_compat import SpeechSegment

torch.set_grad_enabled(False)
agent = StreamSpeechS2STAgent(bench.agent_args(0, "cached"))
u = synth.make_audio(10.0, seed=1234).tolist()
n = 5120


def run():
    agent.reset()
    out = 0
    for i in range(0, len(u), n):
        seg = agent.pushpop(SpeechSegment(content=u[i:i + n], sample_rate=16000, finished=i + n >= len(u)))
        if not seg.is_empty:
            out += len(seg.content)
    return out


for _ in range(2):
    run()
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(3):
    run()
torch.cuda.synchronize()
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(28)
print(s.getvalue()[:6000])

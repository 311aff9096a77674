#!/usr/bin/env python
"""Phase stamps of the single-token MT kernel (version 2): ns between the stamps CTA 0 records in layer 1 of the second step of a
launch (option persistent_profile).  Stamp ids: 0 layer start | 1 QKV done | 2 after barrier | 3 self-attn + Wo partials | 4 | 5 LN + Wcq |
6 | 7 cross-attn + Wco partials | 8 | 9 LN + FC1 | 10 | 11 FC2 | 12 (even ids: after the barrier that follows the odd one)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from streamspeech_b200.agent import StreamSpeechS2STAgent

torch.set_grad_enabled(False)
agent = StreamSpeechS2STAgent(bench.agent_args(0, "cached"))
eng = agent.engine
enc = torch.randn(1024, 256, device="cuda") * 0.5
toks, _ = eng.mt_greedy(enc[:160], None, 40)
toks = (toks + [17] * 40)[:40]
eng.set_option("persistent_profile", 1)
out = {}
for T in (160, 250):
    for _ in range(3):
        eng.encoder_stream_reset()
        eng.mt_greedy(enc[:T], toks[:20], 6, stable_rows=T - 16)
    torch.cuda.synchronize()
    st = eng.persistent_phase_stamps(64)
    n = int(st[0])
    pairs = [(int(st[1 + 2 * i]), int(st[2 + 2 * i])) for i in range(min(n, 30))]
    out[f"T{T}"] = [[pairs[i][0], pairs[i][1] - pairs[0][1], pairs[i][1] - pairs[i - 1][1] if i else 0] for i in range(len(pairs))]
eng.set_option("persistent_profile", 0)
print(json.dumps(out))
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "mt_phases.json"), "w"))

#!/usr/bin/env python
"""Per-phase time of the persistent encoder kernel (kernels_persist.cu), from the %globaltimer stamps it records when
option persistent_profile is set.  Prints the mean duration of each of the 11 phases of a layer (work + barrier) over
all layers and steps, the pure barrier cost, and the per-step GPU time with / without the persistent kernel."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from streamspeech_b200 import synth
from streamspeech_b200.agent import StreamSpeechS2STAgent

PHASES = ["ffn1_w1", "ffn1_w2", "qkv", "attention", "attn_out", "pw1_glu", "depthwise", "pw2", "ffn2_w1", "ffn2_w2", "final_ln"]

torch.set_grad_enabled(False)
agent = StreamSpeechS2STAgent(bench.agent_args(0, "cached"))
eng = agent.engine
u = synth.make_audio(10.0, seed=1234).cuda()
feats = eng.fbank(u)
buf = torch.zeros(1024, 256, device="cuda")
L = agent.cfg.enc_layers
out = {}
for mode in (1, 0):
    eng.set_option("persistent_encoder", mode)
    eng.set_option("persistent_profile", mode)
    acc = [0.0] * len(PHASES)
    barrier, n, gpu = 0.0, 0, []
    for rep in range(2):
        eng.encoder_stream_reset()
        for k in range(1, 32):
            F = 32 * k - 2
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            s.record()
            eng.encoder_stream_step(feats[:F], buf)
            e.record()
            torch.cuda.synchronize()
            if rep == 1 and k > 4:
                gpu.append(s.elapsed_time(e))
                if mode:
                    ts = eng.persistent_phase_stamps(3 + L * len(PHASES))
                    barrier += (ts[2] - ts[0]) / 2
                    for li in range(L):
                        for p in range(len(PHASES)):
                            i = 3 + li * len(PHASES) + p
                            acc[p] += ts[i] - ts[i - 1]
                    n += 1
    key = "persistent" if mode else "per_kernel"
    out[key] = {"gpu_ms_per_step": sum(gpu) / len(gpu)}
    if mode:
        out[key]["barrier_us"] = barrier / n / 1e3
        out[key]["phase_us"] = {PHASES[p]: acc[p] / (n * L) / 1e3 for p in range(len(PHASES))}
        out[key]["layer_us"] = sum(acc) / (n * L) / 1e3
print(json.dumps(out, indent=1))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "persist_phases.json"), "w"), indent=1)
# fine clock64 stamps of CTA 0 inside layer 1 of the last persistent step (see fine_stamp in kernels_persist.cu)
eng.set_option("persistent_encoder", 1)
eng.set_option("persistent_profile", 1)
eng.encoder_stream_reset()
for k in range(1, 20):
    eng.encoder_stream_step(feats[:32 * k - 2], buf)
torch.cuda.synchronize()
ts = eng.persistent_phase_stamps(256 + 60)[256:]
d = [ts[i] - ts[i - 1] for i in range(1, 50)]
print("fine cycle deltas (CTA 0, layer 1):", d)

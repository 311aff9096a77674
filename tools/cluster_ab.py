"""A / B of the two encoder-step kernels on the bench workload (10 s utterance, 320 ms steps): average launch time from CUDA
events on the launching stream (option persistent_time).  Usage: python tools/cluster_ab.py [seconds] [chunk]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from streamspeech_b200 import synth  # noqa: E402
from streamspeech_b200.config import ModelConfig  # noqa: E402
from streamspeech_b200.engine import Engine  # noqa: E402


def main():
    secs = float(sys.argv[1]) if len(sys.argv) > 1 else 10.0
    chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    cfg = ModelConfig()
    e = Engine(cfg, synth.make_model_state_dict(cfg, 0), synth.make_vocoder_state_dict(cfg.vocoder, 1), synth.make_gcmvn(cfg))
    e.set_chunk(chunk, min(chunk, 16) if chunk >= 16 else 8)
    feats = e.fbank(synth.make_audio(secs, seed=0).cuda())
    e.set_option("persistent_time", 1)
    out = {}
    finals = {}
    for rep in range(2):
        for v in (0, 1):
            e.set_option("persistent_encoder_cluster", v)
            buf = torch.zeros(2048, cfg.enc_dim, device="cuda")
            e.encoder_stream_reset()
            e.persistent_time()
            n0 = e.cluster_steps()
            T = 0
            for F in list(range(4 * chunk, feats.shape[0], 4 * chunk)) + [feats.shape[0]]:
                T, _ = e.encoder_stream_step(feats[:F].contiguous(), buf)
            ms, n, by = e.persistent_time()
            finals[v] = buf[:T].clone()
            out[f"rep{rep}_cluster{v}"] = {"us_per_launch": 1e3 * ms / max(n, 1), "launches": n, "cluster_steps": e.cluster_steps() - n0,
                                            "GBps": by / max(ms, 1e-9) / 1e6}
    # phase stamps of the cluster kernel (CTA 0, layer 1) on one more pass
    e.set_option("persistent_profile", 1)
    e.set_option("persistent_encoder_cluster", 1)
    e.encoder_stream_reset()
    for F in list(range(4 * chunk, feats.shape[0], 4 * chunk)) + [feats.shape[0]]:
        T, _ = e.encoder_stream_step(feats[:F].contiguous(), buf)
    torch.cuda.synchronize()
    st = e.persistent_phase_stamps(240)
    n = int(st[0])
    pairs = [(int(st[1 + 2 * i]), int(st[2 + 2 * i])) for i in range(min(n, 60))]
    out["stamps_ns"] = [[pairs[i][0], pairs[i][1] - pairs[0][1], pairs[i][1] - pairs[i - 1][1] if i else 0] if pairs[i][0] < 200 else list(pairs[i]) for i in range(len(pairs))]
    e.set_option("persistent_profile", 0)
    e.set_option("persistent_encoder_cluster", 0)
    e.check_async_error()
    out["maxdiff"] = float((finals[0] - finals[1]).abs().max())
    print(json.dumps(out))
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/cluster_ab.json", "w"), indent=1)


if __name__ == "__main__":
    main()

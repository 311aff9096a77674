"""Derive the calibration constants of the synthetic checkpoint (build-time tool; uses the CPU oracle).

Random-init weights make the decoders degenerate (CTC heads all-blank or never blank, MT decoder never
EOS, unit decoder all-blank or never blank), which would make the streaming benchmark meaningless
(SURVEY.md §8d).  This script measures, on seeded calibration audio, the margins of the seeded model
and derives the few numbers that give token RATES in the range of a trained model:

  * gcmvn mean/std of the synthetic audio family (stand-in for configs/fr-en/gcmvn.npz);
  * CTC heads: blank-bias delta = quantile of the blank margin  -> ~12 % / 10 % non-blank frames;
  * MT decoder: gain of the positional EOS feature (see streamspeech_b200/synth.py) -> EOS at ~30 tokens;
  * unit decoder: <blank> embedding row = a * mu/|mu|^2 with `a` a margin quantile -> ~15 % blank.

Output: streamspeech_b200/data/synth_calibration.npz (a few KB, committed).  Architecture, shapes,
FLOPs and state-dict keys are untouched and both the oracle and the CUDA engine load the same
tensors, so this cannot influence parity.

Run:  python -m oracle.calibrate_synth
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from streamspeech_b200.config import ModelConfig  # noqa: E402
from streamspeech_b200 import synth  # noqa: E402
import oracle.streamspeech_oracle as so  # noqa: E402
from oracle.streamspeech_oracle import StreamSpeechOracle, online_features, ctc_collapse  # noqa: E402

TARGET_NONBLANK = {"source_unigram": 0.22, "ctc_target_unigram": 0.20}
TARGET_EOS_POSITION = 52
TARGET_UNIT_BLANK = 0.10


def main():
    torch.set_grad_enabled(False)
    cfg = ModelConfig()
    out = {}
    raw = torch.cat([online_features(synth.make_audio(10.0, seed=s), None) for s in (100, 101, 102, 103)])
    out["gcmvn_mean"] = raw.mean(0).numpy()
    out["gcmvn_std"] = raw.std(0).numpy()
    gcmvn = {"mean": out["gcmvn_mean"].astype(np.float32), "std": out["gcmvn_std"].astype(np.float32)}
    sd = synth.make_model_state_dict(cfg, seed=0, calibration=None)
    o = StreamSpeechOracle(cfg, sd, None, gcmvn, chunk_size=8)
    # ---- CTC blank bias
    eos_ = []
    for s in (100, 101, 102):
        f = online_features(synth.make_audio(10.0, seed=s), gcmvn)
        eos_.append(o.encoder(f.unsqueeze(0), torch.tensor([f.size(0)]))["encoder_out"][0])
    for name, tgt in TARGET_NONBLANK.items():
        m = []
        for eo in eos_:
            lg = o.ctc_logits(name, eo)[0].clone()
            blank = lg[:, 0].clone()
            lg[:, [0, cfg.pad, cfg.unk]] = -1e30
            m.append(lg.max(dim=1)[0] - blank)
        out[f"{name}_blank_bias_delta"] = float(torch.quantile(torch.cat(m), 1.0 - tgt))
        print(name, "blank bias delta", out[f"{name}_blank_bias_delta"])
    enc_views = [eo[:n] for eo in eos_ for n in (40, 90, 150, 250)]

    # ---- MT: greedy without EOS, record max-other / eos-feature ratio per position
    E = sd["target_unigram_decoder.embed_tokens.weight"]
    E[cfg.eos].zero_()
    curves, seqs = [], []
    for eo in enc_views:
        toks = [cfg.eos]
        lgmax, heos = [], []
        for p in range(64):
            h = o.mt_features(torch.tensor([toks]), eo)[0, -1]
            lg = h @ E.T
            lg[[cfg.pad, cfg.eos]] = -1e30
            lgmax.append(float(lg.max()))
            heos.append(float(h[synth.MT_EOS_DIM]))
            toks.append(int(lg.argmax()))
        curves.append((np.array(lgmax), np.array(heos)))
        seqs.append(toks)

    def median_fire(b):
        fires = []
        for lgmax, heos in curves:
            hit = np.nonzero((b * heos > lgmax) & (np.arange(64) >= 1))[0]
            fires.append(hit[0] if len(hit) else 64)
        return float(np.median(fires)), fires
    lo, hi = 0.0, 10.0
    for _ in range(40):  # EOS fires earlier as b grows
        mid = 0.5 * (lo + hi)
        if median_fire(mid)[0] > TARGET_EOS_POSITION:
            lo = mid
        else:
            hi = mid
    out["mt_eos_b"] = hi
    print("mt_eos_b", hi, "first-fire positions", median_fire(hi)[1])
    print("distinct-token ratio", np.mean([len(set(s)) / len(s) for s in seqs]))
    E[cfg.eos, synth.MT_EOS_DIM] = out["mt_eos_b"]
    o = StreamSpeechOracle(cfg, sd, None, gcmvn, chunk_size=8)
    lens, seqs = [], []
    for eo in enc_views:
        hyp = o.mt_greedy(eo, None, -1)
        lens.append(len(hyp) - 1)
        seqs.append([cfg.eos] + hyp[:-1])
    print("mt lengths with eos", lens)

    # ---- unit decoder blank row
    feats = []
    for eo, toks in zip(enc_views, seqs):
        x = o.mt_features(torch.tensor([toks]), eo).transpose(0, 1)
        t2u = o.t2u_encoder(x, None)
        cap = {}
        orig = so.F.linear

        def hook(x_, w, b=None):
            if w.shape == (cfg.unit_vocab, cfg.unit_dim):
                cap["h"] = x_.reshape(-1, x_.shape[-1]).clone()
            return orig(x_, w, b)
        so.F.linear = hook
        try:
            o.unit_decoder_logits(t2u, None)
        finally:
            so.F.linear = orig
        feats.append(cap["h"])
    U = torch.cat(feats)
    W = sd["decoder.embed_tokens.weight"]
    mu = U.mean(0)
    proj = U @ mu / mu.dot(mu)
    lg = U @ W.T
    lg[:, [cfg.unit_blank, cfg.pad, cfg.unk]] = -1e30
    a = float(torch.quantile(lg.max(dim=1)[0] / proj.clamp_min(1e-3), TARGET_UNIT_BLANK))
    W[cfg.unit_blank] = a * mu / mu.dot(mu)
    out["unit_blank_row"] = W[cfg.unit_blank].numpy()
    lg = U @ W.T
    lg[:, [cfg.pad, cfg.unk]] = -1e30
    am = lg.argmax(1).tolist()
    hyp, _ = ctc_collapse(am, cfg.unit_blank, cfg.pad)
    print(f"unit: n={U.shape[0]} a={a:.3f} blank frac={np.mean(np.array(am) == cfg.unit_blank):.2f} "
          f"-> {25 * len(hyp) / len(am):.2f} units per MT token")
    path = os.path.join(ROOT, "streamspeech_b200", "data", "synth_calibration.npz")
    np.savez(path, **{k: np.asarray(v, dtype=np.float32) for k, v in out.items()})
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()

"""Generate tests/golden/offline_batch.npz from the REAL reference modules (build container only, CPU, ~1 min).

Run:  python -m oracle.gen_golden_offline      (needs /root/reference)

The offline batched generator (researches/ctc_unity/sequence_generator_multi_decoder_ctc.py:163-331)
cannot be imported as a class here (fairseq's SequenceGenerator / search need omegaconf & co.), so the fixture chains the
reference's OWN modules in the order `_generate` calls them, on a padded batch of three utterances of different lengths:

  reference encoder (offline model, chunk_size None)  ->  reference CTCDecoder.generate from its own source
  (researches/ctc_unity/ctc_decoder.py)  ->  MT hypotheses [greedy, restated: oracle.mt_greedy, "parity unpinned" for the search
  loop as in DESIGN.md §2; every decoder forward inside it is the pinned restatement]  ->  reference TransformerDecoderBase
  (features_only) on the padded prev_output_tokens_mt  ->  reference UniTransformerEncoderNoEmb with the padding mask  ->
  reference CTCTransformerUnitDecoder on the batch  ->  reference offline CTCSequenceGenerator.generate from its own source
  (researches/ctc_unity/ctc_generator.py, masks pad / unk / eos, keeps tokens at padded positions).

and asserts that oracle/offline_oracle.py reproduces every stage before writing the fixture.
"""
from __future__ import annotations

import math
import os
import sys

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402
from oracle.gen_golden import (FakeDict, GOLD, REF, build_reference_decoders, build_reference_encoder, extract_class,  # noqa: E402
                               maxdiff)
from oracle.offline_oracle import offline_generate  # noqa: E402
from oracle.streamspeech_oracle import StreamSpeechOracle, online_features  # noqa: E402
from streamspeech_b200 import synth  # noqa: E402
from streamspeech_b200.config import ModelConfig  # noqa: E402


def offline_cfg() -> ModelConfig:
    cfg = ModelConfig()       # full 12-layer encoder: the synthetic checkpoint's token rates are calibrated for it
    cfg.uni_encoder = False   # offline model: bidirectional T2U encoder (N10)
    return cfg


MAX_LEN_B_MT = 60             # two of the three hypotheses hit the forced eos, one ends earlier -> padded MT rows


def main():
    assert ref_loader.available(), "reference tree missing"
    torch.manual_seed(0)
    torch.set_grad_enabled(False)
    ref_loader.load_full()
    import fairseq.utils as futils
    from typing import Dict, List, Optional

    cfg = offline_cfg()
    sd = synth.make_model_state_dict(cfg, seed=0)
    gcmvn = synth.make_gcmvn(cfg)
    wavs = [synth.make_audio(2.0, seed=31), synth.make_audio(1.5, seed=32), synth.make_audio(1.1, seed=33)]
    fl = [online_features(w, gcmvn) for w in wavs]
    lens = torch.tensor([f.size(0) for f in fl])
    feats = torch.zeros(len(fl), int(lens.max()), cfg.feat_dim)
    for b, f in enumerate(fl):
        feats[b, : f.size(0)] = f

    # ---- reference modules
    enc = build_reference_encoder(cfg, sd, None, None)
    mods = build_reference_decoders(cfg, sd)
    glb = {"torch": torch, "nn": nn, "math": math, "utils": futils, "List": List, "Dict": Dict, "Optional": Optional, "Tensor": torch.Tensor}
    RefCTC = extract_class(REF + "/researches/ctc_unity/ctc_decoder.py", "CTCDecoder", glb)
    RefUnitCTC = extract_class(REF + "/researches/ctc_unity/ctc_generator.py", "CTCSequenceGenerator", glb)

    class FakeModel(nn.Module):
        def __init__(self):
            super().__init__()
            self.source_unigram_decoder = mods["source_unigram"]
            self.ctc_target_unigram_decoder = mods["ctc_target_unigram"]
            self.decoder = mods["unit"]

        def max_decoder_positions(self):
            return 1200

        def get_normalized_probs(self, net_output, log_probs, sample=None):
            logits = net_output[0].float()
            return F.log_softmax(logits, dim=-1) if log_probs else F.softmax(logits, dim=-1)

    fake = FakeModel()
    orc = StreamSpeechOracle(cfg, sd, None, gcmvn, chunk_size=None, conv_chunk_size=None)
    mine = offline_generate(orc, feats, lens, max_len_b_mt=MAX_LEN_B_MT)
    report = {}

    # ---- encoder + CTC prints
    enc_out = enc._forward(feats, lens)
    eo = enc_out["encoder_out"][0]
    report["encoder"] = maxdiff(eo, mine["enc_out"])
    assert report["encoder"] < 2e-5, report
    gold = {"feats": feats.numpy(), "lengths": lens.numpy(), "enc_out": eo.numpy(), "out_lengths": mine["out_lengths"].numpy()}
    for name, key in (("source_unigram", "asr"), ("ctc_target_unigram", "st")):
        hyp = RefCTC(FakeDict(6000), [fake]).generate(enc_out, aux_task_name=name)
        for b in range(feats.size(0)):
            assert hyp[b][0]["tokens"].tolist() == mine[key][b]["tokens"], (name, b)
            assert hyp[b][0]["org_tokens"].tolist() == mine[key][b]["org_tokens"], (name, b)
            gold[f"{key}_tokens_{b}"] = hyp[b][0]["tokens"].numpy()
            gold[f"{key}_argmax_{b}"] = hyp[b][0]["org_tokens"].numpy()

    # ---- MT hypotheses (restated greedy search), then the reference decoder on the padded batch
    prev = mine["prev_output_tokens_mt"]
    report["mt_lengths"] = [len(h) for h in mine["mt_hyps"]]
    assert len(set(report["mt_lengths"])) > 1, "fixture should contain padded MT rows"
    ref_feats, _ = mods["mt"](prev, encoder_out=enc_out, features_only=True)
    report["mt_feats"] = maxdiff(ref_feats, mine["mt_feats"])
    assert report["mt_feats"] < 5e-5, report
    gold["prev_output_tokens_mt"] = prev.numpy()
    gold["mt_feats"] = ref_feats.numpy()
    for b, h in enumerate(mine["mt_hyps"]):
        gold[f"mt_hyp_{b}"] = np.array(h)

    # ---- T2U encoder + unit decoder + offline CTC generate, batched
    pad_mask = prev.eq(cfg.pad) if prev.eq(cfg.pad).any() else None
    t2u_ref = mods["t2u"](ref_feats.transpose(0, 1), pad_mask)
    report["t2u"] = maxdiff(t2u_ref["encoder_out"][0], mine["t2u_out"])
    assert report["t2u"] < 5e-5, report
    ul_ref, _ = mods["unit"](None, encoder_out=t2u_ref)
    report["unit_logits"] = maxdiff(ul_ref, mine["unit_logits"])
    assert report["unit_logits"] < 2e-4, report
    hyp = RefUnitCTC(FakeDict(cfg.unit_vocab, blank=cfg.unit_blank), [fake]).generate(t2u_ref)
    lprobs = F.log_softmax(ul_ref.float(), dim=-1)
    for col in (cfg.pad, cfg.unk, cfg.eos):
        lprobs[:, :, col] = -math.inf
    ref_argmax = lprobs.argmax(-1)
    for b in range(feats.size(0)):
        assert hyp[b][0]["tokens"].tolist() == mine["units"][b]["tokens"], b
        assert ref_argmax[b].tolist() == mine["units"][b]["org_tokens"], b
        gold[f"units_{b}"] = hyp[b][0]["tokens"].numpy()
        gold[f"unit_argmax_{b}"] = ref_argmax[b].numpy()
        gold[f"unit_logits_first_{b}"] = ul_ref[b, :4].numpy()
        report[f"n_units_{b}"] = int(hyp[b][0]["tokens"].numel())
    gold["max_len_b_mt"] = np.array(MAX_LEN_B_MT)
    np.savez_compressed(os.path.join(GOLD, "offline_batch.npz"), **gold)
    for k, v in report.items():
        print(f"{k:24s} {v}")
    print("wrote", os.path.join(GOLD, "offline_batch.npz"), os.path.getsize(os.path.join(GOLD, "offline_batch.npz")), "bytes")


if __name__ == "__main__":
    main()

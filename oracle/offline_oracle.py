"""TEST INFRASTRUCTURE (CPU oracle) -- not part of the product path.

CPU restatement of the reference's OFFLINE batched generator with beam_size_mt = 1
(`CTCMultiDecoderSequenceGenerator._generate`, researches/ctc_unity/sequence_generator_multi_decoder_ctc.py:163-331;
SURVEY.md §8 row O1), composed from the per-block restatements of oracle/streamspeech_oracle.py:

  batched padded encoder (offline model: chunk_size None, N10) -> ASR / ST CTC prints over ALL T rows of every sample
  (researches/ctc_unity/ctc_decoder.py:40-111, padded frames are not trimmed) -> greedy MT per sample with the
  encoder padding mask (generate_decoder, beam 1, max_len = max_len_b_mt) -> prev_output_tokens_mt [B, max_tgt_len]
  = [eos, hyp..., pad...] (:262-271) -> mt_decoder(features_only) on the padded batch (:287-291) -> T2U encoder with the
  padding mask (:300-305) -> CTC unit decoder on the batch, with the reference's quirks:
      N1  positional embedding indexed by the batch axis: sample b gets position b + 2
      N2  hypotheses keep the tokens emitted at padded T2U positions (researches/ctc_unity/ctc_generator.py:66-89)
      N3  pad, unk AND eos are masked before the arg-max (:56-58)
Pinned by oracle/gen_golden_offline.py against the reference's own modules (tests/golden/offline_batch.npz).
"""
from __future__ import annotations

from typing import Dict, List

import torch

from .streamspeech_oracle import StreamSpeechOracle


def offline_generate(o: StreamSpeechOracle, feats: torch.Tensor, lengths: torch.Tensor, max_len_b_mt: int = 200) -> Dict[str, object]:
    """feats [B, F, 80] (zero padded), lengths [B].  `o` must be an offline-model oracle (chunk_size None; cfg.uni_encoder as
    the checkpoint was trained).  Returns per-sample lists plus the batched intermediates the parity tests compare."""
    c = o.cfg
    enc = o.encoder(feats, lengths)
    eo = enc["encoder_out"][0]  # [T, B, C]
    T, B, _ = eo.shape
    pad_mask = enc["encoder_padding_mask"][0] if enc["encoder_padding_mask"] else None
    out_len = enc["out_lengths"]
    asr = o.ctc_greedy("source_unigram", eo)
    st = o.ctc_greedy("ctc_target_unigram", eo)
    # 1. MT decoder, beam 1, per sample (the batch only shares the padding mask)
    hyps: List[List[int]] = []
    for b in range(B):
        Tb = int(out_len[b])
        hyp = o.mt_greedy(eo[:Tb, b:b + 1], None, -1, max_len_b=max_len_b_mt)  # tokens + eos
        hyps.append(hyp)
    max_tgt_len = max(len(h) for h in hyps)
    prev = torch.full((B, max_tgt_len), c.pad, dtype=torch.long)
    for b, h in enumerate(hyps):
        toks = h[:-1] if h[-1] == c.eos else h
        prev[b, 0] = c.eos
        prev[b, 1:len(toks) + 1] = torch.tensor(toks, dtype=torch.long)
    x = o.mt_features(prev, eo, pad_mask)  # [B, L, 512]
    mt_pad = prev.eq(c.pad) if prev.eq(c.pad).any() else None
    # 2. T2U encoder, 3. CTC unit decoder + offline CTC generate
    t2u = o.t2u_encoder(x.transpose(0, 1), mt_pad)
    logits = o.unit_decoder_logits(t2u, mt_pad)  # [B, 25 L, V]
    units = o.unit_ctc_greedy(logits, mask_eos=True)
    return {
        "enc_out": eo, "out_lengths": out_len, "asr": asr, "st": st, "mt_hyps": hyps, "prev_output_tokens_mt": prev,
        "mt_feats": x, "t2u_out": t2u, "unit_logits": logits, "units": units,
    }

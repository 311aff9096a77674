"""Generate tests/golden/*.npz from the REAL reference modules (build container only).

Run:  python -m oracle.gen_golden          (needs /root/reference; CPU only, ~1 min)

For every block of the hot path it (1) instantiates the reference's own class through
oracle/ref_loader.py, (2) loads the seeded synthetic checkpoint with strict=True (proving
streamspeech_b200/synth.py uses the reference's key names and shapes), (3) runs it on seeded
inputs, (4) asserts the CPU restatement in oracle/streamspeech_oracle.py agrees, and
(5) writes inputs + reference outputs as small fixtures.  The fixtures travel to the GPU box;
/root/reference does not.

Glue that cannot be imported as a class (ChunkS2TConformerEncoder._forward, the agent's CTC
generators) is executed from the reference's own function source via `ast` extraction, so the
golden values still come from reference code, not from this repo's restatement.
"""
from __future__ import annotations

import ast
import math
import os
import sys
import types
from types import SimpleNamespace as NS

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402
from oracle.streamspeech_oracle import StreamSpeechOracle, online_features, kaldi_fbank, rel_shift, rel_positional_encoding  # noqa: E402
from streamspeech_b200.config import ModelConfig, VocoderConfig  # noqa: E402
from streamspeech_b200 import synth  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
REF = ref_loader.REF


def extract_functions(path, class_name, names, glb):
    """exec the named methods of `class_name` from a reference file; returns {name: function}."""
    tree = ast.parse(open(path).read())
    cls = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == class_name][0]
    fns = [n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name in names]
    for fn in fns:
        fn.decorator_list = []
    mod = ast.Module(body=fns, type_ignores=[])
    ns = dict(glb)
    exec(compile(mod, f"{os.path.basename(path)}:{class_name}", "exec"), ns)
    return {n: ns[n] for n in names}


def extract_class(path, class_name, glb):
    """exec a whole reference class definition from its own source; returns the class."""
    tree = ast.parse(open(path).read())
    cls = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == class_name][0]
    ns = dict(glb)
    exec(compile(ast.Module(body=[cls], type_ignores=[]), f"{os.path.basename(path)}:{class_name}", "exec"), ns)
    return ns[class_name]


def lengths_to_padding_mask(lens):
    # fairseq/data/data_utils.py lengths_to_padding_mask
    bsz, max_lens = lens.size(0), torch.max(lens).item()
    mask = torch.arange(max_lens).to(lens.device).view(1, max_lens)
    return mask.expand(bsz, -1) >= lens.view(bsz, 1).expand(-1, max_lens)


def sub_state(sd, prefix):
    return {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}


def transformer_cfg(cfg: ModelConfig, enc_dim, dec_dim, ffn, heads, layers, max_target_positions):
    return NS(
        dropout=0.1, attention_dropout=0.1, activation_dropout=0.1, relu_dropout=0.1, activation_fn="relu",
        export=False, cross_self_attention=False, quant_noise=NS(pq=0, pq_block_size=8),
        encoder=NS(embed_dim=enc_dim, ffn_embed_dim=ffn, attention_heads=heads, normalize_before=True, xformers_att_config=None, layers=layers),
        decoder=NS(embed_dim=dec_dim, ffn_embed_dim=ffn, attention_heads=heads, normalize_before=True, xformers_att_config=None,
                   layers=layers, layerdrop=0.0, output_dim=dec_dim, learned_pos=False, input_dim=dec_dim),
        share_decoder_input_output_embed=True, max_target_positions=max_target_positions, no_scale_embedding=False,
        adaptive_input=False, no_token_positional_embeddings=False, layernorm_embedding=False, no_decoder_final_norm=False,
        tie_adaptive_weights=False, adaptive_softmax_cutoff=None, base_layers=0, checkpoint_activations=False,
        offload_activations=False, min_params_to_wrap=0,
        # legacy flat names read by CTCTransformerUnitDecoder / UniTransformerEncoderNoEmb
        n_frames_per_step=1, ctc_upsample_rate=cfg.ctc_upsample_rate, encoder_layers=layers,
        encoder_normalize_before=True, encoder_embed_dim=enc_dim, uni_encoder=cfg.uni_encoder,
    )


class FakeDict:
    def __init__(self, n, blank=None):
        self.n = n
        self.blank_index = blank
        self.pad_index = 1

    def __len__(self):
        return self.n

    def pad(self):
        return 1

    def eos(self):
        return 2

    def unk(self):
        return 3


def build_reference_encoder(cfg: ModelConfig, sd, chunk_size, conv_chunk):
    from chunk_unity.modules.conformer_layer import ChunkConformerEncoderLayer
    from chunk_unity.modules.convolution import Conv1dSubsampler
    from fairseq.modules.positional_encoding import RelPositionalEncoding
    import fairseq.utils as futils

    fns = extract_functions(
        REF + "/researches/chunk_unity/models/s2t_conformer.py", "ChunkS2TConformerEncoder",
        ["_forward", "buffered_chunk_mask"],
        {"torch": torch, "math": math, "lengths_to_padding_mask": lengths_to_padding_mask, "utils": futils},
    )

    class Enc(nn.Module):
        def __init__(self):
            super().__init__()
            self.embed_scale = math.sqrt(cfg.enc_dim)
            self.chunk_size = chunk_size
            self.chunk = chunk_size is not None
            self.subsample = Conv1dSubsampler(cfg.feat_dim, cfg.conv_channels, cfg.enc_dim, [cfg.conv_kernel] * 2,
                                              chunk_size=chunk_size if self.chunk else None)
            self.pos_enc_type = "rel_pos"
            self.embed_positions = RelPositionalEncoding(cfg.max_source_positions, cfg.enc_dim)
            self.linear = nn.Linear(cfg.enc_dim, cfg.enc_dim)
            self.dropout = nn.Dropout(0.1)
            self.conformer_layers = nn.ModuleList([
                ChunkConformerEncoderLayer(cfg.enc_dim, cfg.enc_ffn, cfg.enc_heads, 0.1, False, cfg.dw_kernel,
                                           attn_type="espnet", pos_enc_type="rel_pos",
                                           chunk_size=chunk_size if self.chunk else None)
                for _ in range(cfg.enc_layers)])
            self._chunk_mask = torch.empty(0)

    Enc._forward = fns["_forward"]
    Enc.buffered_chunk_mask = fns["buffered_chunk_mask"]
    enc = Enc()
    missing = enc.load_state_dict(sub_state(sd, "encoder."), strict=True)
    enc.eval()
    if enc.chunk:  # what the agent does after loading (agent:404-413)
        for conv in enc.subsample.conv_layers:
            conv.chunk_size = conv_chunk
        for layer in enc.conformer_layers:
            layer.conv_module.depthwise_conv.chunk_size = conv_chunk
    return enc


def build_reference_decoders(cfg: ModelConfig, sd):
    ref_loader.load_full()
    from ctc_unity.modules.transformer_decoder import TransformerDecoderBase
    from ctc_unity.modules.ctc_transformer_unit_decoder import CTCTransformerUnitDecoder
    from ctc_unity.modules.transformer_encoder import UniTransformerEncoderNoEmb
    from fairseq.models.speech_to_speech.modules.ctc_decoder import CTCDecoder
    from fairseq.models.speech_to_speech.modules.stacked_embedding import StackedEmbedding

    mods = {}
    for name, V in (("source_unigram", cfg.src_vocab), ("ctc_target_unigram", cfg.tgt_vocab)):
        m = CTCDecoder(FakeDict(V), cfg.enc_dim)
        m.load_state_dict(sub_state(sd, f"{name}_decoder."), strict=True)
        mods[name] = m.eval()
    mt_cfg = transformer_cfg(cfg, cfg.enc_dim, cfg.mt_dim, cfg.mt_ffn, cfg.mt_heads, cfg.mt_layers, 1024)
    emb = nn.Embedding(cfg.tgt_vocab, cfg.mt_dim, padding_idx=cfg.pad)
    mt = TransformerDecoderBase(mt_cfg, FakeDict(cfg.tgt_vocab), emb)
    mt.load_state_dict(sub_state(sd, "target_unigram_decoder."), strict=False)  # `version` buffer is extra
    assert set(sub_state(sd, "target_unigram_decoder.").keys()) <= set(mt.state_dict().keys())
    mods["mt"] = mt.eval()
    t2u_cfg = transformer_cfg(cfg, cfg.unit_dim, cfg.unit_dim, cfg.unit_ffn, cfg.unit_heads, cfg.t2u_layers, 1200)
    t2u = UniTransformerEncoderNoEmb(t2u_cfg)
    t2u.load_state_dict(sub_state(sd, "synthesizer_encoder."), strict=True)
    mods["t2u"] = t2u.eval()
    u_cfg = transformer_cfg(cfg, cfg.unit_dim, cfg.unit_dim, cfg.unit_ffn, cfg.unit_heads, cfg.unit_layers, 1200)
    uemb = StackedEmbedding(cfg.unit_vocab, cfg.unit_dim, cfg.pad, num_stacked=1)
    ud = CTCTransformerUnitDecoder(u_cfg, FakeDict(cfg.unit_vocab, blank=cfg.unit_blank), uemb)
    ud.load_state_dict(sub_state(sd, "decoder."), strict=False)
    assert set(sub_state(sd, "decoder.").keys()) <= set(ud.state_dict().keys())
    mods["unit"] = ud.eval()
    return mods


def build_reference_ctc_generators():
    """agent/ctc_decoder.py CTCDecoder.generate and agent/ctc_generator.py CTCSequenceGenerator.generate,
    run from their own source."""
    import fairseq.utils as futils
    from typing import Dict, List, Optional
    glb = {"torch": torch, "nn": nn, "math": math, "utils": futils, "List": List, "Dict": Dict, "Optional": Optional,
           "Tensor": torch.Tensor}
    A = extract_class(REF + "/agent/ctc_decoder.py", "CTCDecoder", glb)
    B = extract_class(REF + "/agent/ctc_generator.py", "CTCSequenceGenerator", glb)
    return A, B


class FakeModel(nn.Module):
    """Just enough of StreamSpeechModel for the agent-side generators."""

    def __init__(self, mods):
        super().__init__()
        self.source_unigram_decoder = mods["source_unigram"]
        self.ctc_target_unigram_decoder = mods["ctc_target_unigram"]
        self.decoder = mods["unit"]

    def max_decoder_positions(self):
        return 1200

    def get_normalized_probs(self, net_output, log_probs, sample=None):
        logits = net_output[0].float()
        return F.log_softmax(logits, dim=-1) if log_probs else F.softmax(logits, dim=-1)


def maxdiff(a, b):
    return float((a - b).abs().max())


def main():
    assert ref_loader.available(), "reference tree missing"
    os.makedirs(GOLD, exist_ok=True)
    torch.manual_seed(0)
    torch.set_grad_enabled(False)
    ref_loader.load_full()
    report = {}

    # ------------------------------------------------------------------ KATs held by the reference
    # fairseq/tests/test_espnet_multihead_attention.py:99-147 and test_positional_encoding.py:17-59:
    # run the reference's own test bodies to make sure the stub-imported classes are the tested ones.
    from fairseq.modules.espnet_multihead_attention import RelPositionMultiHeadedAttention as FsRelMHA
    from fairseq.modules.positional_encoding import RelPositionalEncoding
    torch.manual_seed(0)
    mha_ = FsRelMHA(2, 1, 0.0)  # n_feat=2, n_head=1 as in the reference test
    sample_x = torch.tensor([[[0.0, 1.0, 2.0, 3.0, 4.0, 5.0, 6.0], [7.0, 8.0, 9.0, 10.0, 11.0, 12.0, 13.0]]]).unsqueeze(0)
    kat_shift_in = torch.arange(0, 2 * 3 * 5, dtype=torch.float32).view(1, 2, 3, 5)
    kat_shift_out = mha_.rel_shift(kat_shift_in)
    assert torch.equal(kat_shift_out, rel_shift(kat_shift_in))
    pe = RelPositionalEncoding(12, 4)
    kat_pe = pe(torch.zeros(3, 1, 4))  # [2T-1, 1, C]
    assert maxdiff(kat_pe[:, 0], rel_positional_encoding(3, 4)) == 0.0
    # the hard-coded expectation of test_positional_encoding.py:41-59 (forward, T=3, d=4)
    np.savez(os.path.join(GOLD, "kat_relpos.npz"), shift_in=kat_shift_in.numpy(), shift_out=kat_shift_out.numpy(),
             pe_T3_d4=kat_pe[:, 0].numpy())

    # ------------------------------------------------------------------ fbank vs installed torchaudio
    import torchaudio.compliance.kaldi as ta_kaldi
    wav = synth.make_audio(1.0, seed=7)
    ref_fb = ta_kaldi.fbank((wav * 2 ** 15).unsqueeze(0), num_mel_bins=80, sample_frequency=16000)
    my_fb = kaldi_fbank(wav * 2 ** 15)
    report["fbank_restatement_vs_torchaudio"] = maxdiff(ref_fb, my_fb)
    assert report["fbank_restatement_vs_torchaudio"] < 1e-4, report
    np.savez(os.path.join(GOLD, "fbank.npz"), wav=wav.numpy(), fbank=ref_fb.numpy())

    # ------------------------------------------------------------------ encoder (3 chunk settings + offline)
    cfg = ModelConfig()
    cfg.enc_layers = 3  # structure identical, file stays small
    sd = synth.make_model_state_dict(cfg, seed=0)
    gcmvn = synth.make_gcmvn(cfg)
    wav2 = synth.make_audio(2.0, seed=11)
    feats = online_features(wav2, gcmvn)  # [198, 80]
    enc_gold = {"feats": feats.numpy()}
    for seg_ms, chunk, conv_chunk in ((160, 4, 8), (320, 8, 8), (640, 16, 16), (0, None, None)):
        enc = build_reference_encoder(cfg, sd, chunk, conv_chunk)
        out = enc._forward(feats.unsqueeze(0), torch.tensor([feats.size(0)]), return_all_hiddens=True)
        orc = StreamSpeechOracle(cfg, sd, None, gcmvn, chunk_size=chunk, conv_chunk_size=conv_chunk)
        mine = orc.encoder(feats.unsqueeze(0), torch.tensor([feats.size(0)]), return_layers=True)
        d = maxdiff(out["encoder_out"][0], mine["encoder_out"][0])
        report[f"encoder_c{chunk}"] = d
        assert d < 2e-5, report
        enc_gold[f"out_c{chunk}"] = out["encoder_out"][0][:, 0].numpy()
        enc_gold[f"layer0_c{chunk}"] = out["encoder_states"][0][:, 0].numpy()
        # subsampler alone
        sub, _ = enc.subsample(feats.unsqueeze(0), torch.tensor([feats.size(0)]))
        enc_gold[f"sub_c{chunk}"] = sub[:, 0].numpy()
    # batched + padded (offline generator shape, config 3 style): B=2, second one shorter
    enc = build_reference_encoder(cfg, sd, 8, 8)
    fb = torch.zeros(2, feats.size(0), 80)
    fb[0] = feats
    fb[1, :150] = feats[:150]
    lens = torch.tensor([feats.size(0), 150])
    out = enc._forward(fb, lens)
    orc = StreamSpeechOracle(cfg, sd, None, gcmvn, chunk_size=8, conv_chunk_size=8)
    mine = orc.encoder(fb, lens)
    report["encoder_batched"] = maxdiff(out["encoder_out"][0], mine["encoder_out"][0])
    assert report["encoder_batched"] < 2e-5, report
    enc_gold["batched_out"] = out["encoder_out"][0].numpy()
    enc_gold["batched_lens"] = lens.numpy()
    np.savez_compressed(os.path.join(GOLD, "encoder.npz"), **enc_gold)

    # ------------------------------------------------------------------ decoders
    mods = build_reference_decoders(cfg, sd)
    enc = build_reference_encoder(cfg, sd, 8, 8)
    enc_out = enc._forward(feats.unsqueeze(0), torch.tensor([feats.size(0)]))
    eo = enc_out["encoder_out"][0]  # [T,1,256]
    orc = StreamSpeechOracle(cfg, sd, None, gcmvn, chunk_size=8)
    dec_gold = {"enc_out": eo[:, 0].numpy()}
    RefCTC, RefUnitCTC = build_reference_ctc_generators()
    fake = FakeModel(mods)
    for name in ("source_unigram", "ctc_target_unigram"):
        gen = RefCTC(FakeDict(6000), [fake])
        hyp = gen.generate(enc_out, aux_task_name=name)
        mine = orc.ctc_greedy(name, eo)
        assert hyp[0][0]["tokens"].tolist() == mine[0]["tokens"], name
        assert hyp[0][0]["index"] == mine[0]["index"], name
        assert hyp[0][0]["org_tokens"].tolist() == mine[0]["org_tokens"], name
        dec_gold[f"ctc_{name}_argmax"] = hyp[0][0]["org_tokens"].numpy()
        dec_gold[f"ctc_{name}_tokens"] = hyp[0][0]["tokens"].numpy()
        dec_gold[f"ctc_{name}_index"] = np.array(hyp[0][0]["index"])
        report[f"ctc_{name}_ntok"] = len(mine[0]["tokens"])
    # MT decoder: teacher-forced features/logits on a fixed token sequence
    toks = torch.tensor([[2, 17, 256, 4099, 31, 5, 977]])
    ref_logits, _ = mods["mt"](toks, encoder_out=enc_out)
    ref_feats, _ = mods["mt"](toks, encoder_out=enc_out, features_only=True)
    report["mt_logits"] = maxdiff(ref_logits, orc.mt_logits(toks, eo))
    report["mt_feats"] = maxdiff(ref_feats, orc.mt_features(toks, eo))
    assert report["mt_logits"] < 5e-5 and report["mt_feats"] < 2e-5, report
    dec_gold["mt_tokens"] = toks.numpy()
    dec_gold["mt_logits_last"] = ref_logits[0, -1].numpy()
    dec_gold["mt_feats"] = ref_feats[0].numpy()
    # with a trailing pad (whole_word path, agent:576-591)
    toks_pad = torch.tensor([[2, 17, 256, 4099, 1]])
    ref_feats_pad, _ = mods["mt"](toks_pad, encoder_out=enc_out, features_only=True)
    report["mt_feats_pad"] = maxdiff(ref_feats_pad, orc.mt_features(toks_pad, eo))
    assert report["mt_feats_pad"] < 2e-5, report
    dec_gold["mt_tokens_pad"] = toks_pad.numpy()
    dec_gold["mt_feats_pad"] = ref_feats_pad[0].numpy()
    # T2U encoder + unit decoder on the MT features
    x = ref_feats.transpose(0, 1)  # [S,1,512]
    t2u_ref = mods["t2u"](x, None)
    t2u_mine = orc.t2u_encoder(x, None)
    report["t2u"] = maxdiff(t2u_ref["encoder_out"][0], t2u_mine)
    assert report["t2u"] < 2e-5, report
    unit_logits_ref, _ = mods["unit"](None, encoder_out=t2u_ref)
    unit_logits_mine = orc.unit_decoder_logits(t2u_mine, None)
    report["unit_logits"] = maxdiff(unit_logits_ref, unit_logits_mine)
    assert report["unit_logits"] < 1e-4, report
    gen = RefUnitCTC(FakeDict(cfg.unit_vocab, blank=cfg.unit_blank), [fake])
    hyp = gen.generate(t2u_ref, prefix=None)
    mine = orc.unit_ctc_greedy(unit_logits_mine)
    assert hyp[0][0]["tokens"].tolist() == mine[0]["tokens"]
    dec_gold["t2u_out"] = t2u_ref["encoder_out"][0][:, 0].numpy()
    dec_gold["unit_logits_first"] = unit_logits_ref[0, :4].numpy()
    dec_gold["unit_argmax"] = hyp[0][0]["org_tokens"].numpy()
    dec_gold["unit_tokens"] = hyp[0][0]["tokens"].numpy()
    report["unit_ntok"] = len(mine[0]["tokens"])
    # padded T2U input (whole_word): pad mask on the last position
    pm = toks_pad.eq(1)
    xp = ref_feats_pad.transpose(0, 1)
    t2u_ref_p = mods["t2u"](xp, pm)
    t2u_mine_p = orc.t2u_encoder(xp, pm)
    report["t2u_pad"] = maxdiff(t2u_ref_p["encoder_out"][0], t2u_mine_p)
    ul_ref_p, _ = mods["unit"](None, encoder_out=t2u_ref_p)
    ul_mine_p = orc.unit_decoder_logits(t2u_mine_p, pm)
    report["unit_logits_pad"] = maxdiff(ul_ref_p, ul_mine_p)
    assert report["t2u_pad"] < 2e-5 and report["unit_logits_pad"] < 1e-4, report
    dec_gold["unit_argmax_pad"] = ul_ref_p[0].argmax(-1).numpy()
    np.savez_compressed(os.path.join(GOLD, "decoders.npz"), **dec_gold)

    # ------------------------------------------------------------------ vocoder
    from agent.tts.codehifigan import CodeGenerator
    vc = VocoderConfig()
    vsd_wn = synth.make_vocoder_state_dict(vc, seed=1, weight_norm=True)
    gen = CodeGenerator(vc.to_json_dict())
    gen.load_state_dict(vsd_wn, strict=True)
    gen.eval()
    gen.remove_weight_norm()
    code = torch.randint(0, 1000, (1, 24), generator=torch.Generator().manual_seed(3))
    wav_ref, dur_ref = gen(code=code, dur_prediction=True)
    orc_v = StreamSpeechOracle(cfg, sd, vsd_wn, gcmvn)
    wav_mine, dur_mine = orc_v.vocoder(code[0].tolist(), True)
    assert torch.equal(dur_ref, dur_mine)
    report["vocoder_wav"] = maxdiff(wav_ref.squeeze(), wav_mine)
    assert report["vocoder_wav"] < 1e-5, report
    # plain-weight checkpoint must give the same thing
    orc_v2 = StreamSpeechOracle(cfg, sd, synth.make_vocoder_state_dict(vc, seed=1, weight_norm=False), gcmvn)
    assert maxdiff(orc_v2.vocoder(code[0].tolist(), True)[0], wav_mine) < 1e-5
    np.savez_compressed(os.path.join(GOLD, "vocoder.npz"), code=code.numpy(), dur=dur_ref.numpy(),
                        wav=wav_ref.squeeze().numpy().astype(np.float32))

    for k, v in report.items():
        print(f"{k:40s} {v}")
    print("golden fixtures written to", GOLD)


if __name__ == "__main__":
    main()

"""Model dimensions of the StreamSpeech simultaneous S2ST checkpoint family.

Every number here is a *default* taken from the reference's training recipe and arch
functions; a real checkpoint's `cfg` / vocoder `config.json` overrides them:

* encoder / decoder sizes: researches/ctc_unity/train_scripts/train.simul-s2st.sh:11-31,
  researches/ctc_unity/models/streamspeech_model.py:418-430,
  fairseq/fairseq/models/speech_to_speech/s2s_transformer.py:619-669
* multitask heads: configs/fr-en/config_mtl_asr_st_ctcst.yaml:1-36
* unit dictionary = 4 specials + 1000 units + <blank> (researches/ctc_unity/tasks/speech_to_speech_ctc.py:15-17)
* vocoder: public unit-HiFiGAN config (not in the reference tree; read from JSON at run time)
"""
from __future__ import annotations

import ctypes
import json
from dataclasses import dataclass, field, asdict
from typing import List


@dataclass
class VocoderConfig:
    upsample_rates: List[int] = field(default_factory=lambda: [5, 4, 4, 2, 2])
    upsample_kernel_sizes: List[int] = field(default_factory=lambda: [11, 8, 8, 4, 4])
    upsample_initial_channel: int = 512
    resblock_kernel_sizes: List[int] = field(default_factory=lambda: [3, 7, 11])
    resblock_dilation_sizes: List[List[int]] = field(
        default_factory=lambda: [[1, 3, 5], [1, 3, 5], [1, 3, 5]]
    )
    num_embeddings: int = 1000
    embedding_dim: int = 128
    model_in_dim: int = 128
    dur_hidden: int = 128  # dur_predictor_params.var_pred_hidden_dim
    dur_kernel: int = 3  # dur_predictor_params.var_pred_kernel_size

    @property
    def hop(self) -> int:
        h = 1
        for u in self.upsample_rates:
            h *= u
        return h

    def to_json_dict(self) -> dict:
        """Same schema as the reference vocoder `config.json` (agent/tts/codehifigan.py:9-33)."""
        return {
            "resblock": "1",
            "upsample_rates": self.upsample_rates,
            "upsample_kernel_sizes": self.upsample_kernel_sizes,
            "upsample_initial_channel": self.upsample_initial_channel,
            "resblock_kernel_sizes": self.resblock_kernel_sizes,
            "resblock_dilation_sizes": self.resblock_dilation_sizes,
            "num_embeddings": self.num_embeddings,
            "embedding_dim": self.embedding_dim,
            "model_in_dim": self.model_in_dim,
            "dur_predictor_params": {
                "encoder_embed_dim": self.embedding_dim,
                "var_pred_hidden_dim": self.dur_hidden,
                "var_pred_kernel_size": self.dur_kernel,
                "var_pred_dropout": 0.5,
            },
        }

    @staticmethod
    def from_json_dict(d: dict) -> "VocoderConfig":
        dp = d.get("dur_predictor_params") or {}
        return VocoderConfig(
            upsample_rates=list(d["upsample_rates"]),
            upsample_kernel_sizes=list(d["upsample_kernel_sizes"]),
            upsample_initial_channel=int(d["upsample_initial_channel"]),
            resblock_kernel_sizes=list(d["resblock_kernel_sizes"]),
            resblock_dilation_sizes=[list(x) for x in d["resblock_dilation_sizes"]],
            num_embeddings=int(d["num_embeddings"]),
            embedding_dim=int(d["embedding_dim"]),
            model_in_dim=int(d.get("model_in_dim", 80)),
            dur_hidden=int(dp.get("var_pred_hidden_dim", 128)),
            dur_kernel=int(dp.get("var_pred_kernel_size", 3)),
        )

    @staticmethod
    def from_json_file(path: str) -> "VocoderConfig":
        with open(path) as f:
            return VocoderConfig.from_json_dict(json.load(f))


@dataclass
class ModelConfig:
    # front-end
    feat_dim: int = 80
    # chunk-Conformer encoder
    enc_dim: int = 256
    enc_ffn: int = 2048
    enc_heads: int = 4
    enc_layers: int = 12
    dw_kernel: int = 31
    conv_channels: int = 1024  # subsampler mid channels
    conv_kernel: int = 5
    max_source_positions: int = 6000
    # CTC heads (ASR: source_unigram, ST: ctc_target_unigram) and MT dictionary
    src_vocab: int = 6000
    tgt_vocab: int = 6000
    # first-pass MT decoder (target_unigram)
    mt_dim: int = 512
    mt_ffn: int = 2048
    mt_heads: int = 8
    mt_layers: int = 4
    # T2U ("synthesizer") encoder
    t2u_layers: int = 2
    # NAR CTC unit decoder
    unit_dim: int = 512
    unit_ffn: int = 2048
    unit_heads: int = 8
    unit_layers: int = 2
    unit_vocab: int = 1005  # 4 specials + 1000 units + <blank>
    ctc_upsample_rate: int = 25
    # dictionary conventions (fairseq Dictionary: <s>=0 <pad>=1 </s>=2 <unk>=3)
    bos: int = 0
    pad: int = 1
    eos: int = 2
    unk: int = 3
    uni_encoder: bool = True  # --uni-encoder: causal T2U encoder
    vocoder: VocoderConfig = field(default_factory=VocoderConfig)

    @property
    def unit_blank(self) -> int:
        return self.unit_vocab - 1

    @staticmethod
    def from_state_dict(sd) -> "ModelConfig":
        """Infer the dimensions from a fairseq StreamSpeech state dict (checkpoint `cfg` is hydra-pickled and
        not loadable without fairseq; every size is recoverable from tensor shapes)."""
        c = ModelConfig()
        c.conv_channels, c.feat_dim, c.conv_kernel = (int(x) for x in sd["encoder.subsample.conv_layers.0.weight"].shape)
        c.enc_dim = int(sd["encoder.linear.weight"].shape[0])
        c.enc_ffn = int(sd["encoder.conformer_layers.0.ffn1.w_1.weight"].shape[0])
        c.enc_heads = int(sd["encoder.conformer_layers.0.self_attn.pos_bias_u"].shape[0])
        c.dw_kernel = int(sd["encoder.conformer_layers.0.conv_module.depthwise_conv.weight"].shape[-1])
        n = lambda prefix: 1 + max(int(k[len(prefix):].split(".")[0]) for k in sd if k.startswith(prefix))  # noqa: E731
        c.enc_layers = n("encoder.conformer_layers.")
        c.src_vocab = int(sd["source_unigram_decoder.proj.weight"].shape[0])
        c.tgt_vocab, c.mt_dim = (int(x) for x in sd["target_unigram_decoder.embed_tokens.weight"].shape)
        c.mt_ffn = int(sd["target_unigram_decoder.layers.0.fc1.weight"].shape[0])
        c.mt_heads = c.mt_dim // 64
        c.mt_layers = n("target_unigram_decoder.layers.")
        c.t2u_layers = n("synthesizer_encoder.layers.")
        c.unit_vocab, c.unit_dim = (int(x) for x in sd["decoder.embed_tokens.weight"].shape)
        c.unit_ffn = int(sd["decoder.layers.0.fc1.weight"].shape[0])
        c.unit_heads = c.unit_dim // 64
        c.unit_layers = n("decoder.layers.")
        return c

    def tiny(self) -> "ModelConfig":
        """A shrunken variant with the same structure (used by fast CPU tests)."""
        c = ModelConfig(**{k: v for k, v in asdict(self).items() if k != "vocoder"})
        c.enc_layers = 2
        c.mt_layers = 1
        c.t2u_layers = 1
        c.unit_layers = 1
        c.vocoder = VocoderConfig()
        return c

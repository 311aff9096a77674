"""CPU: the index arithmetic of the multi-stream pool's batched subsampler (engine_pool.inc pool_geom + the canonical logical
positions of a group; kernels_gemm.cu load_a4 masking rules) restated in numpy and checked against the oracle's chunk-causal
subsampler on the full prefix: the rows [a0, T) computed from the gathered fbank window with SHIFTED logical positions must equal
the rows of the full computation, for every chunk setting the agents use and for streams at different absolute positions."""
import math

import numpy as np
import pytest
import torch

from oracle.streamspeech_oracle import StreamSpeechOracle
from streamspeech_b200 import synth
from streamspeech_b200.config import ModelConfig

torch.set_grad_enabled(False)


def pool_geom(F, T_final_prev, half=2):
    T1 = (F - 1) // 2 + 1
    T = (T1 - 1) // 2 + 1
    a0 = min(T_final_prev, T)
    t1_lo = max(0, 2 * a0 - half)
    f_lo = max(0, 2 * t1_lo - half)
    return dict(F=F, T1=T1, T=T, a0=a0, nA=T - a0, t1_lo=t1_lo, n1=T1 - t1_lo, f_lo=f_lo, nf=F - f_lo)


def conv_window(x_win, W, b, *, L_in, L_rows, t_offset, x_row0, chunk, k=5, stride=2, pad_left=2):
    """gemm_conv's A operand (kernels_gemm.cu load_a4) + GLU epilogue for one batch element; x_win [x_rows][C]; W [Cout][C][k] (torch)."""
    x_rows, C = x_win.shape
    out = np.zeros((L_rows, W.shape[0]), dtype=np.float64)
    for r in range(L_rows):
        t = r + t_offset
        center = t * stride
        for tap in range(k):
            pos = center + tap - pad_left
            if pos < 0 or pos >= L_in:
                continue
            if chunk > 0 and pos >= (center // chunk + 1) * chunk:
                continue
            pr = pos - x_row0
            if pr < 0 or pr >= x_rows:
                continue
            out[r] += W[:, :, tap].astype(np.float64) @ x_win[pr].astype(np.float64)
        out[r] += b
    half = W.shape[0] // 2
    return (out[:, :half] * (1.0 / (1.0 + np.exp(-out[:, half:])))).astype(np.float32)


@pytest.mark.parametrize("attn,conv", [(4, 4), (8, 8), (4, 8), (16, 16), (12, 8)])
def test_canonical_window_subsampler_equals_full_prefix(attn, conv):
    cfg = ModelConfig().tiny()
    sd = synth.make_model_state_dict(cfg, 0)
    o = StreamSpeechOracle(cfg, sd, None, None, chunk_size=attn, conv_chunk_size=conv)
    G = attn * conv // math.gcd(attn, conv)
    W0, b0 = sd["encoder.subsample.conv_layers.0.weight"].numpy(), sd["encoder.subsample.conv_layers.0.bias"].numpy()
    W1, b1 = sd["encoder.subsample.conv_layers.1.weight"].numpy(), sd["encoder.subsample.conv_layers.1.bias"].numpy()
    feats = torch.randn(800, cfg.feat_dim, generator=torch.Generator().manual_seed(3))
    T_final = 0
    step = 10 * attn  # fbank frames per call (4 per encoder frame, 10 ms each)
    checked = 0
    for F in list(range(step - 2, 800, step)) + [800]:
        g = pool_geom(F, T_final)
        if g["nA"] > 0:
            a0c = 0 if g["a0"] == 0 else G
            shift = g["a0"] - a0c
            win = feats[g["f_lo"]:F].numpy()
            c1 = conv_window(win, W0, b0, L_in=F - 4 * shift, L_rows=g["n1"], t_offset=g["t1_lo"] - 2 * shift, x_row0=g["f_lo"] - 4 * shift, chunk=conv)
            x0 = conv_window(c1, W1, b1, L_in=g["T1"] - 2 * shift, L_rows=g["nA"], t_offset=a0c, x_row0=g["t1_lo"] - 2 * shift, chunk=conv)
            ref, _ = o.subsample(feats[:F].unsqueeze(0), torch.tensor([F]))
            ref = ref[:, 0].numpy()
            assert ref.shape[0] == g["T"]
            assert float(np.abs(x0 - ref[g["a0"]:]).max()) < 2e-5, (F, g)
            checked += 1
        T_final = max(T_final, min(g["T"], (F // (4 * G)) * G))
    assert checked >= 4

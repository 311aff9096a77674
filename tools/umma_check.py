#!/usr/bin/env python
"""Parity of the tcgen05 split-bf16 GEMM (kernels_umma.cu) against fp32: single GEMMs, then the vocoder and the
unit decoder with the tensor-core path switched on.  Run under `timeout` (a wrong mbarrier protocol would spin)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from streamspeech_b200 import synth
from streamspeech_b200.config import ModelConfig
from streamspeech_b200.engine import Engine

torch.set_grad_enabled(False)
cfg = ModelConfig(); cfg.enc_layers = 3
e = Engine(cfg, synth.make_model_state_dict(cfg, 0), synth.make_vocoder_state_dict(cfg.vocoder, 1), None)
g = torch.Generator().manual_seed(0)
res = {}
ok = True
for (M, K, N, act) in [(128, 32, 16, 0), (128, 64, 128, 0), (256, 96, 32, 1), (300, 2816, 256, 0), (1000, 512, 64, 2), (775, 2048, 512, 0), (129, 352, 1536, 0)]:
    x = torch.randn(M, K, generator=g); w = torch.randn(N, K, generator=g) / K ** 0.5; b = torch.randn(N, generator=g)
    ref = torch.nn.functional.linear(x.double(), w.double(), b.double())
    ref = {0: ref, 1: torch.relu(ref), 2: torch.nn.functional.silu(ref)}[act].float()
    f32 = e.op_linear(x.cuda(), w.cuda(), b.cuda(), act).cpu()
    for pieces in (2, 3):
        got = e.op_linear_umma(x.cuda(), w.cuda(), b.cuda(), act, pieces).cpu()
        d = float((got - ref).abs().max()); d32 = float((f32 - ref).abs().max())
        res[f"gemm_{M}x{K}x{N}_act{act}_p{pieces}"] = {"umma_maxdiff": d, "fp32_simt_maxdiff": d32}
        tol = 2e-4 if pieces == 3 else 2e-3
        if not d < tol:
            ok = False
        print(M, K, N, act, "pieces", pieces, "maxdiff", d, "(fp32 SIMT:", d32, ")", "OK" if d < tol else "FAIL")
# timing: the vocoder's heaviest conv shape
M, C, k = 2500, 256, 11
x = torch.randn(M, C * k, device="cuda"); w = torch.randn(C, C * k, device="cuda") / (C * k) ** 0.5; b = torch.zeros(C, device="cuda")
for name, fn in (("simt_fp32", lambda: e.op_linear(x, w, b)), ("umma_p2", lambda: e.op_linear_umma(x, w, b, 0, 2)), ("umma_p3", lambda: e.op_linear_umma(x, w, b, 0, 3))):
    for _ in range(3): fn()
    s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); s.record()
    for _ in range(20): fn()
    t.record(); torch.cuda.synchronize()
    us = s.elapsed_time(t) / 20 * 1e3
    res["time_" + name] = {"us": us, "tflops_fp32_equiv": 2.0 * M * C * C * k / us / 1e6}
    print(name, us, "us", res["time_" + name]["tflops_fp32_equiv"], "TFLOP/s (fp32-equivalent)")
# vocoder end to end
gold = np.load(os.path.join(ROOT, "tests", "golden", "vocoder.npz"))
codes = torch.from_numpy(gold["code"][0].astype(np.int64)).cuda()
dur, cum = e.vocoder_durations(codes, True); total = int(cum[-1].item())
base = e.vocoder_generate(total, 0, total, 0).cpu()
for p in (2, 3):
    e.set_option("umma_vocoder", p)
    wav = e.vocoder_generate(total, 0, total, 0).cpu()
    d = float((wav - torch.from_numpy(gold["wav"])).abs().max())
    res[f"vocoder_umma_p{p}"] = {"maxdiff_vs_reference_fixture": d, "maxdiff_vs_fp32_path": float((wav - base).abs().max())}
    print("vocoder pieces", p, res[f"vocoder_umma_p{p}"])
    ok &= d < 1e-3
e.set_option("umma_vocoder", 0)
# unit decoder with tensor-core linears
dec = np.load(os.path.join(ROOT, "tests", "golden", "decoders.npz"))
feats = torch.from_numpy(dec["mt_feats"]).cuda()
big = feats.repeat(6, 1).contiguous()  # 42 tokens -> 1050 unit positions (M >= 128 takes the tcgen05 path)
ref = e.t2u_unit_decode(big, debug=True)
e.set_option("umma_linear", 3)
got = e.t2u_unit_decode(big, debug=True)
e.set_option("umma_linear", 0)
d = float((got["logits"] - ref["logits"]).abs().max())
same = got["argmax"].tolist() == ref["argmax"].tolist()
res["unit_decoder_umma_p3"] = {"logits_maxdiff_vs_fp32_path": d, "argmax_equal": same}
print("unit decoder umma p3: logits maxdiff", d, "argmax equal", same)
ok &= same and d < 1e-3
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "umma_check.json"), "w"), indent=1)
print("UMMA_CHECK", "PASS" if ok else "FAIL")
sys.exit(0 if ok else 1)

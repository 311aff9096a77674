"""Multi-stream pool (BASELINE configs[3]): every stream of a batched step must produce exactly what the
single-stream agent produces for the same utterance -- tokens bit-exact at every call, encoder rows within the fp32 tolerance."""
import argparse

import pytest
import torch

pytestmark = pytest.mark.gpu

from streamspeech_b200 import synth  # noqa: E402
from streamspeech_b200.config import ModelConfig  # noqa: E402

torch.set_grad_enabled(False)


def asr_args(seg_ms):
    return argparse.Namespace(model_path="synthetic", data_bin=".", config_yaml=None, multitask_config_yaml=None, sample_rate=16000, max_len=200,
                              force_finish=False, vocoder="synthetic", vocoder_cfg=None, dur_prediction=True, lagging_k1=0, lagging_k2=0,
                              segment_size=seg_ms, stride_n=1, stride_n2=1, unit_per_subword=15, source_segment_size=seg_ms,
                              vocoder_context="receptive-field", device_index=0)


@pytest.mark.parametrize("seg_ms,conv,heads", [(160, 4, 1), (320, 8, 2)])
def test_pool_streams_equal_single_stream_agents(seg_ms, conv, heads):
    from streamspeech_b200.agent import StreamSpeechASRAgent
    from streamspeech_b200.scheduler import PooledASRAgent, StreamPool, pushpop_many
    from streamspeech_b200.simuleval_compat import SpeechSegment

    single = StreamSpeechASRAgent(asr_args(seg_ms))
    eng = single.engine
    eng.set_chunk(seg_ms // 40, conv)
    # ragged on purpose: different lengths, one stream joins late, one is a single short chunk
    specs = [(3.1, 11, 0), (2.0, 12, 0), (4.05, 13, 2), (0.5, 14, 1), (2.72, 15, 0), (3.1, 11, 5)]
    n = 16 * seg_ms
    wavs = [synth.make_audio(sec, seed=seed) for sec, seed, _ in specs]
    # reference: every utterance alone through the single-stream agent; record tokens per call and final encoder rows
    ref_tokens, ref_enc = [], []
    for w in wavs:
        single.reset()
        calls = []
        last_enc = None
        for i in range(0, len(w), n):
            fin = i + n >= len(w)
            single.push(SpeechSegment(content=w[i:i + n].tolist(), sample_rate=16000, finished=fin))
            # policy() resets the agent on the final call: grab the encoder rows first through the same code path
            feat = single._features()
            if feat.size(0) > 0:
                enc = single._encode(feat)
                toks0 = single._ctc(0, enc)[0] if heads == 1 else None
                pair = single._ctc_pair(enc) if heads == 2 else None
                calls.append(toks0 if heads == 1 else (pair[0][0], pair[1][0]))
                last_enc = enc.clone()
            else:
                calls.append(None)
        ref_tokens.append(calls)
        ref_enc.append(last_enc)
    # pooled run: all streams concurrently, stream j starts `delay` rounds late
    pool = StreamPool(eng, n_slots=8, max_seconds=10, ctc_heads=heads)
    slots = [pool.acquire() for _ in specs]
    pos = [0] * len(specs)
    call_idx = [0] * len(specs)
    rnd = 0
    worst = 0.0
    while any(p < len(w) for p, w in zip(pos, wavs)):
        active = [j for j, (_, _, delay) in enumerate(specs) if rnd >= delay and pos[j] < len(wavs[j])]
        for j in active:
            pool.push(slots[j], wavs[j][pos[j]:pos[j] + n])
            pos[j] += n
        pool.flush()
        for j in active:
            r = pool.results[slots[j]]
            want = ref_tokens[j][call_idx[j]]
            if want is None:
                assert r["T"] == 0
            elif heads == 1:
                assert r["ctc"][0][0] == want, (j, call_idx[j])
            else:
                assert (r["ctc"][0][0], r["ctc"][1][0]) == want, (j, call_idx[j])
            call_idx[j] += 1
        rnd += 1
    assert max(pool.rows_per_step) >= 4  # streams really were batched
    cfg = ModelConfig()
    for j in range(len(specs)):
        info = eng.pool_info(slots[j])
        T = ref_enc[j].shape[0]
        # view the slot's encoder rows through a tensor that aliases the pool memory
        got = torch.empty(T, cfg.enc_dim, device="cuda")
        torch.cuda.synchronize()
        src = _alias(info["enc_out_ptr"], T * cfg.enc_dim)
        got.copy_(src.view(T, cfg.enc_dim))
        worst = max(worst, float((got - ref_enc[j]).abs().max()))
    assert worst < 2e-4, worst
    single.engine.close()


def _alias(ptr, numel):
    """float32 CUDA tensor over existing device memory (test helper)"""
    class _Arr:
        __cuda_array_interface__ = {"shape": (numel,), "typestr": "<f4", "data": (ptr, False), "version": 3}
    return torch.as_tensor(_Arr(), device="cuda")


def test_pooled_agents_batch_behind_push_pop():
    """The scheduler behind push() / pop(): agents fed round-robin produce the single-stream agent's text, in one engine step per round."""
    from streamspeech_b200.agent import StreamSpeechASRAgent
    from streamspeech_b200.scheduler import PooledASRAgent, StreamPool, pushpop_many
    from streamspeech_b200.simuleval_compat import SpeechSegment

    single = StreamSpeechASRAgent(asr_args(160))
    pool = StreamPool(single.engine, n_slots=4, max_seconds=10, ctc_heads=1)
    agents = [PooledASRAgent(pool, single.dict["source_unigram"]) for _ in range(3)]
    wavs = [synth.make_audio(2.4, seed=s) for s in (21, 22, 23)]
    n = 2560
    texts = [[] for _ in agents]
    for i in range(0, 38400, n):
        fin = i + n >= 38400
        segs = [SpeechSegment(content=w[i:i + n].tolist(), sample_rate=16000, finished=fin) for w in wavs]
        outs = pushpop_many(agents, segs)
        for j, o in enumerate(outs):
            texts[j].append("" if o.is_empty else o.content)
    assert pool.steps == 38400 // n and all(r == 3 for r in pool.rows_per_step)
    for j, w in enumerate(wavs):
        single.reset()
        ref = []
        for i in range(0, len(w), n):
            o = single.pushpop(SpeechSegment(content=w[i:i + n].tolist(), sample_rate=16000, finished=i + n >= len(w)))
            ref.append("" if o.is_empty else o.content)
        assert texts[j] == ref, j
    single.engine.close()


def test_pool_edge_cases_and_errors():
    """Empty / too-short inputs give empty results, misuse and capacity overruns fail loudly (negative status -> EngineError)."""
    from streamspeech_b200.agent import StreamSpeechASRAgent
    from streamspeech_b200.engine import EngineError
    from streamspeech_b200.scheduler import StreamPool

    single = StreamSpeechASRAgent(asr_args(160))
    eng = single.engine
    pool = StreamPool(eng, n_slots=3, max_seconds=2, ctc_heads=1)
    a, b, c = (pool.acquire() for _ in range(3))
    # nothing pushed at all / fewer samples than one fbank frame: T = 0, no tokens
    pool.push(a, [])
    pool.push(b, [0.01] * 300)
    pool.flush()
    assert pool.results[a]["T"] == 0 and pool.results[a]["ctc"][0] == ([], [])
    assert pool.results[b]["T"] == 0 and pool.results[b]["ctc"][0] == ([], [])
    # a stream that got nothing new in a round keeps its result (same T) while the other advances
    w = synth.make_audio(1.0, seed=3)
    pool.push(c, w[:8000])
    pool.flush()
    t1 = pool.results[c]["T"]
    pool.push(c, [])
    pool.reset(b)
    pool.push(b, w[:8000])
    pool.flush()
    assert pool.results[c]["T"] == t1 and pool.results[b]["T"] == t1
    with pytest.raises(EngineError, match="listed twice"):
        eng.pool_step([a, a])
    with pytest.raises(EngineError, match="bad pool slot"):
        eng.pool_step([7])
    with pytest.raises(EngineError, match="capacity"):
        eng.pool_push_audio(a, torch.zeros(16000 * 3))  # the pool was sized for 2 s
    with pytest.raises(EngineError, match="already has a stream pool"):
        eng.pool_create(2, 2)
    single.engine.close()

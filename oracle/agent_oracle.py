"""CPU restatement of `StreamSpeechS2STAgent.policy()` (agent/speech_to_speech.streamspeech.agent.py:422-770)
and of the streaming-ASR agent's policy (agent/speech_to_text.asr.streamspeech.agent.py:385-433).

TEST INFRASTRUCTURE ONLY (see oracle/streamspeech_oracle.py header).  It keeps the reference's
semantics, including what makes it slow: every call recomputes fbank, encoder, CTC heads,
the MT decoder (no incremental state, agent:179), T2U, unit decoder and vocoder on the WHOLE
prefix.  This is the "port" CPU baseline that bench.py times.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional

import torch

from .streamspeech_oracle import StreamSpeechOracle, online_features


@dataclass
class OracleAction:
    kind: str  # "read" | "write"
    wav: Optional[List[float]] = None
    finished: bool = False  # WriteAction.finished
    seg_finished: bool = False  # SpeechSegment.finished
    # extra trace for parity tests (not part of the reference's return value)
    trace: dict = field(default_factory=dict)


class OracleS2STAgent:
    def __init__(self, oracle: StreamSpeechOracle, segment_ms: int = 320, lagging_k1: int = 0, stride_n: int = 1,
                 dur_prediction: bool = True, max_len_b: int = 100, is_word_start=None):
        """is_word_start(token_id) -> bool: `generator_mt.tgt_dict[id].startswith("\u2581")` (agent:545-548), needed by the
        whole-word path (segment >= 640 ms); pass the predicate of the dictionary in use."""
        self.o = oracle
        self.is_word_start = is_word_start
        self.segment_ms = segment_ms
        self.lagging_k1 = lagging_k1
        self.stride_n = stride_n
        self.dur_prediction = dur_prediction
        self.max_len_b = max_len_b
        self.whole_word = segment_ms >= 640  # agent:207-210
        oracle.set_chunk(segment_ms // 40)  # agent:395-413
        self.reset()

    def reset(self):  # agent:328-347
        self.source: List[float] = []
        self.source_finished = False
        self.target_finished = False
        self.tgt_subwords_indices: Optional[List[int]] = None
        self.src_ctc_prefix_length = 0
        self.tgt_ctc_prefix_length = 0
        self.prev_output_tokens_mt: Optional[List[int]] = None
        self.unit: Optional[List[int]] = None
        self.unfinished_wav = None

    def push(self, samples, finished: bool = False):  # SimulEval agents/agent.py:71-83, states.py:33-47
        self.source.extend(samples)
        self.source_finished = finished

    def _finish_or_read(self, trace):
        if not self.source_finished:
            return OracleAction("read", trace=trace)
        return OracleAction("write", wav=[], finished=True, seg_finished=True, trace=trace)

    def policy(self) -> OracleAction:
        o, c = self.o, self.o.cfg
        trace = {}
        feature = online_features(torch.tensor(self.source, dtype=torch.float32), o.gcmvn)
        if feature.size(0) == 0 and not self.source_finished:
            return OracleAction("read", trace=trace)
        enc = o.encoder(feature.unsqueeze(0), torch.tensor([feature.size(0)]))
        eo = enc["encoder_out"][0]
        asr = o.ctc_greedy("source_unigram", eo)[0]
        st = o.ctc_greedy("ctc_target_unigram", eo)[0]
        trace["asr_tokens"], trace["st_tokens"] = asr["tokens"], st["tokens"]
        trace["enc_frames"] = eo.size(0)
        if not self.source_finished:  # agent:480-509
            src_len, tgt_len = len(asr["tokens"]), len(st["tokens"])
            if src_len < self.src_ctc_prefix_length + self.stride_n or tgt_len < self.tgt_ctc_prefix_length + self.stride_n:
                return OracleAction("read", trace=trace)
            self.src_ctc_prefix_length = max(src_len, self.src_ctc_prefix_length)
            self.tgt_ctc_prefix_length = max(tgt_len, self.tgt_ctc_prefix_length)
            subword_tokens = ((tgt_len - self.lagging_k1) // self.stride_n) * self.stride_n
            if self.whole_word:
                subword_tokens += 1
            new_subword_tokens = subword_tokens - len(self.tgt_subwords_indices) if self.tgt_subwords_indices is not None else subword_tokens
            if new_subword_tokens < 1:
                return OracleAction("read", trace=trace)
        else:
            new_subword_tokens = -1
        new_subword_tokens = int(new_subword_tokens)
        trace["new_subword_tokens"] = new_subword_tokens

        # 1. MT decoder (agent:520-538)
        hyp = o.mt_greedy(eo, self.tgt_subwords_indices, new_subword_tokens, max_len_b=self.max_len_b)
        tgt = hyp[:-1] if hyp[-1] == c.eos else hyp
        finalized_tokens = hyp
        if self.whole_word:  # agent:540-574 (the incremental-state surgery there is dead: use_incremental_states=False, agent:179)
            j = 999999
            if not self.source_finished:
                if self.is_word_start is None:
                    raise ValueError("whole_word path needs is_word_start (dictionary predicate)")
                for j in range(len(tgt) - 1, -1, -1):
                    if self.is_word_start(tgt[j]):
                        break
                tgt = tgt[:j]
                finalized_tokens = finalized_tokens[:j]
                if j == 0:
                    return OracleAction("read", trace=trace)
        max_tgt_len = len(finalized_tokens) + (1 if self.whole_word else 0)
        prev = [c.pad] * max_tgt_len
        prev[0] = c.eos
        tmp = finalized_tokens[:-1] if finalized_tokens[-1] == c.eos else finalized_tokens
        prev[1: len(tmp) + 1] = tmp
        trace["mt_tokens"] = list(tmp)
        if self.tgt_subwords_indices is not None and self.tgt_subwords_indices == tgt:  # agent:609-626
            return self._finish_or_read(trace)
        self.tgt_subwords_indices = tgt
        if not self.source_finished and self.prev_output_tokens_mt is not None:  # agent:629-636
            if self.prev_output_tokens_mt == prev or len(prev) <= len(self.prev_output_tokens_mt):
                return OracleAction("read", trace=trace)
        self.prev_output_tokens_mt = prev
        prev_t = torch.tensor([prev], dtype=torch.long)
        x = o.mt_features(prev_t, eo).transpose(0, 1)  # agent:638-652
        pad_mask = prev_t.eq(c.pad) if prev_t.eq(c.pad).any() else None
        # 2. T2U encoder + CTC unit decoder (agent:662-689)
        t2u = o.t2u_encoder(x, pad_mask)
        logits = o.unit_decoder_logits(t2u, pad_mask)
        fin = o.unit_ctc_greedy(logits)[0]
        if len(fin["tokens"]) == 0:
            return self._finish_or_read(trace)
        tmp = fin["tokens"]
        if tmp[-1] == c.eos:
            tmp = tmp[:-1]
        # dictionary strings of the unit dict are "<s>", "<pad>", "</s>", "<unk>", "0".."999", "<blank>"
        # (agent:708-718): <s>/</s> are dropped, everything else is int(symbol)
        unit = [t - 4 for t in tmp if t not in (c.bos, c.eos)]
        trace["units"] = list(unit)
        cur_unit = unit if self.unit is None else unit[len(self.unit):]
        if len(unit) < 1 or len(cur_unit) < 1:
            return self._finish_or_read(trace)
        wav, dur = o.vocoder(unit, self.dur_prediction)  # agent:743-748
        cur_wav_length = int(dur[:, -len(cur_unit):].sum()) * 320
        new_wav = wav[-cur_wav_length:]
        trace["dur"] = dur.view(-1).tolist()
        self.unit = unit
        if self.source_finished and new_subword_tokens == -1:
            self.target_finished = True
            self.reset()  # resets the states too: both flags below read False afterwards (agent:759-770)
        return OracleAction("write", wav=new_wav.tolist(), finished=self.target_finished,
                            seg_finished=self.source_finished, trace=trace)


class OracleASRAgent:
    """agent/speech_to_text.asr.streamspeech.agent.py:385-433: fbank -> encoder -> ASR CTC -> text delta.
    Conv chunk = min(chunk, 16) there (N8), not the S2ST rule."""

    def __init__(self, oracle: StreamSpeechOracle, segment_ms: int = 160):
        self.o = oracle
        c = segment_ms // 40
        oracle.set_chunk(c, min(c, 16))
        self.reset()

    def reset(self):
        self.source: List[float] = []
        self.source_finished = False
        self.emitted = 0

    def push(self, samples, finished=False):
        self.source.extend(samples)
        self.source_finished = finished

    def policy(self):
        feature = online_features(torch.tensor(self.source, dtype=torch.float32), self.o.gcmvn)
        if feature.size(0) == 0 and not self.source_finished:
            return OracleAction("read")
        enc = self.o.encoder(feature.unsqueeze(0), torch.tensor([feature.size(0)]))
        toks = self.o.ctc_greedy("source_unigram", enc["encoder_out"][0])[0]["tokens"]
        new = toks[self.emitted:]
        self.emitted = len(toks)
        return OracleAction("write", trace={"asr_tokens": toks, "new": new}, finished=self.source_finished)


class OracleS2TTAgent:
    """agent/speech_to_text.s2tt.streamspeech.agent.py:381-545: fbank -> encoder -> ASR / ST CTC -> policy gate -> MT decoder
    WITH incremental states across policy() calls (generator_mt: use_incremental_states=True, max_len_a=1, max_len_b=200,
    :161-179) -> text delta."""

    def __init__(self, oracle: StreamSpeechOracle, segment_ms: int = 320, lagging_k1: int = 0, stride_n: int = 1, symbols=None,
                 max_decoder_positions: int = 1200):
        """symbols(token_id) -> dictionary string (generator_mt.tgt_dict[c])."""
        self.o = oracle
        self.lagging_k1, self.stride_n = lagging_k1, stride_n
        self.symbols = symbols if symbols is not None else (lambda t: str(t))
        self.max_decoder_positions = max_decoder_positions
        ch = segment_ms // 40
        oracle.set_chunk(ch, min(ch, 16))  # :359-366 (the ASR agent's rule, N8)
        self.reset()

    def reset(self):  # :300-321 (reset_incremental_states inside the try)
        self.source: List[float] = []
        self.source_finished = False
        self.target_finished = False
        self.tgt_subwords_indices: Optional[List[int]] = None
        self.src_ctc_prefix_length = 0
        self.tgt_ctc_prefix_length = 0
        self.tgt_text = ""
        self.state = self.o.mt_incremental_state()

    def push(self, samples, finished: bool = False):
        self.source.extend(samples)
        self.source_finished = finished

    def policy(self) -> OracleAction:
        o, c = self.o, self.o.cfg
        trace = {}
        feature = online_features(torch.tensor(self.source, dtype=torch.float32), o.gcmvn)
        if feature.size(0) == 0 and not self.source_finished:
            return OracleAction("read", trace=trace)
        enc = o.encoder(feature.unsqueeze(0), torch.tensor([feature.size(0)]))
        eo = enc["encoder_out"][0]
        asr = o.ctc_greedy("source_unigram", eo)[0]
        st = o.ctc_greedy("ctc_target_unigram", eo)[0]
        trace["asr_tokens"], trace["st_tokens"] = asr["tokens"], st["tokens"]
        if not self.source_finished:  # :437-465
            src_len, tgt_len = len(asr["tokens"]), len(st["tokens"])
            if src_len < self.src_ctc_prefix_length + self.stride_n or tgt_len < self.tgt_ctc_prefix_length + self.stride_n:
                return OracleAction("read", trace=trace)
            self.src_ctc_prefix_length = max(src_len, self.src_ctc_prefix_length)
            self.tgt_ctc_prefix_length = max(tgt_len, self.tgt_ctc_prefix_length)
            subword_tokens = ((tgt_len - self.lagging_k1) // self.stride_n) * self.stride_n
            new_subword_tokens = subword_tokens - len(self.tgt_subwords_indices) if self.tgt_subwords_indices is not None else subword_tokens
            if new_subword_tokens < 1:
                return OracleAction("read", trace=trace)
        else:
            new_subword_tokens = -1
        new_subword_tokens = int(new_subword_tokens)
        trace["new_subword_tokens"] = new_subword_tokens
        # max_len for -1: min(int(max_len_a * src_len + max_len_b), self.max_len - 1) with max_len_a = 1, max_len_b = 200 and
        # src_len = src_tokens.size(1) = number of fbank frames (sequence_generator.py:197-215)
        max_len_full = min(int(1 * feature.size(0) + 200), self.max_decoder_positions - 1)
        hyp = o.mt_greedy_incremental(self.state, eo, self.tgt_subwords_indices, new_subword_tokens, max_len_full)
        tgt = hyp[:-1] if hyp[-1] == c.eos else hyp
        trace["mt_tokens"] = list(tgt)
        if self.tgt_subwords_indices is not None and self.tgt_subwords_indices == tgt:  # :519-529
            if not self.source_finished:
                return OracleAction("read", trace=trace)
            return OracleAction("write", wav="", finished=True, trace=trace)
        self.tgt_subwords_indices = tgt
        text = " ".join(self.symbols(t) for t in tgt)  # :532-535
        new_text = text[len(self.tgt_text):]
        self.tgt_text = text
        if self.source_finished and new_subword_tokens == -1:
            self.target_finished = True
            self.reset()
        return OracleAction("write", wav=new_text, finished=self.target_finished, trace=trace)

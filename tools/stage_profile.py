#!/usr/bin/env python
"""Per-stage cost of one streamed utterance (wall time with a device sync around every engine call, kernel launches
per stage).  Diagnostic only: numbers are NOT bench values (the syncs remove CPU/GPU overlap)."""
import argparse
import collections
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from streamspeech_b200 import synth  # noqa: E402
from streamspeech_b200.agent import StreamSpeechS2STAgent  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--encoder-mode", default="cached")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "stage_profile.json"))
    a = ap.parse_args()
    torch.set_grad_enabled(False)
    agent = StreamSpeechS2STAgent(bench.agent_args(0, a.encoder_mode))
    eng = agent.engine
    stats = collections.defaultdict(lambda: [0, 0.0, 0])
    for name in ("fbank", "encoder", "encoder_stream_step", "ctc_greedy", "ctc_greedy_rows_pair", "mt_greedy", "mt_features", "t2u_unit_decode",
                 "vocoder_durations", "vocoder_generate"):
        fn = getattr(eng, name)

        def wrap(*args, _fn=fn, _name=name, **kw):
            torch.cuda.synchronize()
            l0 = eng.launch_count()
            t0 = time.perf_counter()
            r = _fn(*args, **kw)
            torch.cuda.synchronize()
            s = stats[_name]
            s[0] += 1
            s[1] += time.perf_counter() - t0
            s[2] += eng.launch_count() - l0
            return r
        setattr(eng, name, wrap)
    u = synth.make_audio(10.0, seed=1234).cuda()
    n = 5120

    def run():
        agent.reset()
        for i in range(0, u.numel(), n):
            end = min(i + n, u.numel())
            agent.step_resident(u, end, end >= u.numel())
    run()
    stats.clear()
    t0 = time.perf_counter()
    run()
    total = time.perf_counter() - t0
    rows = {k: {"calls": v[0], "ms": v[1] * 1e3, "launches": v[2], "us_per_launch": v[1] * 1e6 / max(v[2], 1)} for k, v in stats.items()}
    rows["_total_wall_ms"] = total * 1e3
    rows["_host_other_ms"] = total * 1e3 - sum(v[1] for v in stats.values()) * 1e3
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    json.dump(rows, open(a.out, "w"), indent=1)
    for k, v in sorted(rows.items(), key=lambda kv: -(kv[1]["ms"] if isinstance(kv[1], dict) else 0)):
        print(k, v)


if __name__ == "__main__":
    main()

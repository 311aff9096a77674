"""Stub-import loader for the reference's own module files (TEST INFRASTRUCTURE ONLY).

Only usable in the build container, where /root/reference exists.  It registers bare
package objects so the reference `__init__.py` files (which need omegaconf/hydra/...)
never run, then imports the leaf module files unchanged.  Used by oracle/gen_golden.py
to (a) validate the CPU restatement in oracle/streamspeech_oracle.py against the real
reference classes and (b) dump the golden vectors committed under tests/golden/.

Nothing in the product path, the GPU tests, smoke() or bench.py imports this file.
Recipe documented in SURVEY.md Appendix A.
"""
import os
import sys
import types

REF = os.environ.get("STREAMSPEECH_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF, "researches", "ctc_unity"))


_loaded = False


def load():
    """Register stubs; afterwards `import chunk_unity.modules.conformer_layer` etc. work."""
    global _loaded
    if _loaded:
        return
    if not available():
        raise RuntimeError(f"reference tree not found at {REF}")

    def stub(name, path):
        m = types.ModuleType(name)
        m.__path__ = [path]
        sys.modules[name] = m
        return m

    f = stub("fairseq", REF + "/fairseq/fairseq")
    fm = stub("fairseq.modules", REF + "/fairseq/fairseq/modules")
    fmo = stub("fairseq.models", REF + "/fairseq/fairseq/models")
    stub("fairseq.models.text_to_speech", REF + "/fairseq/fairseq/models/text_to_speech")
    stub("fairseq.models.speech_to_speech", REF + "/fairseq/fairseq/models/speech_to_speech")
    stub(
        "fairseq.models.speech_to_speech.modules",
        REF + "/fairseq/fairseq/models/speech_to_speech/modules",
    )
    for p in ("uni_unity", "chunk_unity", "ctc_unity"):
        stub(p, REF + "/researches/" + p)
        stub(p + ".modules", REF + "/researches/" + p + "/modules")
        stub(p + ".models", REF + "/researches/" + p + "/models")
    oc = types.ModuleType("omegaconf")
    oc.OmegaConf = type("OmegaConf", (), {"is_config": staticmethod(lambda o: False)})
    sys.modules["omegaconf"] = oc
    t = types.ModuleType("fairseq.models.transformer")
    t.TransformerConfig = object
    sys.modules["fairseq.models.transformer"] = t

    import fairseq.utils as U

    f.utils = U
    from fairseq.modules.gelu import gelu, gelu_accurate

    fm.gelu, fm.gelu_accurate = gelu, gelu_accurate
    from fairseq.models.fairseq_decoder import FairseqDecoder

    fmo.FairseqDecoder = FairseqDecoder
    from fairseq.models.fairseq_encoder import FairseqEncoder

    fmo.FairseqEncoder = FairseqEncoder
    from fairseq.models.fairseq_incremental_decoder import FairseqIncrementalDecoder

    fmo.FairseqIncrementalDecoder = FairseqIncrementalDecoder
    from fairseq.modules.layer_norm import LayerNorm

    fm.LayerNorm = LayerNorm
    from fairseq.modules.fairseq_dropout import FairseqDropout

    fm.FairseqDropout = FairseqDropout
    from fairseq.modules.positional_encoding import RelPositionalEncoding

    fm.RelPositionalEncoding = RelPositionalEncoding
    _loaded = True


_full = False


def load_full():
    """Extra stubs so the decoder-side reference classes import unchanged:
    TransformerDecoderBase / CTCTransformerUnitDecoder / UniTransformerEncoderNoEmb
    (researches/ctc_unity/modules/*.py) and CodeGenerator (agent/tts/codehifigan.py).
    VariancePredictor is exec'd from its own class source (fastspeech2.py imports hydra's II)."""
    global _full
    load()
    if _full:
        return
    import ast

    import torch.nn as nn

    f = sys.modules["fairseq"]
    fm = sys.modules["fairseq.modules"]
    d = types.ModuleType("fairseq.distributed")
    d.fsdp_wrap = lambda m, **kw: m
    sys.modules["fairseq.distributed"] = d
    f.distributed = d

    class _Unused:
        def __init__(self, *a, **k):
            raise RuntimeError("stubbed class must not be instantiated on the hot path")

    fm.AdaptiveSoftmax = _Unused
    fm.BaseLayer = _Unused
    fm.LayerDropModuleList = nn.ModuleList
    from fairseq.modules.sinusoidal_positional_embedding import SinusoidalPositionalEmbedding

    fm.SinusoidalPositionalEmbedding = SinusoidalPositionalEmbedding
    from fairseq.modules.learned_positional_embedding import LearnedPositionalEmbedding

    fm.LearnedPositionalEmbedding = LearnedPositionalEmbedding
    from fairseq.modules.positional_embedding import PositionalEmbedding

    fm.PositionalEmbedding = PositionalEmbedding
    ca = types.ModuleType("fairseq.modules.checkpoint_activations")
    ca.checkpoint_wrapper = lambda m, **kw: m
    sys.modules["fairseq.modules.checkpoint_activations"] = ca
    tm = sys.modules["fairseq.models.transformer"]

    class TransformerConfig:  # the NS cfg built by gen_golden is already in the nested dataclass shape
        @staticmethod
        def from_namespace(a):
            return a

    tm.TransformerConfig = TransformerConfig
    tm.Linear = lambda i, o, bias=True: nn.Linear(i, o, bias)
    tm.TransformerModelBase = object
    cu = types.ModuleType("fairseq.checkpoint_utils")
    sys.modules["fairseq.checkpoint_utils"] = cu
    f.checkpoint_utils = cu
    s2t = types.ModuleType("fairseq.models.speech_to_text")
    s2t.S2TTransformerEncoder = object
    sys.modules["fairseq.models.speech_to_text"] = s2t
    sys.modules["fairseq.models.text_to_speech"].TTSTransformerDecoder = object

    # VariancePredictor: run the reference's own class body (fairseq/models/text_to_speech/fastspeech2.py:117-151)
    src = open(REF + "/fairseq/fairseq/models/text_to_speech/fastspeech2.py").read()
    tree = ast.parse(src)
    node = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "VariancePredictor"][0]
    ns = {"nn": nn, "FairseqDropout": fm.FairseqDropout}
    exec(compile(ast.Module(body=[node], type_ignores=[]), "fastspeech2.py:VariancePredictor", "exec"), ns)
    fs2 = types.ModuleType("fairseq.models.text_to_speech.fastspeech2")
    fs2.VariancePredictor = ns["VariancePredictor"]
    sys.modules["fairseq.models.text_to_speech.fastspeech2"] = fs2
    # agent/tts/codehifigan.py imports cleanly once the two modules above exist
    ag = types.ModuleType("agent")
    ag.__path__ = [REF + "/agent"]
    sys.modules["agent"] = ag
    agt = types.ModuleType("agent.tts")
    agt.__path__ = [REF + "/agent/tts"]
    sys.modules["agent.tts"] = agt
    _full = True

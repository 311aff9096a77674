"""The SimulEval agent API the reference agents subclass.

If `simuleval` is importable the real classes are re-exported, so the agents in
streamspeech_b200.agent plug into `simuleval --agent ...` unchanged.  Otherwise (this image lacks
simuleval's yt_dlp/pydub/tornado dependencies) API-identical mirrors are defined, following
SimulEval/simuleval/agents/{agent,states,actions}.py and data/segments.py field for field.
"""
from __future__ import annotations

try:  # pragma: no cover - not installable in the build image
    from simuleval.agents import SpeechToSpeechAgent, SpeechToTextAgent  # type: ignore
    from simuleval.agents.actions import Action, ReadAction, WriteAction  # type: ignore
    from simuleval.agents.states import AgentStates  # type: ignore
    from simuleval.data.segments import EmptySegment, Segment, SpeechSegment, TextSegment  # type: ignore
    from simuleval.utils import entrypoint  # type: ignore
    from simuleval.utils.agent import EVALUATION_SYSTEM_LIST  # type: ignore

    HAVE_SIMULEVAL = True
except Exception:  # noqa: BLE001
    HAVE_SIMULEVAL = False
    import json
    from argparse import ArgumentParser, Namespace
    from dataclasses import dataclass, field
    from inspect import signature
    from typing import List, Optional, Union

    @dataclass
    class Segment:  # data/segments.py:11-25
        index: int = 0
        content: list = field(default_factory=list)
        finished: bool = False
        is_empty: bool = False
        data_type: str = None

        def json(self) -> str:
            return json.dumps({a: v for a, v in self.__dict__.items()})

        @classmethod
        def from_json(cls, json_string: str):
            return cls(**json.loads(json_string))

    @dataclass
    class EmptySegment(Segment):
        is_empty: bool = True

    @dataclass
    class TextSegment(Segment):
        content: str = ""
        data_type: str = "text"

    @dataclass
    class SpeechSegment(Segment):
        sample_rate: int = -1
        data_type: str = "speech"

    class Action:  # agents/actions.py
        def is_read(self) -> bool:
            raise NotImplementedError

    class ReadAction(Action):
        def is_read(self) -> bool:
            return True

        def __repr__(self) -> str:
            return "ReadAction()"

    @dataclass
    class WriteAction(Action):
        content: Union[str, List[float], Segment]
        finished: bool

        def is_read(self) -> bool:
            return False

    class AgentStates:  # agents/states.py
        def __init__(self) -> None:
            self.reset()

        def reset(self) -> None:
            self.source = []
            self.target = []
            self.source_finished = False
            self.target_finished = False
            self.source_sample_rate = 0
            self.target_sample_rate = 0

        def update_source(self, segment: Segment):
            self.source_finished = segment.finished
            if isinstance(segment, EmptySegment):
                return
            elif isinstance(segment, TextSegment):
                self.source.append(segment.content)
            elif isinstance(segment, SpeechSegment):
                self.source += segment.content
                self.source_sample_rate = segment.sample_rate
            else:
                raise NotImplementedError

        def update_target(self, segment: Segment):
            self.target_finished = segment.finished
            if not self.target_finished:
                if isinstance(segment, EmptySegment):
                    return
                elif isinstance(segment, TextSegment):
                    self.target.append(segment.content)
                elif isinstance(segment, SpeechSegment):
                    self.target += segment.content
                    self.target_sample_rate = segment.sample_rate
                else:
                    raise NotImplementedError

    SEGMENT_TYPE_DICT = {"text": TextSegment, "speech": SpeechSegment}

    class GenericAgent:  # agents/agent.py:18-176
        source_type = None
        target_type = None

        def __init__(self, args: Optional[Namespace] = None) -> None:
            if args is not None:
                self.args = args
            assert self.source_type
            assert self.target_type
            self.device = "cpu"
            self.states = self.build_states()
            self.reset()

        def build_states(self) -> AgentStates:
            return AgentStates()

        def reset(self) -> None:
            self.states.reset()

        def policy(self, states: Optional[AgentStates] = None) -> Action:
            raise NotImplementedError

        def push(self, source_segment: Segment, states: Optional[AgentStates] = None) -> None:
            if states is None:
                states = self.states
            states.update_source(source_segment)

        def pop(self, states: Optional[AgentStates] = None) -> Segment:
            if len(signature(self.policy).parameters) == 0:
                is_stateless = False
                if states:
                    raise RuntimeError("Feeding states to stateful agents.")
            else:
                is_stateless = True
            if states is None:
                states = self.states
            if states.target_finished:
                return EmptySegment(finished=True)
            action = self.policy(states) if is_stateless else self.policy()
            if not isinstance(action, Action):
                raise RuntimeError(f"The return value of {self.policy.__qualname__} is not an {Action.__qualname__} instance")
            if action.is_read():
                return EmptySegment()
            if isinstance(action.content, Segment):
                return action.content
            segment = SEGMENT_TYPE_DICT[self.target_type](index=0, content=action.content, finished=action.finished)
            states.update_target(segment)
            return segment

        def pushpop(self, segment: Segment, states: Optional[AgentStates] = None) -> Segment:
            self.push(segment, states)
            return self.pop(states)

        @staticmethod
        def add_args(parser: ArgumentParser):
            pass

        @classmethod
        def from_args(cls, args):
            return cls(args)

        def to(self, device: str, *args, **kwargs) -> None:
            pass

    class SpeechToTextAgent(GenericAgent):
        source_type: str = "speech"
        target_type: str = "text"

    class SpeechToSpeechAgent(GenericAgent):
        source_type: str = "speech"
        target_type: str = "speech"

    EVALUATION_SYSTEM_LIST = []  # simuleval/utils/agent.py:20

    def entrypoint(klass):  # simuleval/utils/__init__.py:10-12
        EVALUATION_SYSTEM_LIST.append(klass)
        return klass

"""bauklank-audio-stretch_b200: the Signalsmith-Stretch hot path of hanskerkhof/BAUKLANK-audio-stretch as
hand-written CUDA for B200 (sm_100a), behind the reference's own operator surface.

Import name: ``bauklank_audio_stretch_b200`` (the repo-root shim maps it onto this directory, whose name carries a
hyphen as the project layout asks).
"""
from ._capi import load_library, Segment, EXPORTS  # noqa: F401
from .batch import BatchStretch, KioskDrive, StreamingDrive, TableDrive, TraceDrive, segment, trace_events  # noqa: F401
from .worklet import WorkletTimeline, ControllerMapper  # noqa: F401
from .engine import StretchEngine  # noqa: F401
from . import shard  # noqa: F401
from . import audio  # noqa: F401
from .audio import decode_audio  # noqa: F401

__all__ = ["StretchEngine", "BatchStretch", "KioskDrive", "StreamingDrive", "TableDrive", "TraceDrive", "trace_events", "WorkletTimeline", "ControllerMapper", "segment", "load_library", "Segment", "EXPORTS", "shard", "audio", "decode_audio"]

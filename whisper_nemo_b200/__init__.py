"""B200-native diarization hot path of johnx102/whisper-nemo.

`ClusteringDiarizer` is the drop-in for the NeMo class the reference drives from
diarize.py:200-201 / nemo_process.py:31-32.  Everything numerical runs in libb200d.so
(hand-written sm_100a CUDA behind the C ABI of include/b200d.h); there is no CPU fallback.
"""
from .config import DiarConfig, create_config, load_config  # noqa: F401

__version__ = "0.1.0"


def __getattr__(name):
    # torch / CUDA are only touched when the diarizer is actually requested
    if name in ("ClusteringDiarizer", "NeuralDiarizer"):
        from . import diarizer

        return getattr(diarizer, name)
    raise AttributeError(name)

"""Config boundary: the reference's YAML schema, unchanged.

The reference builds its config with OmegaConf (helpers.py:252-303:
`OmegaConf.load("nemo_msdd_configs/diar_infer_telephonic.yaml")` followed by
attribute-style overrides).  omegaconf is not installed in this image, so
`DiarConfig` is a small attribute-dict that behaves like a DictConfig for every
access pattern the hot path uses (`cfg.a.b`, `cfg.get("k", d)`, `cfg["k"]`,
`"k" in cfg`, assignment), treats OmegaConf's mandatory marker `???` as
missing, and accepts a real DictConfig (duck-typed) when one is handed in.
"""
import copy
import os
from typing import Any

import yaml

MISSING = "???"
_CONF_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "conf")


class MissingMandatoryValue(KeyError):
    pass


class DiarConfig(dict):
    """dict with attribute access; nested dicts are wrapped on the way in."""

    def __init__(self, data=None):
        super().__init__()
        if data is not None:
            for k, v in dict(data).items():
                self[k] = v

    @staticmethod
    def _wrap(v):
        if isinstance(v, DiarConfig):
            return v
        if isinstance(v, dict):
            return DiarConfig(v)
        if hasattr(v, "items") and hasattr(v, "keys") and not isinstance(v, (str, bytes)):  # DictConfig
            return DiarConfig({k: v[k] for k in v.keys()})
        if type(v).__name__ == "ListConfig":
            return [DiarConfig._wrap(x) for x in v]
        return v

    def __setitem__(self, k, v):
        super().__setitem__(k, self._wrap(v))

    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)
        try:
            v = self[k]
        except KeyError:
            raise AttributeError(k) from None
        if isinstance(v, str) and v == MISSING:
            raise MissingMandatoryValue(f"Missing mandatory value: {k}")
        return v

    def __setattr__(self, k, v):
        self[k] = v

    def get(self, k, default=None):
        v = super().get(k, default)
        if isinstance(v, str) and v == MISSING:
            return default
        return v

    def __deepcopy__(self, memo):
        return DiarConfig({k: copy.deepcopy(v, memo) for k, v in self.items()})

    def to_dict(self):
        return {k: (v.to_dict() if isinstance(v, DiarConfig) else v) for k, v in self.items()}


def load_config(path_or_domain: str) -> DiarConfig:
    """Load a diar_infer_*.yaml.  Accepts a file path or one of the reference's domain
    names ("telephonic" | "meeting" | "general", helpers.py:253)."""
    path = path_or_domain
    if path_or_domain in ("telephonic", "meeting", "general"):
        path = os.path.join(_CONF_DIR, f"diar_infer_{path_or_domain}.yaml")
    with open(path, "r") as f:
        return DiarConfig(yaml.safe_load(f))


def as_config(cfg: Any) -> DiarConfig:
    """Accept DiarConfig / dict / OmegaConf DictConfig."""
    if isinstance(cfg, DiarConfig):
        return cfg
    return DiarConfig._wrap(cfg if isinstance(cfg, dict) else cfg)


def create_config(output_dir: str, domain_type: str = "telephonic", manifest_entry: dict = None) -> DiarConfig:
    """The reference's helpers.create_config (helpers.py:252-303), for users who switch over:
    loads the domain YAML, writes the one-line input manifest for <output_dir>/mono_file.wav,
    and applies the same overrides."""
    import json

    config = load_config(domain_type)
    data_dir = os.path.join(output_dir, "data")
    os.makedirs(data_dir, exist_ok=True)
    meta = {
        "audio_filepath": os.path.join(output_dir, "mono_file.wav"),
        "offset": 0,
        "duration": None,
        "label": "infer",
        "text": "-",
        "rttm_filepath": None,
        "uem_filepath": None,
    }
    if manifest_entry:
        meta.update(manifest_entry)
    with open(os.path.join(data_dir, "input_manifest.json"), "w") as fp:
        json.dump(meta, fp)
        fp.write("\n")
    config.num_workers = 0
    config.diarizer.manifest_filepath = os.path.join(data_dir, "input_manifest.json")
    config.diarizer.out_dir = output_dir
    config.diarizer.speaker_embeddings.model_path = "titanet_large"
    config.diarizer.oracle_vad = False
    config.diarizer.clustering.parameters.oracle_num_speakers = False
    config.diarizer.vad.model_path = "vad_multilingual_marblenet"
    config.diarizer.vad.parameters.onset = 0.8
    config.diarizer.vad.parameters.offset = 0.6
    config.diarizer.vad.parameters.pad_offset = -0.05
    config.diarizer.msdd_model.model_path = "diar_msdd_telephonic"
    return config

"""YAML config loading: base config overridden by the dataset's file (utils/merge_yaml.py of the reference)."""
import yaml


def deep_merge(dct, merge_dct):
    """Recursive in-place merge: nested dicts are merged key by key, anything else is replaced."""
    for key, value in merge_dct.items():
        if isinstance(dct.get(key), dict) and isinstance(value, dict):
            deep_merge(dct[key], value)
        else:
            dct[key] = value


def load_and_merge_yaml(base_filepath, override_filepath):
    with open(base_filepath, "r") as f:
        cfg = yaml.safe_load(f) or {}
    with open(override_filepath, "r") as f:
        deep_merge(cfg, yaml.safe_load(f) or {})
    return cfg

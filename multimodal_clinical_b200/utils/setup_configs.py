"""``setup_configs()``: parse ``--dir``, merge ``utils/base_cfg.yaml`` with ``<dir>/<dir>.yaml`` into an
argparse.Namespace and seed everything (utils/setup_configs.py of the reference).  Config files are looked
up in the working directory first (the reference's layout) and then inside this package."""
import argparse
import os

from .lightning_compat import seed_everything
from .merge_yaml import load_and_merge_yaml

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _find(*parts):
    for root in (os.getcwd(), _PKG):
        p = os.path.join(root, *parts)
        if os.path.exists(p):
            return p
    raise FileNotFoundError(os.path.join(*parts))


def setup_configs(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument("--dir", type=str, default=None)
    args, _ = parser.parse_known_args(argv)
    if not args.dir:
        raise NotImplementedError("No directory provided, please specify flag --dir")
    cfg = load_and_merge_yaml(_find("utils", "base_cfg.yaml"), _find(args.dir, args.dir + ".yaml"))
    for key, val in cfg.items():
        setattr(args, key, val)
    seed_everything(args.seed, workers=True)
    return args

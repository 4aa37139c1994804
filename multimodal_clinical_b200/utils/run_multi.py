"""Driver shared by mustard/run_training.py and avmnist/run_training.py of the reference (stand-alone scripts there):
``--config <yaml>`` (default: the packaged one) -> args, seed, loaders, ``pl.Trainer(precision="32")``, fit, test."""
import argparse
import os

import torch
import yaml
from torch.utils.data import DataLoader

from .lightning_compat import pl, seed_everything


def load_args(argv, default_yaml):
    parser = argparse.ArgumentParser()
    parser.add_argument("--config", "--configs", type=str, default=None)
    parser.add_argument("--dir", type=str, default=None)
    args, _ = parser.parse_known_args(argv)
    with open(args.config or default_yaml, "r") as fh:
        for key, val in yaml.safe_load(fh).items():
            setattr(args, key, val)
    seed_everything(args.seed, workers=True)
    return args


def fit_and_test(args, model, datasets):
    if not torch.cuda.is_available():
        raise NotImplementedError("It is not advised to train without a GPU")
    train_dataset, val_dataset, test_dataset = datasets
    kw = dict(batch_size=args.batch_size, num_workers=0)
    model = model.cuda()
    trainer = pl.Trainer(strategy="auto", max_epochs=args.num_epochs, logger=None, deterministic=True, default_root_dir="ckpts/",
                         precision="32", num_sanity_val_steps=0, log_every_n_steps=10)
    trainer.fit(model, train_dataloaders=DataLoader(train_dataset, **kw), val_dataloaders=DataLoader(val_dataset, shuffle=False, **kw))
    trainer.test(model, dataloaders=DataLoader(test_dataset, shuffle=False, **kw))
    return trainer


def packaged_yaml(pkg_file, name):
    return os.path.join(os.path.dirname(os.path.abspath(pkg_file)), name)

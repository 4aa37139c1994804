"""``python -m multimodal_clinical_b200.main --dir {cremad,food101,enrico,mustard,avmnist}`` (main.py of the reference; the
last two are stand-alone ``run_training.py`` scripts there and take ``--config <yaml>``)."""
import argparse


def main(argv=None):
    parser = argparse.ArgumentParser(description="which directory to run")
    parser.add_argument("--dir", type=str, default=None, help="directory to run")
    arg, _ = parser.parse_known_args(argv)
    if arg.dir == "cremad":
        from .cremad.run_trainer import run_training
    elif arg.dir == "food101":
        from .food101.run_training import run_training
    elif arg.dir == "enrico":
        from .enrico.run_training import run_training
    elif arg.dir == "mustard":
        from .mustard.run_training import run_training
    elif arg.dir == "avmnist":
        from .avmnist.run_training import run_training
    else:
        raise NotImplementedError("Please specify a directory to run")
    return run_training(argv)


if __name__ == "__main__":
    main()

"""``run_trainer(args, model, train_loader, val_loader, test_loader, overfit_batches=0)``: fit, reload the
best checkpoint by ``val_epoch/val_avg_acc``, test (utils/run_trainer.py of the reference).  Uses
PyTorch-Lightning when installed, the built-in trainer of lightning_compat otherwise."""
from datetime import datetime

import torch

from .lightning_compat import pl


def run_trainer(args, model, train_loader, val_loader, test_loader, overfit_batches=0):
    wandb_logger = None
    if args.use_wandb:
        from pytorch_lightning.loggers import WandbLogger      # needs Lightning + wandb, as in the reference
        wandb_logger = WandbLogger(group=args.group_name)
        file_name = wandb_logger.experiment.name + "_best"
        wandb_logger.log_hyperparams(args)
    else:
        file_name = datetime.now().strftime("%Y%m%d_%H%M%S") + "_best"
    lr_logger = pl.callbacks.LearningRateMonitor(logging_interval='epoch')
    checkpoint_logger = pl.callbacks.ModelCheckpoint(
        dirpath=args.data_path + "_ckpts/" + args.group_name + "/", filename=file_name, save_top_k=1,
        monitor="val_epoch/val_avg_acc", mode="max")
    if not torch.cuda.is_available():
        raise NotImplementedError("It is not advised to train without a GPU")
    model = model.cuda()
    trainer = pl.Trainer(strategy="auto", max_epochs=args.num_epochs, logger=wandb_logger, deterministic=True,
                         default_root_dir="ckpts/", precision=getattr(args, "precision", "bf16-mixed"),
                         num_sanity_val_steps=0, log_every_n_steps=30, callbacks=[lr_logger, checkpoint_logger],
                         overfit_batches=overfit_batches)
    trainer.fit(model, train_dataloaders=train_loader, val_dataloaders=val_loader)
    model.load_state_dict(torch.load(checkpoint_logger.best_model_path)["state_dict"])
    trainer.test(model, dataloaders=test_loader)
    return trainer

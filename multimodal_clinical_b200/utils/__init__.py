"""Counterparts of the reference's utils/ package (EMA, BaseModel, run_trainer, setup_configs)."""

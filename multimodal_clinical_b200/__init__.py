"""B200-native late-fusion training step (sm_100a CUDA behind a C ABI) with the module / config API of
Nano1337/multimodal-clinical: ``main --dir``, ``utils.BaseModel``, the per-dataset ``FusionNet`` heads,
``existing_algos.{QMF, OGM_GE}`` and ``utils.EMA``.  See DESIGN.md and INTEGRATION.md."""
__version__ = "0.1.0"

def get_model(args):
    """Model types of mustard/run_training.py on the fused path (jlogits: mean fusion of the heads' logits)."""
    if args.model_type == "jlogits":
        from .joint_model import MultimodalMustardModel
    else:
        raise NotImplementedError("Model type not implemented")
    return MultimodalMustardModel(args)

def get_model(args):
    """enrico/__init__.py of the reference; types on the fused path."""
    if args.model_type == "jlogits":
        from .joint_model import MultimodalEnricoModel
    else:
        raise NotImplementedError("Model type not implemented")
    return MultimodalEnricoModel(args)

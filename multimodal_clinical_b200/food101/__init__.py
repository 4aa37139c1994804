def get_model(args):
    """food101/__init__.py of the reference; types on the fused path."""
    if args.model_type == "jlogits":
        from .joint_model import MultimodalFoodModel
    elif args.model_type == "ogm_ge":
        from .joint_model_ogm_ge import MultimodalFoodModel
    elif args.model_type == "qmf":
        from .joint_model_qmf import MultimodalFoodModel
    else:
        raise NotImplementedError("Model type not implemented")
    return MultimodalFoodModel(args)

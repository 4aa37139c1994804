def get_model(args):
    """Model factory keyed on args.model_type (cremad/__init__.py of the reference).  The types on the
    fused path are implemented; the others name reference models outside it."""
    if args.model_type == "jlogits":
        from .joint_model import MultimodalCremadModel
    elif args.model_type == "ogm_ge":
        from .joint_model_ogm_ge import MultimodalCremadModel
    elif args.model_type == "ensemble_ogm_ge":
        from .ensemble_model_noised import MultimodalCremadModel
    elif args.model_type == "qmf":
        from .joint_model_qmf import MultimodalCremadModel
    elif args.model_type == "qmf_ablate":
        from .joint_model_qmf_ablate import MultimodalCremadModel
    elif args.model_type == "qmf_ablate_Ljoint":
        from .joint_model_qmf_ablate_Ljoint import MultimodalCremadModel
    elif args.model_type == "qmf_ablate_Lunimodal":
        from .joint_model_qmf_ablate_Lunimodal import MultimodalCremadModel
    elif args.model_type == "ogm_ge_lreg":
        from .joint_model_ogm_ge_lreg import MultimodalCremadModel
    else:
        raise NotImplementedError("Model type not implemented")
    return MultimodalCremadModel(args)

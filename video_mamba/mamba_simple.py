from videomamba_b200.mixer import InferenceParamsLike, Mamba

__all__ = ["InferenceParamsLike", "Mamba"]

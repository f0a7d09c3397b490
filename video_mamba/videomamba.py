from videomamba_b200.model import PretrainVideoMamba, build_videomamba, load_state_dict

__all__ = ["PretrainVideoMamba", "build_videomamba", "load_state_dict"]

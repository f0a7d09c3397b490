from videomamba_b200.refiner import BiMambaRefinerBlock

__all__ = ["BiMambaRefinerBlock"]

from videomamba_b200.determinism import (DeterminismConfig, add_determinism_args,
                                         configure_determinism, configure_determinism_from_args)

__all__ = ["DeterminismConfig", "add_determinism_args", "configure_determinism",
           "configure_determinism_from_args"]

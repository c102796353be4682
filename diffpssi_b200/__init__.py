"""diffpssi_b200 -- B200-native engine for DiffPSSi's time-stepping hot path, behind the reference's Pss API.

    from diffpssi_b200 import PowerSystemSimulation as Pss
    sim = Pss(parallel_sims=4096, sim_time=10, time_step=0.005, solver='heun', grid_data=grid)   # backend='cuda'
"""
from .components import (Bus, Line, Transformer, Load, Shunt, SynchMachine, SEXS, TGOV1, STAB1,  # noqa: F401
                         ScEvent, ParamEvent, PIController)
from .pss import PowerSystemSimulation, Pss  # noqa: F401
from .optimization import PowerSystemOptimization, CustomBFGSREALOptimizer, GraphedStep  # noqa: F401
from . import engine  # noqa: F401

__version__ = '0.1.0'

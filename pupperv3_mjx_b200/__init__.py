"""pupperv3_mjx_b200 -- B200-native batched PupperV3Env step behind the reference's Brax PipelineEnv API."""
from . import abi, config, domain_randomization, mjcf, obstacles, prng, system, utils  # noqa: F401
from .environment import PupperV3Env, State  # noqa: F401

__all__ = ["PupperV3Env", "State", "abi", "config", "domain_randomization", "mjcf", "obstacles", "prng", "system", "utils"]

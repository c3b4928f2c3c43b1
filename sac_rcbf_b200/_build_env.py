"""build_env(args) -- name -> env mapping of the reference's build_env.py:1-16."""
from .envs import SimulatedCarsEnv, UnicycleEnv


def build_env(args, **kw):
    if args.env_name == 'Unicycle':
        return UnicycleEnv(**kw)
    if args.env_name == 'SimulatedCars':
        return SimulatedCarsEnv(**kw)
    raise Exception('Env {} not supported!'.format(args.env_name))

"""B200-native batched backend for the GridworldCtf step loop of g-nightingale/marl-ctf-development.

Public surface:
    GridworldCtfGPU   batched env (B envs per GPU), device tensors in / out
    GridworldCtf      single-env view with the reference's exact methods (drop-in for ppo.py / utils.py)
    compile_config    (AGENT_CONFIG, SCENARIO, kwargs) -> ctf_config_t
    experiment_env_config / experiment_names   the reference's nine experiment configs as data
"""
from .config import METRIC_NAMES, compile_config, env_dims  # noqa: F401
from .scenario_io import experiment_env_config, experiment_names, load_env_config  # noqa: F401


def __getattr__(name):
    # env.py imports torch; keep `import marl_ctf_development_b200` light for host-only users
    if name in ("GridworldCtfGPU", "GridworldCtf", "metrics_dict"):
        from . import env

        return getattr(env, name)
    raise AttributeError(name)

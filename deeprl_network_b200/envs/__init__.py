from .cacc_env import CACCEnv  # noqa: F401

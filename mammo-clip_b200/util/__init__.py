from .global_env import GlobalEnv  # noqa: F401

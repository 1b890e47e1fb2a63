"""paddlenlp/utils/tools.py: get_env_device() — this build has exactly one device back-end."""


def get_env_device() -> str:
    return "gpu"

from .images import *  # noqa: F401,F403

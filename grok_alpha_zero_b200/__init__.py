"""B200-native self-play engine: drop-in for Grok_Alpha_Zero's MCTS + evaluation hot path."""
__all__ = ["Engine"]


def __getattr__(name):
    if name == "Engine":
        from .engine import Engine
        return Engine
    raise AttributeError(name)

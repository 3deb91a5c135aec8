"""B200-native self-play engine: drop-in for Grok_Alpha_Zero's MCTS + evaluation hot path.

Reference-named entry points (lazy imports; nothing here touches CUDA until an engine or network is created):
    from grok_alpha_zero_b200.MCTS import MCTS
    from grok_alpha_zero_b200.MCTS_Gumbel import MCTS_Gumbel
    from grok_alpha_zero_b200.games import Gomoku, Connect4, TicTacToe
    from grok_alpha_zero_b200.Client_Server import Parallelized_Session, start_server, create_shared_memory
    from grok_alpha_zero_b200.Self_Play import run_self_play
"""
__all__ = ["Engine", "Net", "GazSession"]


def __getattr__(name):
    if name == "Engine":
        from .engine import Engine
        return Engine
    if name == "Net":
        from .net import Net
        return Net
    if name == "GazSession":
        from .session import GazSession
        return GazSession
    raise AttributeError(name)

class ReplayBuffer:
    def __init__(self, *a, **k): raise NotImplementedError
class ConcatReplayBuffer(ReplayBuffer):
    pass

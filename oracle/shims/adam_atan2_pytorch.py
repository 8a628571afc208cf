class MuonAdamAtan2:  # off the hot path
    def __init__(self, *a, **k): raise NotImplementedError

class PoPE:
    def __init__(self, *a, **k): raise NotImplementedError
class AxialPoPE:
    def __init__(self, *a, **k): raise NotImplementedError
def flash_attn_with_pope(*a, **k): raise NotImplementedError

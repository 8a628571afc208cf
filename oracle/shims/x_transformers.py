class Decoder:
    def __init__(self, *a, **k): raise NotImplementedError

"""`dreamer4.mocks` (reference dreamer4/mocks.py)."""
from dreamer4_b200.mocks import MockDictEnv, MockEnv

"""`dreamer4.env` names on this path (reference dreamer4/env.py:353-553 DynamicsWorldModelWrapper)."""
from dreamer4_b200.env import DynamicsWorldModelWrapper

"""`dreamer4.trainers` names on this path (reference dreamer4/trainers.py:1330-1468 DreamTrainer, 1472-1790 SimTrainer)."""
from dreamer4_b200.trainer import DreamTrainer, SimTrainer

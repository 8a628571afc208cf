"""dreamer4_b200 — B200-native imagination hot path of Dreamer 4 (generate + learn_from_experience).

Public names mirror the reference package (`dreamer4/__init__.py:1-15`, `dreamer4/dreamer4.py`) for the
classes on this path."""
from .experience import Actions, Embeds, Experience, Predictions, combine_experiences
from .dynamics import DynamicsWorldModel, ModelConfig, exists, default
from .trainer import DreamTrainer, SimTrainer
from .env import DynamicsWorldModelWrapper
from .tokenizer import AxialSpaceTimeTransformer, VideoTokenizer, TokenizerConfig
from .registry import register_activation, register_reward_encoder

__all__ = ['Actions', 'Embeds', 'Predictions', 'Experience', 'combine_experiences', 'DynamicsWorldModel', 'DreamTrainer', 'SimTrainer', 'DynamicsWorldModelWrapper', 'ModelConfig', 'exists', 'default',
           'VideoTokenizer', 'TokenizerConfig', 'AxialSpaceTimeTransformer', 'register_activation', 'register_reward_encoder']

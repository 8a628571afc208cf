"""`dreamer4` import name over the B200-native implementation (dreamer4_b200): the names the reference package exports
(reference dreamer4/__init__.py:1-15), so `from dreamer4 import DynamicsWorldModel, VideoTokenizer, ...` resolves to the CUDA path.
Trainers that belong to the reference's pre-training side (VideoTokenizerTrainer, BehaviorCloneTrainer) are outside this path
(SURVEY.md section 8): importing them raises ImportError naming that."""
from dreamer4.dreamer4 import (
    VideoTokenizer,
    DynamicsWorldModel,
    AxialSpaceTimeTransformer,
    register_activation,
    register_reward_encoder,
)
from dreamer4.trainers import DreamTrainer, SimTrainer
from dreamer4.env import DynamicsWorldModelWrapper

_OUT_OF_SCOPE = ('VideoTokenizerTrainer', 'BehaviorCloneTrainer')


def __getattr__(name):
    if name in _OUT_OF_SCOPE:
        raise ImportError(f'dreamer4.{name}: the pre-training trainers are outside the B200 hot path this package builds '
                          '(generate / learn_from_experience and their callers; SURVEY.md section 8)')
    raise AttributeError(f'module {__name__!r} has no attribute {name!r}')

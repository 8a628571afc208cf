import pickle
from functools import wraps
from pathlib import Path
import torch

def save_load(klass = None, **_):
    """Records init kwargs as `_config`; adds .save/.load/.init_and_load (hot-path subset)."""
    def decorate(klass):
        orig_init = klass.__init__

        @wraps(orig_init)
        def __init__(self, *args, **kwargs):
            self._config = (args, kwargs)
            orig_init(self, *args, **kwargs)

        def save(self, path, overwrite = True):
            path = Path(path)
            assert overwrite or not path.exists()
            torch.save(dict(model = self.state_dict(), config = pickle.dumps(self._config)), str(path))

        def load(self, path, strict = True):
            pkg = torch.load(str(path), map_location = 'cpu', weights_only = False)
            self.load_state_dict(pkg['model'], strict = strict)

        @classmethod
        def init_and_load(cls, path, strict = True):
            pkg = torch.load(str(path), map_location = 'cpu', weights_only = False)
            args, kwargs = pickle.loads(pkg['config'])
            model = cls(*args, **kwargs)
            model.load_state_dict(pkg['model'], strict = strict)
            return model

        klass.__init__ = __init__
        klass.save, klass.load, klass.init_and_load = save, load, init_and_load
        return klass
    return decorate(klass) if klass is not None else decorate

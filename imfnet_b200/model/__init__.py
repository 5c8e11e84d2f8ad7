"""Model registry with the reference's `load_model(name) -> class` contract (/root/reference/model/__init__.py:16-30)."""
import logging

from . import resunet as _resunets

MODELS = [getattr(_resunets, a) for a in dir(_resunets) if 'Net' in a and isinstance(getattr(_resunets, a), type)]


def load_model(name):
    """Class of the model called `name`, or None (after logging the options) when it is unknown."""
    table = {cls.__name__: cls for cls in MODELS}
    if name not in table:
        logging.info(f'Invalid model index. You put {name}. Options are:')
        for cls in MODELS:
            logging.info('\t* {}'.format(cls.__name__))
        return None
    return table[name]

"""Process-wide weights epoch.

Weight-derived caches (equalised-LR conv weights, the concatenated style affines, the folded attention weights of stage W,
captured CUDA graphs of the forward) are keyed on parameter storage + version counters.  A CUDA-graph replay of a training
step, or an in-place ``copy_`` performed by a captured optimizer, updates the parameters WITHOUT touching those counters, so
every such cache key also carries this epoch; whoever changes weights behind autograd's back bumps it
(``Trainer.step_graphed``, ``Generator.load_state_dict``)."""

_EPOCH = [0]


def weights_epoch() -> int:
    return _EPOCH[0]


def bump_weights_epoch() -> int:
    _EPOCH[0] += 1
    return _EPOCH[0]

"""Plugging into the reference's registries (registry/register.py:9-45, models/builder.py:40).

The reference selects losses by name: ``loss_decode=dict(type='CrossEntropyLoss', ...)`` ->
``build_loss(cfg)`` -> ``LOSS.get(type)(**cfg)`` (models/builder.py:262-283). ``Register.register``
refuses duplicate keys (register.py:15-18), so a drop-in either registers under new names or
overwrites the class-level ``_storage`` entry. Both are offered.
"""
from .losses import CrossEntropyLoss, DiceLoss, LovaszLoss, TverskyLoss

B200_LOSSES = {'CrossEntropyLoss': CrossEntropyLoss, 'DiceLoss': DiceLoss, 'TverskyLoss': TverskyLoss, 'LovaszLoss': LovaszLoss}


def install(loss_registry, override=True, prefix='B200'):
    """Make the fused losses reachable through the reference's ``LOSS`` registry.

    override=True : ``LOSS._storage['CrossEntropyLoss'|'DiceLoss'|'TverskyLoss'|'LovaszLoss']`` now build this package's classes, so existing
                    network configs run on the fused kernels unchanged.
    always        : also registers ``B200CrossEntropyLoss`` / ``B200DiceLoss`` for configs that opt in by name.
    Returns the dict of names installed.
    """
    storage = loss_registry._storage
    installed = {}
    for name, cls in B200_LOSSES.items():
        alias = prefix + name
        if alias not in storage:
            storage[alias] = cls
        installed[alias] = cls
        if override:
            storage[name] = cls
            installed[name] = cls
    return installed


def build_loss(cfg):
    """Stand-alone equivalent of models/builder.py:262-283 restricted to this package's losses."""
    if not isinstance(cfg, dict):
        raise TypeError('The loss cfg must be a dict')
    if 'type' not in cfg:
        raise KeyError('The loss cfg dict must contain the key "type"')
    cfg_ = cfg.copy()
    loss_type = cfg_.pop('type')
    name = loss_type[4:] if loss_type.startswith('B200') else loss_type
    if name not in B200_LOSSES:
        raise KeyError(f'Cannot find {loss_type} in LOSS Register !')
    return B200_LOSSES[name](**cfg_)

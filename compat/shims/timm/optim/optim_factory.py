"""`param_groups_layer_decay(model, weight_decay)` as main_lidar_upsampling.py:282 calls it.

Behaviour follows timm's documentation of the function: parameters are numbered in `named_parameters()` order and cut into
"layers" of 12 names; layer i of n gets `lr_scale = layer_decay ** (n - 1 - i)` (0.75 by default); 1-D parameters and the
names in `no_weight_decay_list` get weight decay 0.  One group per (layer, decay / no_decay).  `lr_sched.py:16-20` of the
reference multiplies the scheduled lr by `lr_scale`."""


def add_weight_decay(model, weight_decay=1e-5, no_weight_decay_list=()):
    decay, no_decay = [], []
    for name, p in model.named_parameters():
        if not p.requires_grad:
            continue
        (no_decay if (p.ndim <= 1 or name.endswith(".bias") or name in no_weight_decay_list) else decay).append(p)
    return [{"params": no_decay, "weight_decay": 0.0}, {"params": decay, "weight_decay": weight_decay}]


def param_groups_layer_decay(model, weight_decay=0.05, no_weight_decay_list=(), layer_decay=0.75, end_layer_decay=None, verbose=False):
    skip = set(no_weight_decay_list)
    names = [n for n, _ in model.named_parameters()]
    per_layer = 12
    layer_of = {n: i // per_layer for i, n in enumerate(names)}
    n_layers = (len(names) + per_layer - 1) // per_layer + 1          # trunk chunks + one (empty here) head layer
    scales = [layer_decay ** (n_layers - 1 - i) for i in range(n_layers)]
    groups = {}
    for name, p in model.named_parameters():
        if not p.requires_grad:
            continue
        no_decay = p.ndim == 1 or name in skip
        lid = layer_of[name]
        key = (lid, no_decay)
        if key not in groups:
            groups[key] = {"lr_scale": scales[lid], "weight_decay": 0.0 if no_decay else weight_decay, "params": []}
        groups[key]["params"].append(p)
    return list(groups.values())

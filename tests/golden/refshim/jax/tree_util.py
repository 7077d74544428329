import torch


def tree_leaves(tree, is_leaf=None):
    out = []

    def walk(t):
        if is_leaf is not None and is_leaf(t):
            out.append(t)
        elif isinstance(t, (tuple, list)):
            for x in t:
                walk(x)
        elif isinstance(t, dict):
            for x in t.values():
                walk(x)
        elif getattr(t, "_is_module", False):  # equinox.Module stand-in: fields in definition order
            for x in vars(t).values():
                walk(x)
        elif t is not None:
            out.append(t)

    walk(tree)
    return out


def tree_map(f, tree, *rest):
    if isinstance(tree, (tuple, list)):
        return type(tree)(tree_map(f, t, *[r[i] for r in rest]) for i, t in enumerate(tree))
    return f(tree, *rest)

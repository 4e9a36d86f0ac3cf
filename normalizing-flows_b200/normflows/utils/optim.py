"""normflows/utils/optim.py:4-25."""


def clear_grad(model):
    for param in model.parameters():
        param.grad = None


def set_requires_grad(module, flag):
    for param in module.parameters():
        param.requires_grad = flag


def update_lipschitz(model, n_iterations):
    """normflows/utils/optim.py:28-31: refresh the spectral-norm estimates (power iteration) of every induced-norm layer."""
    from ..nets.lipschitz import InducedNormLinear
    for m in model.modules():
        if isinstance(m, InducedNormLinear):
            m.compute_weight(update=True, n_iterations=n_iterations)

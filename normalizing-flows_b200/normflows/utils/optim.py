"""normflows/utils/optim.py:4-25."""


def clear_grad(model):
    for param in model.parameters():
        param.grad = None


def set_requires_grad(module, flag):
    for param in module.parameters():
        param.requires_grad = flag

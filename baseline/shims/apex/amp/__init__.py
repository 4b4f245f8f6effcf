import contextlib


def initialize(model, optimizer, opt_level="O2", loss_scale="dynamic", **kw):
    return model, optimizer


@contextlib.contextmanager
def scale_loss(loss, optimizer):
    yield loss


def master_params(optimizer):
    for g in optimizer.param_groups:
        for p in g["params"]:
            yield p

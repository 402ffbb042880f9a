"""apex.amp no-op: the kernels already compute in a 16-bit storage type with fp32 accumulation (bf16 by default, no loss scaling
needed; IEEE fp16 through hrviton_b200.ops.set_precision('fp16'))."""
import contextlib


def initialize(models, optimizers=None, opt_level="O1", num_losses=1, **kwargs):
    return (models, optimizers) if optimizers is not None else models


@contextlib.contextmanager
def scale_loss(loss, optimizer, loss_id=0, **kwargs):
    yield loss

"""`MipLRDecay` with the reference's constructor (utils/lr_schedule.py:6-60): a per-step scheduler for a stock
torch optimiser (what `BaseSystem.configure_optimizers` upstream pairs with `torch.optim.Adam`,
systems/base_system.py:81-87).  The rate itself is `systems.base_system.mip_lr_decay`, the function `FlatAdam` evaluates
on the host for its fused update, so both optimiser routes follow the same schedule."""
import torch

from ..systems.base_system import mip_lr_decay


class MipLRDecay(torch.optim.lr_scheduler.LRScheduler):
    def __init__(self, optimizer, lr_init: float, lr_final: float, max_steps: int, lr_delay_steps: int,
                 lr_delay_mult: float):
        self.lr_init, self.lr_final, self.max_steps = lr_init, lr_final, max_steps
        self.lr_delay_steps, self.lr_delay_mult = lr_delay_steps, lr_delay_mult
        super().__init__(optimizer)

    def get_lr(self):
        # one rate for every parameter group; `last_epoch` counts scheduler steps (Lightning: interval='step')
        rate = mip_lr_decay(self.last_epoch, self.lr_init, self.lr_final, self.max_steps, self.lr_delay_steps,
                            self.lr_delay_mult)
        return [rate for _ in self.optimizer.param_groups]
